// kernels.cuh -- the kernels of the explicit FV residual + update path, templated on
// arithmetic mode (see physics.cuh).  Included by kernels_fast.cu (default FMA contraction)
// and kernels_faithful.cu (-fmad=false).
//
//   ghost_kernel  : boundary-condition ghost states + neighbour ghost lines + per-block dt
//                   (generate_ghost_cells SpatialDiscretization.cpp:17-99, the receive side of
//                   the halo exchange :263-269, find_dt main.cpp:19-37)
//   march_kernel  : (march.cuh) one warp strip of one block: TMA-staged state+node rows with halo
//                   columns -> per-cell derived state -> one LDFSS per face (EulerFlux.cpp:121) ->
//                   residual (SpatialDiscretization.cpp:481-1011) -> sum of squares
//                   (main.cpp:39-57) -> Jacobian solve + 1/ycc (main.cpp:307-331) ->
//                   unk += dv and the UpdateState clamps (main.cpp:394-404)
//   test kernels  : the point-wise device functions over arrays (test hooks of the C ABI)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"
#include "physics.cuh"
#include "halo.cuh"
#include "march.cuh"
#include "kernels_common.cuh"
#include "dg.cuh"

namespace par2d {

template <bool FAITHFUL>
__device__ __forceinline__ double block_dt_of(const StepArgs& a, const Gas& g, const BlockDev& b, const double* S) {
    const double dt = block_dt_value(S, b.plane, b.pitch, b.min_face, g.gamRs, g.cfl);
    if (isnan(dt)) atomicOr(a.flags, FLAG_NAN);
    return dt;
}

// ------------------------------------------------------------------ ghost fill
template <bool FAITHFUL>
__global__ void __launch_bounds__(256) ghost_kernel(StepArgs a, Gas g, int write_dt, FinalizeArgs fin, int nfin) {
    if (*a.stop) return;
    // rows 0..nfin-1 of the grid: the residual bookkeeping of the PREVIOUS step (nfin = owned blocks, or 0
    // when nothing is pending), overlapped with this step's ghost fill
    if ((int)blockIdx.y < nfin) {
        __shared__ __align__(16) double stage[FIN_CHUNK * 4];
        if (blockIdx.x == 0) finalize_block(fin, blockIdx.y, stage);
        return;
    }
    if (a.wait_flags != nullptr) {
        // peer-mapped mode: the neighbours' edge lines and every block's dt of this step's state are
        // written into this rank's arena by the other GPUs; wait for their completion flags
        __shared__ int arrived;
        if (threadIdx.x == 0) {
            arrived = wait_peers(a.wait_flags, a.nranks, a.rank, a.wait_seq, a.timeout_ns) ? 1 : 0;
            if (!arrived) atomicOr(a.flags, FLAG_COMM);
        }
        __syncthreads();
    }
    if (a.dt_ready && blockIdx.x == 0 && (int)blockIdx.y == nfin && threadIdx.x == 0) {
        // global dt = max over blocks (main.cpp:270-292) of the table the exchange completed before this launch
        double dt = 0.0;
        for (int t = 0; t < a.nblocks_global; t++) dt = fmax(dt, __ldcg(a.dt_table + t));
        *a.dt_used = dt;
    }
    const BlockDev& b = a.blocks[blockIdx.y - nfin];   // grid.y = [pending finalize rows +] owned blocks, grid.x covers the longest perimeter
    int n = blockIdx.x * blockDim.x + threadIdx.x, side;
    if (n >= 2 * (b.ncx + b.ncy)) return;
    if (n < b.ncx) side = SIDE_BOTTOM;
    else if ((n -= b.ncx) < b.ncy) side = SIDE_RIGHT;
    else if ((n -= b.ncy) < b.ncx) side = SIDE_TOP;
    else { n -= b.ncx; side = SIDE_LEFT; }

    double* S = b.state[a.cur];
    int gc, gr;
    ghost_cell_pos(b.ncx, b.ncy, side, n, gc, gr);
    const long long go = (long long)gr * b.pitch + gc;

    if (write_dt && side == SIDE_BOTTOM && n == 0) a.dt_table[b.gid] = block_dt_of<FAITHFUL>(a, g, b, S);

    if (b.nb_kind[side] == NB_NONE) {
        bc_ghost_fill(a.gas_dev, a.blocks + (blockIdx.y - nfin), S, side, n);
        return;
    }
    double gh[4];
    if (b.nb_kind[side] == NB_LOCAL) {
        // the neighbour's edge line, copied verbatim (SpatialDiscretization.cpp:263-266)
        const BlockDev& o = a.blocks[b.nb_block[side]];
        int c, r;
        edge_cell_pos(o.ncx, o.ncy, b.nb_side[side], n, c, r);
        const double* So = o.state[a.cur];
        const long long oo = (long long)r * o.pitch + c;
#pragma unroll
        for (int v = 0; v < 4; v++) gh[v] = So[v * o.plane + oo];
        gh[3] = fmax(gh[3], 201.0);
    } else {
        const double* line = b.recv[side] + a.recv_off + 4 * (long long)n;
#pragma unroll
        for (int v = 0; v < 4; v++) gh[v] = __ldcg(line + v);   // written by another GPU: read at L2
        gh[3] = fmax(gh[3], 201.0);
    }
#pragma unroll
    for (int v = 0; v < 4; v++) S[v * b.plane + go] = gh[v];
}

// ------------------------------------------------------------------ pack (multi-rank send side)
template <bool FAITHFUL>
__global__ void __launch_bounds__(256) pack_kernel(StepArgs a, Gas g, const SendLine* lines, int nlines, PutArgs put) {
    if (*a.stop) return;
    if (put.peer_mode) {
        // producing event seq: every peer must have produced event seq-1 (see PutArgs)
        __shared__ int arrived;
        if (threadIdx.x == 0) {
            arrived = wait_peers(put.local_flags, put.nranks, put.rank, put.seq - 1, put.timeout_ns) ? 1 : 0;
            if (!arrived) atomicOr(a.flags, FLAG_COMM);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int k = threadIdx.x; k < a.nowned; k += blockDim.x) {
            const BlockDev& b = a.blocks[k];
            const double dt = block_dt_of<FAITHFUL>(a, g, b, b.state[a.cur]);
            if (put.peer_mode) {
                for (int r = 0; r < put.nranks; r++) put.dt_dst[r][put.parity * put.dt_stride + b.gid] = dt;
            } else {
                a.dt_table[b.gid] = dt;
            }
        }
    }
    for (int l = blockIdx.y; l < nlines; l += gridDim.y) {
        const SendLine ln = lines[l];
        const BlockDev& b = a.blocks[ln.block];
        const double* S = b.state[a.cur];
        double* dst = ln.dst + put.parity * ln.parity_stride;
        for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < ln.n; n += gridDim.x * blockDim.x) {
            int c, r;
            edge_cell_pos(b.ncx, b.ncy, ln.side, n, c, r);
            const long long o = (long long)r * b.pitch + c;
            if (put.dg) {
                // second-order branch: the reconstruction at the two Gauss points of this rank's edge face,
                // usend[k + 2*NVAR*j] = point 1, usend[k + NVAR + 2*NVAR*j] = point 2 (SpatialDiscretization.cpp:243-256)
                DgCell cell;
                dg_load_cell(g, b, a.cur, c - COL0, r - 1, cell);
                const double G = dg_gauss();
                double2* d2 = reinterpret_cast<double2*>(dst + 8 * (long long)n);
#pragma unroll
                for (int pt = 0; pt < 2; pt++) {
                    const double sgp = pt == 0 ? G : -G;
                    double xsi, eta, u[4];
                    if (ln.side == SIDE_LEFT) { xsi = -1.0; eta = sgp; }
                    else if (ln.side == SIDE_RIGHT) { xsi = 1.0; eta = sgp; }
                    else if (ln.side == SIDE_BOTTOM) { xsi = sgp; eta = -1.0; }
                    else { xsi = sgp; eta = 1.0; }
                    dg_u_val(g, cell, xsi, eta, u);
                    d2[2 * pt] = make_double2(u[0], u[1]);
                    d2[2 * pt + 1] = make_double2(u[2], u[3]);
                }
                continue;
            }
            // one 32-byte AoS record per edge cell (usend[k + NVAR*j], SpatialDiscretization.cpp:241)
            double2* d2 = reinterpret_cast<double2*>(dst + 4 * (long long)n);
            d2[0] = make_double2(S[o], S[b.plane + o]);
            d2[1] = make_double2(S[2 * b.plane + o], S[3 * b.plane + o]);
        }
    }
    if (put.peer_mode) {
        // all stores of this CTA are ordered before its ticket; the last CTA raises this rank's flag in
        // every peer's arena (release at system scope)
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned total = gridDim.x * gridDim.y;
            if (atomicAdd(put.ticket, 1u) == total - 1) {
                *put.ticket = 0u;
                __threadfence_system();
                for (int r = 0; r < put.nranks; r++)
                    if (r != put.rank) st_release_sys(put.flag_dst[r], put.seq);
            }
        }
    }
}

// ------------------------------------------------------------------ stop test across ranks, end of a batch
// The launches of a batch test iterations up to the last but two (march.cuh, stop_test); the last two are tested
// here, in order, once every rank's bookkeeping of them has been published.
template <bool FAITHFUL>
__global__ void stop_tail_kernel(StopArgs sx, int* stop, unsigned* flags, long long it_first, long long it_last) {
    for (long long it = it_first; it <= it_last; it++) stop_test(sx, it, stop, flags, 0x7fffffff);
}

// ------------------------------------------------------------------ point-wise test kernels
template <bool FAITHFUL, bool TPG, bool AXI>
__global__ void test_ldfss_kernel(Gas g, long long n, const double* face, const double* uL, const double* uR, double* flux, double* parr) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned fl = 0;
    const double nx = face[4 * i], ny = face[4 * i + 1], len = face[4 * i + 2], yf = face[4 * i + 3];
    FaceGeom fg;
    fg.nx = nx; fg.ny = ny; fg.len = len; fg.yf = yf;
    // a face with this unit normal and length: d = len*(-ny, nx)
    fg.dx = -ny * len; fg.dy = nx * len; fg.rl = 0.0;
    const FaceState L = face_state<FAITHFUL, TPG>(g, uL[4 * i], uL[4 * i + 1], uL[4 * i + 2], fmax(uL[4 * i + 3], 201.0), fl);
    const FaceState R = face_state<FAITHFUL, TPG>(g, uR[4 * i], uR[4 * i + 1], uR[4 * i + 2], fmax(uR[4 * i + 3], 201.0), fl);
    const Flux f = ldfss<FAITHFUL, AXI>(fg, L, R);
    flux[4 * i] = f.f0; flux[4 * i + 1] = f.f1; flux[4 * i + 2] = f.f2; flux[4 * i + 3] = f.f3;
    parr[i] = f.parr;
}

template <bool FAITHFUL>
__global__ void test_boundary_kernel(Gas g, long long n, const int* btype, const double* norm, const double* uInt, double* uGhost) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ui[4] = {uInt[4 * i], uInt[4 * i + 1], uInt[4 * i + 2], fmax(uInt[4 * i + 3], 201.0)};
    const double v2 = ui[1] * ui[1] + ui[2] * ui[2];
    const double p = fmax(ui[0] * g.Rs * ui[3], 1.0);
    const double aint = sqrt(g.gamRs * ui[3]);
    double gh[4] = {uGhost[4 * i], uGhost[4 * i + 1], uGhost[4 * i + 2], uGhost[4 * i + 3]};
    boundary_state(g, btype[i], norm[2 * i], norm[2 * i + 1], ui, aint, p, v2, gh);
    uGhost[4 * i] = gh[0]; uGhost[4 * i + 1] = gh[1]; uGhost[4 * i + 2] = gh[2]; uGhost[4 * i + 3] = gh[3];
}

template <bool FAITHFUL, bool TPG>
__global__ void test_update_state_kernel(Gas g, long long n, double* unk, double* derived) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned fl = 0;
    const double T = fmax(unk[4 * i + 3], 201.0);
    unk[4 * i + 3] = T;
    const double rho = unk[4 * i], u = unk[4 * i + 1], v = unk[4 * i + 2];
    const FaceState s = face_state<FAITHFUL, TPG>(g, rho, u, v, T, fl);
    const double v2 = u * u + v * v;
    const double h = enthalpy<TPG>(g, T, fl);
    const double Cp = heat_capacity<TPG>(g, T, fl);
    double* o = derived + 8 * i;
    o[0] = s.p; o[1] = FAITHFUL ? s.a : 2.0 * s.a;   // the lean FaceState carries a/2
    o[2] = h; o[3] = s.h0; o[4] = h - g.Rs * T + 0.5 * v2; o[5] = v2; o[6] = Cp; o[7] = Cp - g.Rs;
}

template <bool FAITHFUL, bool TPG, bool AXI>
__global__ void test_solve_kernel(Gas g, long long n, const double* dt_ycc, const double* unk, const double* res, double* dv) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned fl = 0;
    const double rho = unk[4 * i], u = unk[4 * i + 1], v = unk[4 * i + 2], T = fmax(unk[4 * i + 3], 201.0);
    const double v2 = u * u + v * v;
    const double h = enthalpy<TPG>(g, T, fl);
    const double Cv = heat_capacity<TPG>(g, T, fl) - g.Rs;
    const double e0 = h - g.Rs * T + 0.5 * v2;
    double inv_cv = g.inv_cv_cpg;
    if constexpr (TPG && !FAITHFUL) inv_cv = fast_rcp(Cv);
    const double r[4] = {res[4 * i], res[4 * i + 1], res[4 * i + 2], res[4 * i + 3]};
    double d[4];
    solve_cell<FAITHFUL, AXI>(dt_ycc[2 * i], dt_ycc[2 * i + 1], rho, u, v, e0, Cv, inv_cv, r, d, fl);
    dv[4 * i] = d[0]; dv[4 * i + 1] = d[1]; dv[4 * i + 2] = d[2]; dv[4 * i + 3] = d[3];
}

// ------------------------------------------------------------------ launch table of one arithmetic mode
struct KernelTable {
    void (*preload)(bool tpg, bool axi);
    void (*ghost)(cudaStream_t, const StepArgs&, const Gas&, int write_dt, const FinalizeArgs* pending);
    void (*pack)(cudaStream_t, const StepArgs&, const Gas&, const SendLine*, int nlines, int max_n, const PutArgs&);
    void (*march)(cudaStream_t, const StepArgs&, const Gas&, bool tpg, bool axi);
    void (*march_push)(cudaStream_t, const StepArgs&, const Gas&, const PutArgs&, const FinalizeArgs&, bool tpg, bool axi);
    int (*coop_capacity)(bool tpg, bool axi);   // CTAs of the persistent kernel that are resident at once on the current device
    cudaError_t (*march_coop)(cudaStream_t, const StepArgs&, const Gas&, const FinalizeArgs&, double* dt_used_alt, double* partials_alt, const DeferArgs&, bool tpg, bool axi);
    void (*stop_tail)(cudaStream_t, const StopArgs&, int* stop, unsigned* flags, long long it_first, long long it_last);
    void (*dg_ghost)(cudaStream_t, const StepArgs&, const Gas&, const DgArgs&, int write_dt);
    void (*dg_cell)(cudaStream_t, const StepArgs&, const Gas&, const DgArgs&, bool tpg, bool axi);
    void (*dg_limit)(cudaStream_t, const StepArgs&, const DgArgs&);
    void (*test_ldfss)(cudaStream_t, const Gas&, bool tpg, bool axi, long long n, const double*, const double*, const double*, double*, double*);
    void (*test_boundary)(cudaStream_t, const Gas&, long long n, const int*, const double*, const double*, double*);
    void (*test_update_state)(cudaStream_t, const Gas&, bool tpg, long long n, double*, double*);
    void (*test_solve)(cudaStream_t, const Gas&, bool tpg, bool axi, long long n, const double*, const double*, const double*, double*);
};

// the production step of PAR2D_MODE_FAST lives in a module of its own (kernels_march.cu)
void launch_march_tuned(cudaStream_t s, const StepArgs& a, const Gas& g, bool tpg, bool axi);
void preload_march_tuned();

template <bool FAITHFUL>
struct Launch {
    template <bool HOOKS>
    static void march_h(cudaStream_t s, const StepArgs& a, const Gas& g, bool tpg, bool axi) {
        const dim3 grid = a.tile_list ? dim3(a.tile_count, 1) : dim3(a.max_tiles, a.nowned);   // one warp strip per CTA
        if (a.tile_list && a.tile_count == 0) return;
        if (tpg) { if (axi) march_kernel<FAITHFUL, true, true, HOOKS><<<grid, MARCH_THREADS, 0, s>>>(a, g); else march_kernel<FAITHFUL, true, false, HOOKS><<<grid, MARCH_THREADS, 0, s>>>(a, g); }
        else { if (axi) march_kernel<FAITHFUL, false, true, HOOKS><<<grid, MARCH_THREADS, 0, s>>>(a, g); else march_kernel<FAITHFUL, false, false, HOOKS><<<grid, MARCH_THREADS, 0, s>>>(a, g); }
    }
    // the production step (advancing, no residual dump) runs the instantiation without the test hooks
    static void march(cudaStream_t s, const StepArgs& a, const Gas& g, bool tpg, bool axi) {
        if (a.update && a.res_out == nullptr) {
            if constexpr (FAITHFUL) march_h<false>(s, a, g, tpg, axi);
            else launch_march_tuned(s, a, g, tpg, axi);
        } else {
            march_h<true>(s, a, g, tpg, axi);
        }
    }
    // the fused step: bookkeeping of the previous step + all strips, perimeter strips pushing ghosts / edge lines / dt
    static void march_push(cudaStream_t s, const StepArgs& a, const Gas& g, const PutArgs& put, const FinalizeArgs& fin, bool tpg, bool axi) {
        // across ranks: tile-list order; single rank: (strips of the largest block, [bookkeeping row +] owned blocks)
        const dim3 grid = a.tile_list ? dim3(a.nfin + a.tile_count, 1) : dim3(a.max_tiles > a.nfin ? a.max_tiles : a.nfin, a.nowned + (a.nfin ? 1 : 0));
        if (tpg) { if (axi) march_fused_kernel<FAITHFUL, true, true><<<grid, MARCH_THREADS, 0, s>>>(a, g, put, fin); else march_fused_kernel<FAITHFUL, true, false><<<grid, MARCH_THREADS, 0, s>>>(a, g, put, fin); }
        else { if (axi) march_fused_kernel<FAITHFUL, false, true><<<grid, MARCH_THREADS, 0, s>>>(a, g, put, fin); else march_fused_kernel<FAITHFUL, false, false><<<grid, MARCH_THREADS, 0, s>>>(a, g, put, fin); }
    }
    // Loads every kernel a ctx of this gas model can launch.  With CUDA's lazy module loading the FIRST launch of a
    // kernel may need the context to go idle - a deadlock if a kernel of another ctx of the same process is spinning
    // on a flag that only the kernel being loaded would raise (par2d_peer_connect_local).  par2d_create calls this
    // while nothing is running.
    static void preload(bool tpg, bool axi) {
        cudaFuncAttributes fa;
        auto touch = [&](const void* f) { cudaFuncGetAttributes(&fa, f); };
        if constexpr (!FAITHFUL) preload_march_tuned();
        touch((const void*)ghost_kernel<FAITHFUL>);
        touch((const void*)pack_kernel<FAITHFUL>);
        touch((const void*)stop_tail_kernel<FAITHFUL>);
        touch((const void*)dg_ghost_kernel<FAITHFUL>);
        touch((const void*)dg_limit_kernel<FAITHFUL>);
        touch(coop_fn(tpg, axi));
        if (tpg) {
            if (axi) { touch((const void*)march_kernel<FAITHFUL, true, true, false>); touch((const void*)march_kernel<FAITHFUL, true, true, true>); touch((const void*)march_fused_kernel<FAITHFUL, true, true>); touch((const void*)dg_cell_kernel<FAITHFUL, true, true>); }
            else { touch((const void*)march_kernel<FAITHFUL, true, false, false>); touch((const void*)march_kernel<FAITHFUL, true, false, true>); touch((const void*)march_fused_kernel<FAITHFUL, true, false>); touch((const void*)dg_cell_kernel<FAITHFUL, true, false>); }
        } else {
            if (axi) { touch((const void*)march_kernel<FAITHFUL, false, true, false>); touch((const void*)march_kernel<FAITHFUL, false, true, true>); touch((const void*)march_fused_kernel<FAITHFUL, false, true>); touch((const void*)dg_cell_kernel<FAITHFUL, false, true>); }
            else { touch((const void*)march_kernel<FAITHFUL, false, false, false>); touch((const void*)march_kernel<FAITHFUL, false, false, true>); touch((const void*)march_fused_kernel<FAITHFUL, false, false>); touch((const void*)dg_cell_kernel<FAITHFUL, false, false>); }
        }
        cudaGetLastError();
    }
    static const void* coop_fn(bool tpg, bool axi) {
        if (tpg) return axi ? (const void*)march_persist_kernel<FAITHFUL, true, true> : (const void*)march_persist_kernel<FAITHFUL, true, false>;
        return axi ? (const void*)march_persist_kernel<FAITHFUL, false, true> : (const void*)march_persist_kernel<FAITHFUL, false, false>;
    }
    static int coop_capacity(bool tpg, bool axi) {
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_fn(tpg, axi), MARCH_THREADS, 0) != cudaSuccess) return 0;
        return per_sm * sms;
    }
    static cudaError_t march_coop(cudaStream_t s, const StepArgs& a, const Gas& g, const FinalizeArgs& fin, double* dt_used_alt, double* partials_alt, const DeferArgs& df,
                                  bool tpg, bool axi) {
        void* args[] = {(void*)&a, (void*)&g, (void*)&fin, (void*)&dt_used_alt, (void*)&partials_alt, (void*)&df};
        return cudaLaunchCooperativeKernel(coop_fn(tpg, axi), dim3((df.on ? 0 : a.nowned) + a.tile_count), dim3(MARCH_THREADS), args, 0, s);
    }
    static void stop_tail(cudaStream_t s, const StopArgs& sx, int* stop, unsigned* flags, long long it_first, long long it_last) {
        stop_tail_kernel<FAITHFUL><<<1, 1, 0, s>>>(sx, stop, flags, it_first, it_last);
    }
    static void ghost(cudaStream_t s, const StepArgs& a, const Gas& g, int write_dt, const FinalizeArgs* pending) {
        FinalizeArgs fin;
        memset(&fin, 0, sizeof fin);
        if (pending) fin = *pending;
        const int nfin = pending ? a.nowned : 0;
        ghost_kernel<FAITHFUL><<<dim3((a.max_faces + 255) / 256, a.nowned + nfin), 256, 0, s>>>(a, g, write_dt, fin, nfin);
    }
    static void pack(cudaStream_t s, const StepArgs& a, const Gas& g, const SendLine* lines, int nlines, int max_n, const PutArgs& put) {
        dim3 grid((unsigned)((max_n + 255) / 256 > 0 ? (max_n + 255) / 256 : 1), (unsigned)(nlines > 0 ? nlines : 1));
        pack_kernel<FAITHFUL><<<grid, 256, 0, s>>>(a, g, lines, nlines, put);
    }
    // second-order branch (dg.cuh)
    static void dg_ghost(cudaStream_t s, const StepArgs& a, const Gas& g, const DgArgs& d, int write_dt) {
        dg_ghost_kernel<FAITHFUL><<<dim3((d.max_points + 255) / 256, a.nowned), 256, 0, s>>>(a, g, d, write_dt);
    }
    static void dg_cell(cudaStream_t s, const StepArgs& a, const Gas& g, const DgArgs& d, bool tpg, bool axi) {
        const dim3 grid((d.max_tiles + DG_WARPS - 1) / DG_WARPS, a.nowned);   // DG_WARPS strips per CTA
        if (tpg) { if (axi) dg_cell_kernel<FAITHFUL, true, true><<<grid, 32 * DG_WARPS, 0, s>>>(a, g, d); else dg_cell_kernel<FAITHFUL, true, false><<<grid, 32 * DG_WARPS, 0, s>>>(a, g, d); }
        else { if (axi) dg_cell_kernel<FAITHFUL, false, true><<<grid, 32 * DG_WARPS, 0, s>>>(a, g, d); else dg_cell_kernel<FAITHFUL, false, false><<<grid, 32 * DG_WARPS, 0, s>>>(a, g, d); }
    }
    static void dg_limit(cudaStream_t s, const StepArgs& a, const DgArgs& d) {
        dg_limit_kernel<FAITHFUL><<<dim3(d.max_lim_tiles, a.nowned), DG_THREADS, 0, s>>>(a, d);
    }
    static void test_ldfss(cudaStream_t s, const Gas& g, bool tpg, bool axi, long long n, const double* face, const double* uL, const double* uR, double* flux, double* parr) {
        const int grid = (int)((n + 127) / 128);
        if (tpg) { if (axi) test_ldfss_kernel<FAITHFUL, true, true><<<grid, 128, 0, s>>>(g, n, face, uL, uR, flux, parr); else test_ldfss_kernel<FAITHFUL, true, false><<<grid, 128, 0, s>>>(g, n, face, uL, uR, flux, parr); }
        else { if (axi) test_ldfss_kernel<FAITHFUL, false, true><<<grid, 128, 0, s>>>(g, n, face, uL, uR, flux, parr); else test_ldfss_kernel<FAITHFUL, false, false><<<grid, 128, 0, s>>>(g, n, face, uL, uR, flux, parr); }
    }
    static void test_boundary(cudaStream_t s, const Gas& g, long long n, const int* bt, const double* norm, const double* ui, double* ug) {
        test_boundary_kernel<FAITHFUL><<<(int)((n + 127) / 128), 128, 0, s>>>(g, n, bt, norm, ui, ug);
    }
    static void test_update_state(cudaStream_t s, const Gas& g, bool tpg, long long n, double* unk, double* der) {
        const int grid = (int)((n + 127) / 128);
        if (tpg) test_update_state_kernel<FAITHFUL, true><<<grid, 128, 0, s>>>(g, n, unk, der);
        else test_update_state_kernel<FAITHFUL, false><<<grid, 128, 0, s>>>(g, n, unk, der);
    }
    static void test_solve(cudaStream_t s, const Gas& g, bool tpg, bool axi, long long n, const double* dy, const double* unk, const double* res, double* dv) {
        const int grid = (int)((n + 127) / 128);
        if (tpg) { if (axi) test_solve_kernel<FAITHFUL, true, true><<<grid, 128, 0, s>>>(g, n, dy, unk, res, dv); else test_solve_kernel<FAITHFUL, true, false><<<grid, 128, 0, s>>>(g, n, dy, unk, res, dv); }
        else { if (axi) test_solve_kernel<FAITHFUL, false, true><<<grid, 128, 0, s>>>(g, n, dy, unk, res, dv); else test_solve_kernel<FAITHFUL, false, false><<<grid, 128, 0, s>>>(g, n, dy, unk, res, dv); }
    }
    static const KernelTable* table() {
        static const KernelTable t = {&preload, &ghost, &pack, &march, &march_push, &coop_capacity, &march_coop, &stop_tail, &dg_ghost, &dg_cell, &dg_limit, &test_ldfss, &test_boundary, &test_update_state, &test_solve};
        return &t;
    }
};

const KernelTable* kernels_fast();
const KernelTable* kernels_faithful();

}  // namespace par2d
