// kernels.cuh -- the kernels of the explicit FV residual + update path, templated on
// arithmetic mode (see physics.cuh).  Included by kernels_fast.cu (default FMA contraction)
// and kernels_faithful.cu (-fmad=false).
//
//   ghost_kernel  : boundary-condition ghost states + neighbour ghost lines + per-block dt
//                   (generate_ghost_cells SpatialDiscretization.cpp:17-99, the receive side of
//                   the halo exchange :263-269, find_dt main.cpp:19-37)
//   step_kernel   : one tile of one block: TMA-staged state+node tile with halo ->
//                   per-cell derived state -> one LDFSS per face (EulerFlux.cpp:121) ->
//                   residual (SpatialDiscretization.cpp:481-1011) -> sum of squares
//                   (main.cpp:39-57) -> Jacobian solve + 1/ycc (main.cpp:307-331) ->
//                   unk += dv and the UpdateState clamps (main.cpp:394-404)
//   test kernels  : the point-wise device functions over arrays (test hooks of the C ABI)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"
#include "physics.cuh"
#include "march.cuh"

namespace par2d {

// ------------------------------------------------------------------ PTX helpers (TMA + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a mis-programmed copy must surface as an error, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 20); spin++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}
// 3-D tiled bulk tensor load global -> shared, completion on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tmap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}

// ------------------------------------------------------------------ small helpers
// (col,row) in the padded arrays of the n-th cell along a side of a block
__device__ __forceinline__ void edge_cell_pos(int ncx, int ncy, int side, int n, int& col, int& row) {
    switch (side) {
        case SIDE_BOTTOM: col = COL0 + n; row = 1; break;
        case SIDE_RIGHT: col = COL0 + ncx - 1; row = n + 1; break;
        case SIDE_TOP: col = COL0 + n; row = ncy; break;
        default: col = COL0; row = n + 1; break;
    }
}
__device__ __forceinline__ void ghost_cell_pos(int ncx, int ncy, int side, int n, int& col, int& row) {
    switch (side) {
        case SIDE_BOTTOM: col = COL0 + n; row = 0; break;
        case SIDE_RIGHT: col = COL0 + ncx; row = n + 1; break;
        case SIDE_TOP: col = COL0 + n; row = ncy + 1; break;
        default: col = COL0 - 1; row = n + 1; break;
    }
}

// per-block time step from cell (0,0): find_dt, main.cpp:19-37
template <bool FAITHFUL>
__device__ __forceinline__ void block_dt(const StepArgs& a, const Gas& g, const BlockDev& b) {
    const double* S = b.state[a.cur];
    const long long o = (long long)b.pitch + COL0;
    const double u = S[b.plane + o], v = S[2 * b.plane + o], T = S[3 * b.plane + o];
    const double v2 = u * u + v * v;
    const double vmax = sqrt(v2) + sqrt(g.gamRs * T);
    const double dt = g.cfl * b.min_face / vmax;
    if (isnan(dt)) atomicOr(a.flags, FLAG_NAN);
    a.dt_table[b.gid] = dt;
}

// ------------------------------------------------------------------ ghost fill
template <bool FAITHFUL>
__global__ void __launch_bounds__(256) ghost_kernel(StepArgs a, Gas g, int write_dt) {
    if (*a.stop) return;
    const BlockDev b = a.blocks[blockIdx.y];   // grid.y = owned blocks, grid.x covers the longest perimeter
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

    if (write_dt && side == SIDE_BOTTOM && n == 0) block_dt<FAITHFUL>(a, g, b);

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
    } else if (b.nb_kind[side] == NB_REMOTE) {
        const double* line = b.recv[side] + 4 * (long long)n;
#pragma unroll
        for (int v = 0; v < 4; v++) gh[v] = line[v];
        gh[3] = fmax(gh[3], 201.0);
    } else {
        // outward normal of the boundary face (SpatialDiscretization.cpp:33-34,51-52,71-72,91-92);
        // the right side reads the normal of column 0, as the reference does
        const double* X = b.nodes;
        const double* Y = b.nodes + b.plane;
        long long p0, p1;
        double sgn;
        int code;
        if (side == SIDE_BOTTOM) { p0 = (long long)b.pitch + COL0 + n; p1 = p0 + 1; sgn = 1.0; code = b.ibound[n]; }
        else if (side == SIDE_RIGHT) { p0 = (long long)(n + 1) * b.pitch + COL0; p1 = p0 + b.pitch; sgn = 1.0; code = b.ibound[b.ncx + n]; }
        else if (side == SIDE_TOP) { p0 = (long long)(b.ncy + 1) * b.pitch + COL0 + n; p1 = p0 + 1; sgn = -1.0; code = b.ibound[2 * b.ncx + b.ncy - 1 - n]; }
        else { p0 = (long long)(n + 1) * b.pitch + COL0; p1 = p0 + b.pitch; sgn = -1.0; code = b.ibound[2 * b.ncx + 2 * b.ncy - 1 - n]; }
        const FaceGeom fg = face_geom<true>(X[p0], Y[p0], X[p1], Y[p1]);   // O(perimeter): unit normal as MeshModule.cpp:52-81
        int c, r;
        edge_cell_pos(b.ncx, b.ncy, side, n, c, r);
        const long long io = (long long)r * b.pitch + c;
        double ui[4];
#pragma unroll
        for (int v = 0; v < 4; v++) ui[v] = S[v * b.plane + io];
        const double v2 = ui[1] * ui[1] + ui[2] * ui[2];
        const double p = fmax(ui[0] * g.Rs * ui[3], 1.0);
        const double aint = sqrt(g.gamRs * ui[3]);
#pragma unroll
        for (int v = 0; v < 4; v++) gh[v] = S[v * b.plane + go];  // stays as is if the BC leaves it unwritten
        boundary_state(g, code, sgn * fg.nx, sgn * fg.ny, ui, aint, p, v2, gh);
    }
#pragma unroll
    for (int v = 0; v < 4; v++) S[v * b.plane + go] = gh[v];
}

// ------------------------------------------------------------------ pack (multi-rank send side)
template <bool FAITHFUL>
__global__ void __launch_bounds__(256) pack_kernel(StepArgs a, Gas g, const SendLine* lines, int nlines) {
    if (*a.stop) return;
    if (blockIdx.x == 0 && threadIdx.x < a.nowned) block_dt<FAITHFUL>(a, g, a.blocks[threadIdx.x]);
    for (int l = blockIdx.y; l < nlines; l += gridDim.y) {
        const SendLine ln = lines[l];
        const BlockDev& b = a.blocks[ln.block];
        const double* S = b.state[a.cur];
        for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < ln.n; n += gridDim.x * blockDim.x) {
            int c, r;
            edge_cell_pos(b.ncx, b.ncy, ln.side, n, c, r);
            const long long o = (long long)r * b.pitch + c;
#pragma unroll
            for (int v = 0; v < 4; v++) ln.dst[4 * (long long)n + v] = S[v * b.plane + o];
        }
    }
}

// ------------------------------------------------------------------ the step kernel
struct StepSmem {
    // TMA destinations.  The box starts one column left of the halo (HOFF) because a tiled copy of
    // 8-byte elements must start on a 16-byte boundary (odd start columns raise an illegal-instruction
    // fault on sm_100a, see tools/tma_probe): halo column hx lives at [hx + HOFF].
    double U[4][HY][HXL];  // rho,u,v,T with halo
    double N[2][HY][HXL];  // node x,y
    double P[HY][HX], A[HY][HX], H[HY][HX];  // p, a, h0 per halo cell
    double F[5][TY][TX + 1];   // i-face fluxes + parr
    double G[5][TY + 1][TX];   // j-face fluxes + parr
    double red[STEP_THREADS / 32][4];
    double dt;
    uint64_t bar;
};

template <bool FAITHFUL, bool TPG, bool AXI, bool USE_TMA>
__global__ void __launch_bounds__(STEP_THREADS, 2) step_kernel(const StepArgs a, const Gas g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StepSmem& sm = *reinterpret_cast<StepSmem*>(smem_raw);
    if (*a.stop) return;

    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int k = blockIdx.y;                  // grid.y = owned blocks, grid.x = tiles of the largest block
    const BlockDev& b = a.blocks[k];
    const int lt = blockIdx.x;
    if (lt >= b.tiles_x * b.tiles_y) return;
    const int tile = b.tile0 + lt;
    const int i0 = (lt % b.tiles_x) * TX, j0 = (lt / b.tiles_x) * TY;
    const int ncx = b.ncx, ncy = b.ncy, pitch = b.pitch;
    const long long plane = b.plane;
    const double* __restrict__ Sin = b.state[a.cur];
    double* __restrict__ Sout = b.state[a.cur ^ 1];

    // ---- stage the haloed tile: box origin = array column i0+COL0-1, array row j0
    if constexpr (USE_TMA) {
        if (tid == 0) mbar_init(&sm.bar, 1);
        __syncthreads();
        if (tid == 0) {
            const char* maps = static_cast<const char*>(a.tmaps) + (size_t)k * 3 * 128;
            // descriptors live in global memory: make them visible to the tensor-map proxy
            tmap_acquire(maps + a.cur * 128);
            tmap_acquire(maps + 2 * 128);
            mbar_expect_tx(&sm.bar, (uint32_t)(sizeof(sm.U) + sizeof(sm.N)));
            tma_load_3d(&sm.U[0][0][0], maps + a.cur * 128, i0 + COL0 - 1 - HOFF, j0, 0, &sm.bar);
            tma_load_3d(&sm.N[0][0][0], maps + 2 * 128, i0 + COL0 - 1 - HOFF, j0, 0, &sm.bar);
        }
    } else {
        const int rows = ncy + 2;
        for (int idx = tid; idx < 6 * HY * HXL; idx += STEP_THREADS) {
            const int pl = idx / (HY * HXL), rem = idx % (HY * HXL), hy = rem / HXL, hx = rem % HXL;
            const int col = i0 + COL0 - 1 - HOFF + hx, row = j0 + hy;
            const bool in = col < pitch && row < rows;
            double val = 0.0;
            if (in) val = pl < 4 ? Sin[pl * plane + (long long)row * pitch + col] : b.nodes[(pl - 4) * plane + (long long)row * pitch + col];
            if (pl < 4) sm.U[pl][hy][hx] = val; else sm.N[pl - 4][hy][hx] = val;
        }
    }
    // global dt = max over blocks (main.cpp:270-292), overlapped with the copy
    if (tid < 32) {
        double m = 0.0;
        for (int q = tid; q < a.nblocks_global; q += 32) m = fmax(m, a.dt_table[q]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
        if (tid == 0) sm.dt = m;
    }
    if constexpr (USE_TMA) {
        if (!mbar_wait(&sm.bar, 0)) { atomicOr(a.flags, FLAG_NAN | 0x80000000u); return; }
    } else {
        __syncthreads();
    }

    unsigned flags = 0;
    // ---- phase A: derived state of every staged cell (own + halo)
    for (int idx = tid; idx < HY * HX; idx += STEP_THREADS) {
        const int hy = idx / HX, hx = idx % HX;
        const int gi = i0 - 1 + hx, gj = j0 - 1 + hy;  // cell index, -1 and ncx/ncy are the ghost ring
        const bool xin = gi >= 0 && gi < ncx, yin = gj >= 0 && gj < ncy;
        const bool real = (xin && gj >= -1 && gj <= ncy) || (yin && gi >= -1 && gi <= ncx);
        unsigned fl = 0;
        const FaceState s = face_state<FAITHFUL, TPG>(g, sm.U[0][hy][hx + HOFF], sm.U[1][hy][hx + HOFF], sm.U[2][hy][hx + HOFF], sm.U[3][hy][hx + HOFF], fl);
        sm.P[hy][hx] = s.p; sm.A[hy][hx] = s.a; sm.H[hy][hx] = s.h0;
        if (real) flags |= fl;
    }
    __syncthreads();

    auto load_state = [&](int hx, int hy) {
        FaceState s;
        s.rho = sm.U[0][hy][hx + HOFF]; s.u = sm.U[1][hy][hx + HOFF]; s.v = sm.U[2][hy][hx + HOFF];
        s.p = sm.P[hy][hx]; s.a = sm.A[hy][hx]; s.h0 = sm.H[hy][hx];
        return s;
    };

    // ---- phase B: one LDFSS per face
    const int hx = tx + 1, hy = ty + 1;
    const FaceState own = load_state(hx, hy);
    {   // face between cell (i-1,j) [left state] and (i,j): node (i,j) -> (i,j+1)
        const FaceGeom fg = face_geom<FAITHFUL>(sm.N[0][hy][hx + HOFF], sm.N[1][hy][hx + HOFF], sm.N[0][hy + 1][hx + HOFF], sm.N[1][hy + 1][hx + HOFF]);
        const Flux f = ldfss<FAITHFUL, AXI>(fg, load_state(hx - 1, hy), own);
        sm.F[0][ty][tx] = f.f0; sm.F[1][ty][tx] = f.f1; sm.F[2][ty][tx] = f.f2; sm.F[3][ty][tx] = f.f3; sm.F[4][ty][tx] = f.parr;
    }
    {   // face between cell (i,j) [left state = upper cell] and (i,j-1): node (i,j) -> (i+1,j)
        const FaceGeom fg = face_geom<FAITHFUL>(sm.N[0][hy][hx + HOFF], sm.N[1][hy][hx + HOFF], sm.N[0][hy][hx + 1 + HOFF], sm.N[1][hy][hx + 1 + HOFF]);
        const Flux f = ldfss<FAITHFUL, AXI>(fg, own, load_state(hx, hy - 1));
        sm.G[0][ty][tx] = f.f0; sm.G[1][ty][tx] = f.f1; sm.G[2][ty][tx] = f.f2; sm.G[3][ty][tx] = f.f3; sm.G[4][ty][tx] = f.parr;
    }
    if (tid < TY) {   // closing column of i-faces (warp 0)
        const int r = tid + 1;
        const FaceGeom fg = face_geom<FAITHFUL>(sm.N[0][r][TX + 1 + HOFF], sm.N[1][r][TX + 1 + HOFF], sm.N[0][r + 1][TX + 1 + HOFF], sm.N[1][r + 1][TX + 1 + HOFF]);
        const Flux f = ldfss<FAITHFUL, AXI>(fg, load_state(TX, r), load_state(TX + 1, r));
        sm.F[0][tid][TX] = f.f0; sm.F[1][tid][TX] = f.f1; sm.F[2][tid][TX] = f.f2; sm.F[3][tid][TX] = f.f3; sm.F[4][tid][TX] = f.parr;
    } else if (tid >= 32 && tid < 32 + TX) {   // closing row of j-faces (warp 1)
        const int c = tid - 32 + 1;
        const FaceGeom fg = face_geom<FAITHFUL>(sm.N[0][TY + 1][c + HOFF], sm.N[1][TY + 1][c + HOFF], sm.N[0][TY + 1][c + 1 + HOFF], sm.N[1][TY + 1][c + 1 + HOFF]);
        const Flux f = ldfss<FAITHFUL, AXI>(fg, load_state(c, TY + 1), load_state(c, TY));
        sm.G[0][TY][c - 1] = f.f0; sm.G[1][TY][c - 1] = f.f1; sm.G[2][TY][c - 1] = f.f2; sm.G[3][TY][c - 1] = f.f3; sm.G[4][TY][c - 1] = f.parr;
    }
    __syncthreads();

    // ---- phase C: residual, norms, solve, update
    const int i = i0 + tx, j = j0 + ty;
    const bool valid = i < ncx && j < ncy;
    double sq[4] = {0.0, 0.0, 0.0, 0.0};
    if (valid) {
        const double x1 = sm.N[0][hy][hx + HOFF], x2 = sm.N[0][hy][hx + 1 + HOFF], x3 = sm.N[0][hy + 1][hx + 1 + HOFF], x4 = sm.N[0][hy + 1][hx + HOFF];
        const double y1 = sm.N[1][hy][hx + HOFF], y2 = sm.N[1][hy][hx + 1 + HOFF], y3 = sm.N[1][hy + 1][hx + 1 + HOFF], y4 = sm.N[1][hy + 1][hx + HOFF];
        const double vol = cell_volume(x1, y1, x2, y2, x3, y3, x4, y4);
        const double ycc = 0.25 * (y1 + y2 + y3 + y4);
        const double yc = AXI ? ycc : 0.0;
        double r[4];
        if constexpr (FAITHFUL) {
            // accumulation order of SpatialDiscretization.cpp:481-625 (interior faces: left, right,
            // bottom, top) then :643-977 (boundary faces: left, right, bottom, top)
            const bool bl = i == 0, br = i == ncx - 1, bb = j == 0, bt = j == ncy - 1;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                double fl = sm.F[v][ty][tx], fr = sm.F[v][ty][tx + 1], gb = sm.G[v][ty][tx], gt = sm.G[v][ty + 1][tx];
                if (v == 2) {
                    fl = (fl + (yc * sm.F[4][ty][tx])); fr = (fr + (yc * sm.F[4][ty][tx + 1]));
                    gb = (gb + (yc * sm.G[4][ty][tx])); gt = (gt + (yc * sm.G[4][ty + 1][tx]));
                }
                double acc = 0.0;
                if (!bl) acc += fl;
                if (!br) acc -= fr;
                if (!bb) acc -= gb;
                if (!bt) acc += gt;
                if (bl) acc += fl;
                if (br) acc -= fr;
                if (bb) acc -= gb;
                if (bt) acc += gt;
                r[v] = acc / vol;
            }
        } else {
            const double ivol = fast_rcp(vol);
#pragma unroll
            for (int v = 0; v < 4; v++) r[v] = (sm.F[v][ty][tx] - sm.F[v][ty][tx + 1]) + (sm.G[v][ty + 1][tx] - sm.G[v][ty][tx]);
            if constexpr (AXI) r[2] = fma(ycc, (sm.F[4][ty][tx] - sm.F[4][ty][tx + 1]) + (sm.G[4][ty + 1][tx] - sm.G[4][ty][tx]), r[2]);
#pragma unroll
            for (int v = 0; v < 4; v++) r[v] *= ivol;
        }
        if (isnan(r[0]) || isnan(r[1]) || isnan(r[2]) || isnan(r[3])) flags |= FLAG_NAN;
#pragma unroll
        for (int v = 0; v < 4; v++) sq[v] = r[v] * r[v];

        if (a.res_out != nullptr && a.res_block == k) {
            double* o = a.res_out + 4 * ((long long)j * ncx + i);
            o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
        }
        if (a.update) {
            const double T = sm.U[3][hy][hx + HOFF];
            const double v2 = own.u * own.u + own.v * own.v;
            unsigned fl2 = 0;
            const double h = enthalpy<TPG>(g, T, fl2);
            const double Cv = heat_capacity<TPG>(g, T, fl2) - g.Rs;
            const double e0 = h - g.Rs * T + 0.5 * v2;
            double inv_cv = g.inv_cv_cpg;
            if constexpr (TPG && !FAITHFUL) inv_cv = fast_rcp(Cv);
            double dv[4];
            solve_cell<FAITHFUL, AXI>(sm.dt, ycc, own.rho, own.u, own.v, e0, Cv, inv_cv, r, dv, flags);
            const double nrho = own.rho + dv[0], nu = own.u + dv[1], nv = own.v + dv[2];
            double nT = T + dv[3];
            if (!(nrho > 0.0)) flags |= FLAG_DENSITY;
            nT = fmax(nT, 201.0);
            const long long o = (long long)(j + 1) * pitch + COL0 + i;
            Sout[o] = nrho; Sout[plane + o] = nu; Sout[2 * plane + o] = nv; Sout[3 * plane + o] = nT;
        }
    }
    // ---- sum of squares of this tile, fixed order: lanes (xor tree), then warps 0..7
#pragma unroll
    for (int v = 0; v < 4; v++) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sq[v] += __shfl_xor_sync(0xffffffffu, sq[v], s);
    }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int v = 0; v < 4; v++) sm.red[tid >> 5][v] = sq[v];
    }
    if (flags) atomicOr(a.flags, flags);
    __syncthreads();
    if (tid < 4) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < STEP_THREADS / 32; w++) s += sm.red[w][tid];
        a.partials[4 * (long long)tile + tid] = s;
    }
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
    o[0] = s.p; o[1] = s.a; o[2] = h; o[3] = s.h0; o[4] = h - g.Rs * T + 0.5 * v2; o[5] = v2; o[6] = Cp; o[7] = Cp - g.Rs;
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
    size_t step_smem;
    cudaError_t (*prepare)();  // opt in to the dynamic shared memory size
    void (*ghost)(cudaStream_t, const StepArgs&, const Gas&, int write_dt);
    void (*pack)(cudaStream_t, const StepArgs&, const Gas&, const SendLine*, int nlines, int max_n);
    void (*step)(cudaStream_t, const StepArgs&, const Gas&, bool tpg, bool axi, bool use_tma);
    void (*march)(cudaStream_t, const StepArgs&, const Gas&, bool tpg, bool axi);
    void (*test_ldfss)(cudaStream_t, const Gas&, bool tpg, bool axi, long long n, const double*, const double*, const double*, double*, double*);
    void (*test_boundary)(cudaStream_t, const Gas&, long long n, const int*, const double*, const double*, double*);
    void (*test_update_state)(cudaStream_t, const Gas&, bool tpg, long long n, double*, double*);
    void (*test_solve)(cudaStream_t, const Gas&, bool tpg, bool axi, long long n, const double*, const double*, const double*, double*);
};

template <bool FAITHFUL>
struct Launch {
    template <bool TPG, bool AXI, bool TMA>
    static void step_one(cudaStream_t s, const StepArgs& a, const Gas& g) {
        step_kernel<FAITHFUL, TPG, AXI, TMA><<<dim3(a.max_tiles, a.nowned), STEP_THREADS, sizeof(StepSmem), s>>>(a, g);
    }
    template <bool TPG, bool AXI>
    static void step_tma(cudaStream_t s, const StepArgs& a, const Gas& g, bool tma) {
        if (tma) step_one<TPG, AXI, true>(s, a, g); else step_one<TPG, AXI, false>(s, a, g);
    }
    static void step(cudaStream_t s, const StepArgs& a, const Gas& g, bool tpg, bool axi, bool tma) {
        if (tpg) { if (axi) step_tma<true, true>(s, a, g, tma); else step_tma<true, false>(s, a, g, tma); }
        else { if (axi) step_tma<false, true>(s, a, g, tma); else step_tma<false, false>(s, a, g, tma); }
    }
    template <bool TPG, bool AXI>
    static cudaError_t prep_one() {
        cudaError_t e = cudaFuncSetAttribute(step_kernel<FAITHFUL, TPG, AXI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem));
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(step_kernel<FAITHFUL, TPG, AXI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem));
    }
    static cudaError_t prepare() {
        cudaError_t e;
        if ((e = prep_one<false, false>()) != cudaSuccess) return e;
        if ((e = prep_one<false, true>()) != cudaSuccess) return e;
        if ((e = prep_one<true, false>()) != cudaSuccess) return e;
        return prep_one<true, true>();
    }
    static void march(cudaStream_t s, const StepArgs& a, const Gas& g, bool tpg, bool axi) {
        const dim3 grid((a.max_tiles + MARCH_WARPS - 1) / MARCH_WARPS, a.nowned);
        if (tpg) { if (axi) march_kernel<FAITHFUL, true, true><<<grid, MARCH_THREADS, 0, s>>>(a, g); else march_kernel<FAITHFUL, true, false><<<grid, MARCH_THREADS, 0, s>>>(a, g); }
        else { if (axi) march_kernel<FAITHFUL, false, true><<<grid, MARCH_THREADS, 0, s>>>(a, g); else march_kernel<FAITHFUL, false, false><<<grid, MARCH_THREADS, 0, s>>>(a, g); }
    }
    static void ghost(cudaStream_t s, const StepArgs& a, const Gas& g, int write_dt) {
        ghost_kernel<FAITHFUL><<<dim3((a.max_faces + 255) / 256, a.nowned), 256, 0, s>>>(a, g, write_dt);
    }
    static void pack(cudaStream_t s, const StepArgs& a, const Gas& g, const SendLine* lines, int nlines, int max_n) {
        dim3 grid((unsigned)((max_n + 255) / 256 > 0 ? (max_n + 255) / 256 : 1), (unsigned)(nlines > 0 ? nlines : 1));
        pack_kernel<FAITHFUL><<<grid, 256, 0, s>>>(a, g, lines, nlines);
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
        static const KernelTable t = {sizeof(StepSmem), &prepare, &ghost, &pack, &step, &march, &test_ldfss, &test_boundary, &test_update_state, &test_solve};
        return &t;
    }
};

const KernelTable* kernels_fast();
const KernelTable* kernels_faithful();

}  // namespace par2d
