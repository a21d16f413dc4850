// kernels_common.cuh -- launchers of the mode-independent kernels (kernels_common.cu).
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
namespace par2d {

// Residual bookkeeping of one step (main.cpp:39-57,365-392,494).  Done by the first WARP of one CTA per owned
// block k (the stand-alone finalize_kernel, the bookkeeping rows of the ghost kernel, the bookkeeping CTAs of the
// fused / persistent step kernels): sums the block's strip partials in a FIXED order - lane l adds strips l, l+32,
// l+64, ... in that order, then a xor-shuffle tree over the 32 lane sums - and takes the square roots
// (main.cpp:53-56); the last CTA to finish adds the blocks in block order and appends dt, sum ||res||, sum res0 and
// relres to the device history.  No multiply-add is involved, so the result does not depend on the contraction
// mode of the including TU; one warp does all of it in ~2 us for the 100-strip blocks of the shipped cases, which
// is what lets it ride inside a persistent launch without holding up the steps.
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long fin_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fin_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// FIN_CHUNK strips are staged per pass when the calling CTA is larger than a warp (finalize_kernel, the bookkeeping
// rows of the ghost kernel): every thread fetches strips into shared memory, the first warp then adds them in the
// canonical order - the same bits as the warp-only form, but with the whole CTA's loads in flight (in situ the
// partial sums come from HBM: a single warp walking 2208 strips of a 2048^2 block took 33 us, measured).
constexpr int FIN_CHUNK = 1024;


// stage: FIN_CHUNK*4 doubles of shared memory, or null when the caller is a single warp (then blockDim.x must be 32).
// With staging ALL threads of the CTA must call; only the first warp returns with work done.
// First part: the four residual norms of block k (main.cpp:39-57) from its strips' sums, in norm[] of every lane of the
// first warp.  Returns false for the threads of a staged CTA that are done (everything but the first warp).
__device__ __forceinline__ bool finalize_block_norms(const FinalizeArgs& a, int k, double* stage, double norm[4]) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const BlockDev& b = a.blocks[k];
    norm[0] = norm[1] = norm[2] = norm[3] = 0.0;
    // first order: one group of 4 sums; second-order branch: three groups (dudt, duxdt, duydt) whose square roots
    // are added as main.cpp:360-364 does, ressum + (ressumx + ressumy)
    const int ngroups = a.dg_partials ? 3 : 1;
    double parked[4] = {0.0, 0.0, 0.0, 0.0};   // ressumx while ressumy is being reduced
    for (int grp = 0; grp < ngroups; grp++) {
        const int ntiles = a.dg_partials ? b.dg_ntiles : b.tiles_x * b.tiles_y;
        const double* src = a.dg_partials ? a.dg_partials + 12 * (long long)b.dg_tile0 + 4 * grp : a.partials + 4 * (long long)b.tile0;
        const int stride = a.dg_partials ? 12 : 4;
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        if (stage != nullptr) {
            for (int base = 0; base < ntiles; base += FIN_CHUNK) {
                const int n = min(FIN_CHUNK, ntiles - base);
                __syncthreads();
                for (int t = threadIdx.x; t < n; t += blockDim.x) {
                    const double2* p = reinterpret_cast<const double2*>(src + (long long)stride * (base + t));
                    const double2 p0 = __ldcg(p), p1 = __ldcg(p + 1);
                    reinterpret_cast<double2*>(stage)[2 * t] = p0;
                    reinterpret_cast<double2*>(stage)[2 * t + 1] = p1;
                }
                __syncthreads();
                if (threadIdx.x < 32) {
                    for (int t = lane; t < n; t += 32) {
                        const double2 p0 = reinterpret_cast<const double2*>(stage)[2 * t], p1 = reinterpret_cast<const double2*>(stage)[2 * t + 1];
                        s[0] += p0.x; s[1] += p0.y; s[2] += p1.x; s[3] += p1.y;
                    }
                }
            }
            if (threadIdx.x >= 32 && grp == ngroups - 1) return false;
            if (threadIdx.x >= 32) continue;
        } else {
            // eight strips per lane requested together, then added in strip order: the trip count is the same for every
            // lane and no pass waits for more than one round trip to L2 (a plain `for (t = lane; ...)` is compiled into a
            // serial remainder loop of up to seven round trips ahead of the unrolled one)
#pragma unroll 1
            for (int base = 0; base < ntiles; base += 32 * 8) {
                double2 p0[8], p1[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int t = base + 32 * u + lane;
                    p0[u] = p1[u] = make_double2(0.0, 0.0);
                    if (t < ntiles) {
                        const double2* p = reinterpret_cast<const double2*>(src + (long long)stride * t);
                        p0[u] = __ldcg(p); p1[u] = __ldcg(p + 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (base + 32 * u + lane < ntiles) { s[0] += p0[u].x; s[1] += p0[u].y; s[2] += p1[u].x; s[3] += p1[u].y; }
                }
            }
        }
        // lane tree: the five levels as a LOOP over the four sums (this code runs once per step out of a cold instruction
        // cache, where the bytes of code touched count, not the instructions executed); same operands, same order
#pragma unroll 1
        for (int w = 16; w > 0; w >>= 1) {
#pragma unroll
            for (int v = 0; v < 4; v++) s[v] += __shfl_xor_sync(full, s[v], w);
        }
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const double rt = sqrt(s[v]);
            if (grp == 0) norm[v] = rt;
            else if (grp == 1) parked[v] = rt;
            else norm[v] = norm[v] + (parked[v] + rt);
        }
    }
    return true;
}

__device__ __forceinline__ void finalize_block(const FinalizeArgs& a, int k, double* stage = nullptr, long long* fprof = nullptr) {
    const long long fp0 = fprof ? clock64() : 0;   // diagnostics of the persistent launch (tools/diag_persist.py); null in production
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    double norm[4];
    if (!finalize_block_norms(a, k, stage, norm)) return;
    if (fprof) fprof[0] = clock64() - fp0;         // strip sums, lane tree, square roots
    int last = 1;
    if (a.nowned > 1) {                            // (a single block has nobody to wait for: no ticket, no fences)
        if (lane == 0) {
            for (int v = 0; v < 4; v++) a.blk_norm[4 * k + v] = norm[v];
            __threadfence();
            last = atomicAdd(a.ticket, 1u) == (unsigned)(a.nowned - 1);
        }
        last = __shfl_sync(full, last, 0);
    } else if (lane == 0) {
        for (int v = 0; v < 4; v++) a.blk_norm[4 * k + v] = norm[v];
    }
    if (fprof) { fprof[1] = clock64() - fp0; fprof[2] = last; }   // + fence, ticket
    if (!last) return;
    if (a.nowned > 1) __threadfence();
    else __syncwarp();
    // last CTA: per-block sums by one lane per block (the global loads overlap), then lane 0 adds the
    // blocks in block order (main.cpp:379-392 gathers in rank order)
    const long long it = *a.iter;
    const int pos = *a.hist_pos;
    const StopArgs& sx = a.stop_x;
    const int slot = (int)((it + 1) & 3);
    double tot = 0.0, tot0 = 0.0;
    for (int q0 = 0; q0 < a.nowned; q0 += 32) {
        const int q = q0 + lane;
        double sum = 0.0, sum0 = 0.0;
        if (q < a.nowned) {
            for (int v = 0; v < 4; v++) {
                const double rs = __ldcg(a.blk_norm + 4 * q + v);
                double r0 = a.res0[4 * q + v];
                if (it == 0) r0 = rs;
                if (r0 < 1e-16) r0 = fmax(rs, 1e-16);
                a.res0[4 * q + v] = r0;
                sum += rs;
                sum0 += r0;
            }
            if (pos < a.hist_cap) {
                a.blk_hist[2 * ((long long)pos * a.nowned + q) + 0] = sum;
                a.blk_hist[2 * ((long long)pos * a.nowned + q) + 1] = sum0;
            }
            if (sx.enabled) {
                // this block's sums into every rank's residual table (the gather of main.cpp:379-392, one-sided)
                const long long o = 2 * ((long long)slot * sx.nblocks_global + a.blocks[q].gid);
                for (int r = 0; r < sx.nranks; r++) { sx.res_dst[r][o] = sum; sx.res_dst[r][o + 1] = sum0; }
            }
        }
        const int n = min(32, a.nowned - q0);
        for (int r = 0; r < n; r++) {           // block order
            tot += __shfl_sync(full, sum, r);
            tot0 += __shfl_sync(full, sum0, r);
        }
    }
    if (sx.enabled) __threadfence_system();
    __syncwarp();
    if (fprof) fprof[3] = clock64() - fp0;         // + fence, iteration counters, per-block sums
    if (lane != 0) return;
    if (sx.enabled) {
        for (int r = 0; r < sx.nranks; r++) fin_st_release_sys(sx.rflag_dst[r], (unsigned long long)(it + 1));
    }
    const double relres = tot / tot0;
    if (pos < a.hist_cap) {
        double* h = a.step_hist + 4 * (long long)pos;
        h[0] = __ldcg(a.dt_used); h[1] = tot; h[2] = tot0; h[3] = relres;
        a.flag_hist[pos] = __ldcg(a.flags);
    }
    *a.iter = it + 1;
    *a.hist_pos = pos + 1;
    if (a.nowned > 1) *a.ticket = 0u;
    // single rank: the local sums are the global ones (main.cpp:506-511); across ranks the test is made on the
    // published sums by the first CTA of a later launch (stop_test, march.cuh)
    if (a.stop_on_tol && !sx.enabled && relres < a.tol) *a.stop = 1;
    if (fprof) fprof[4] = clock64() - fp0;         // + history record
}
#endif

void preload_common();
void launch_finalize(cudaStream_t s, const FinalizeArgs& a);
// Deferred bookkeeping of `nsteps` steps whose strip sums, dts and error-bit snapshots were kept by a persistent launch
// (DeferArgs): the same arithmetic in the same order as finalize_block, step after step, as three short kernels.
void launch_finalize_deferred(cudaStream_t s, const FinalizeArgs& a, const DeferArgs& df, int nsteps, long long it0, int pos0, double* norm_ring, double* r0_ring);
void launch_aos_to_planes(cudaStream_t s, const BlockDev& b, int buf, const double* aos);
void launch_planes_to_aos(cudaStream_t s, const BlockDev& b, int buf, double* aos);
void launch_fill_uniform(cudaStream_t s, const BlockDev& b, int buf, const double u[4]);
void launch_nodes_to_planes(cudaStream_t s, const BlockDev& b, const double* x, const double* y);
void launch_cell_metric(cudaStream_t s, const BlockDev& b, int axi);
void launch_min_face(cudaStream_t s, const BlockDev& b, double* out, int nparts);
void launch_slopes_transpose(cudaStream_t s, const double* in, double* out, long long ncell, int to_soa);
void launch_face_normals(cudaStream_t s, const BlockDev& b, double2* out);
void launch_fp64_probe(cudaStream_t s, double* out, int grid, int iters);
}  // namespace par2d
