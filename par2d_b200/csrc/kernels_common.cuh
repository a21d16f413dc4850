// kernels_common.cuh -- launchers of the mode-independent kernels (kernels_common.cu).
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
namespace par2d {

// Residual bookkeeping of one step (main.cpp:39-57,365-392,494).  Called by one CTA of 256 threads per
// owned block k: sums the block's strip partials in a fixed order (thread-strided, then a shared-memory
// tree) and takes the square roots (main.cpp:53-56); the last CTA to finish adds the blocks in block
// order and appends dt, sum ||res||, sum res0 and relres to the device history.  No multiply-add is
// involved, so the result does not depend on the contraction mode of the including TU.
// Runs either as its own launch (finalize_kernel, end of a batch) or inside the NEXT step's ghost
// kernel (kernels.cuh): the bookkeeping of step n overlaps the ghost fill of step n+1.
#ifdef __CUDACC__
__device__ __forceinline__ void finalize_block(const FinalizeArgs& a, int k, double (*red)[4], bool* last) {
    const int tid = threadIdx.x;
    const BlockDev& b = a.blocks[k];
    double norm[4] = {0.0, 0.0, 0.0, 0.0};   // thread 0: the block's norms
    // first order: one group of 4 sums; second-order branch: three groups (dudt, duxdt, duydt) whose square roots
    // are added as main.cpp:360-364 does, ressum + (ressumx + ressumy)
    const int ngroups = a.dg_partials ? 3 : 1;
    double parked[4] = {0.0, 0.0, 0.0, 0.0};   // thread 0: ressumx while ressumy is being reduced
    for (int grp = 0; grp < ngroups; grp++) {
        const int ntiles = a.dg_partials ? b.dg_ntiles : b.tiles_x * b.tiles_y;
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int t = tid; t < ntiles; t += 256) {
            const double2* p = a.dg_partials ? reinterpret_cast<const double2*>(a.dg_partials + 12 * (long long)(b.dg_tile0 + t) + 4 * grp)
                                             : reinterpret_cast<const double2*>(a.partials + 4 * (long long)(b.tile0 + t));
            const double2 p0 = __ldcg(p), p1 = __ldcg(p + 1);
            s[0] += p0.x; s[1] += p0.y; s[2] += p1.x; s[3] += p1.y;
        }
        __syncthreads();
        for (int v = 0; v < 4; v++) red[tid][v] = s[v];
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (tid < w)
                for (int v = 0; v < 4; v++) red[tid][v] += red[tid + w][v];
            __syncthreads();
        }
        if (tid == 0) {
            for (int v = 0; v < 4; v++) {
                const double rt = sqrt(red[0][v]);
                if (grp == 0) norm[v] = rt;
                else if (grp == 1) parked[v] = rt;
                else norm[v] = norm[v] + (parked[v] + rt);
            }
        }
    }
    if (tid == 0) {
        for (int v = 0; v < 4; v++) a.blk_norm[4 * k + v] = norm[v];
        __threadfence();
        *last = atomicAdd(a.ticket, 1u) == (unsigned)(a.nowned - 1);
    }
    __syncthreads();
    if (!*last) return;
    __threadfence();
    // last CTA: per-block sums by one thread per block (the global loads overlap), then thread 0 adds the
    // blocks in block order (main.cpp:379-392 gathers in rank order)
    const long long it = *a.iter;
    const int pos = *a.hist_pos;
    double tot = 0.0, tot0 = 0.0;
    for (int q0 = 0; q0 < a.nowned; q0 += 256) {
        const int q = q0 + tid;
        if (q < a.nowned) {
            double sum = 0.0, sum0 = 0.0;
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
            red[tid][0] = sum;
            red[tid][1] = sum0;
        }
        __syncthreads();
        if (tid == 0) {
            const int n = min(256, a.nowned - q0);
            for (int r = 0; r < n; r++) { tot += red[r][0]; tot0 += red[r][1]; }
        }
        __syncthreads();
    }
    if (tid != 0) return;
    const double relres = tot / tot0;
    if (pos < a.hist_cap) {
        double* h = a.step_hist + 4 * (long long)pos;
        h[0] = __ldcg(a.dt_used); h[1] = tot; h[2] = tot0; h[3] = relres;
    }
    *a.iter = it + 1;
    *a.hist_pos = pos + 1;
    *a.ticket = 0u;
    if (a.stop_on_tol && relres < a.tol) *a.stop = 1;
}
#endif

void launch_finalize(cudaStream_t s, const FinalizeArgs& a);
void launch_aos_to_planes(cudaStream_t s, const BlockDev& b, int buf, const double* aos);
void launch_planes_to_aos(cudaStream_t s, const BlockDev& b, int buf, double* aos);
void launch_fill_uniform(cudaStream_t s, const BlockDev& b, int buf, const double u[4]);
void launch_nodes_to_planes(cudaStream_t s, const BlockDev& b, const double* x, const double* y);
void launch_cell_metric(cudaStream_t s, const BlockDev& b, int axi);
void launch_min_face(cudaStream_t s, const BlockDev& b, double* out, int nparts);
void launch_fp64_probe(cudaStream_t s, double* out, int grid, int iters);
}  // namespace par2d
