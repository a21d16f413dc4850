// kernels_common.cu -- mode-independent kernels: residual-norm finalisation and the
// reference's residual bookkeeping, AoS<->padded-SoA state transposes, the static part
// of find_dt, and an fp64 FMA throughput probe.  Compiled with -fmad=false: min_face
// must be the reference's value bit for bit (main.cpp:26-32, MeshModule.cpp:56,74).
#include <cuda_runtime.h>
#include "common.cuh"
#include "kernels_common.cuh"

namespace par2d {

// Stand-alone launch of the step bookkeeping (end of a batch; see finalize_block in kernels_common.cuh).
__global__ void __launch_bounds__(256) finalize_kernel(FinalizeArgs a) {
    __shared__ __align__(16) double stage[FIN_CHUNK * 4];
    if (*a.stop) return;
    finalize_block(a, blockIdx.x, stage);
}

void launch_finalize(cudaStream_t s, const FinalizeArgs& a) {
    finalize_kernel<<<a.nowned, 256, 0, s>>>(a);
}

// ---- deferred bookkeeping of a persistent launch (DeferArgs): finalize_block, one step after the other, as three kernels
// 1. the norms of every (step, block): the first part of finalize_block, one warp each
__global__ void __launch_bounds__(32) fin_norms_kernel(FinalizeArgs a, DeferArgs df, double* norm_ring) {
    const int step = blockIdx.x, k = blockIdx.y;
    a.partials = df.partials + step * df.stride;
    double norm[4];
    finalize_block_norms(a, k, nullptr, norm);
    if (threadIdx.x < 4) {
        const double n = threadIdx.x == 0 ? norm[0] : threadIdx.x == 1 ? norm[1] : threadIdx.x == 2 ? norm[2] : norm[3];
        norm_ring[4 * ((long long)step * a.nowned + k) + threadIdx.x] = n;
    }
}
// 2. the res0 rules (main.cpp:365-374) are the one thing that runs from step to step: one thread per (block, variable)
__global__ void fin_res0_kernel(FinalizeArgs a, int nsteps, long long it0, const double* __restrict__ norm_ring, double* __restrict__ r0_ring) {
    const int n4 = 4 * a.nowned;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += gridDim.x * blockDim.x) {
        double r0 = a.res0[e];
        for (int s0 = 0; s0 < nsteps; s0 += 8) {
            double rs[8];
#pragma unroll
            for (int u = 0; u < 8; u++) rs[u] = s0 + u < nsteps ? norm_ring[(long long)(s0 + u) * n4 + e] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (s0 + u < nsteps) {
                    if (it0 + s0 + u == 0) r0 = rs[u];
                    if (r0 < 1e-16) r0 = fmax(rs[u], 1e-16);
                    r0_ring[(long long)(s0 + u) * n4 + e] = r0;
                }
            }
        }
        a.res0[e] = r0;
    }
}
// 3. the history records: one thread per step adds the blocks in block order (main.cpp:379-392)
__global__ void fin_hist_kernel(FinalizeArgs a, DeferArgs df, int nsteps, long long it0, int pos0, const double* __restrict__ norm_ring, const double* __restrict__ r0_ring) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsteps) return;
    const int pos = pos0 + s;
    double tot = 0.0, tot0 = 0.0;
    for (int q = 0; q < a.nowned; q++) {
        double sum = 0.0, sum0 = 0.0;
        for (int v = 0; v < 4; v++) {
            sum += norm_ring[4 * ((long long)s * a.nowned + q) + v];
            sum0 += r0_ring[4 * ((long long)s * a.nowned + q) + v];
        }
        if (pos < a.hist_cap) {
            a.blk_hist[2 * ((long long)pos * a.nowned + q) + 0] = sum;
            a.blk_hist[2 * ((long long)pos * a.nowned + q) + 1] = sum0;
        }
        tot += sum;
        tot0 += sum0;
    }
    const double relres = tot / tot0;
    if (pos < a.hist_cap) {
        double* h = a.step_hist + 4 * (long long)pos;
        h[0] = df.dt[s]; h[1] = tot; h[2] = tot0; h[3] = relres;
        a.flag_hist[pos] = s + 1 < nsteps ? df.flags[s] : *a.flags;
    }
    if (s == 0) { *a.iter = it0 + nsteps; *a.hist_pos = pos0 + nsteps; }
}

void launch_finalize_deferred(cudaStream_t s, const FinalizeArgs& a, const DeferArgs& df, int nsteps, long long it0, int pos0, double* norm_ring, double* r0_ring) {
    fin_norms_kernel<<<dim3(nsteps, a.nowned), 32, 0, s>>>(a, df, norm_ring);
    const int n4 = 4 * a.nowned;
    fin_res0_kernel<<<(n4 + 127) / 128, 128, 0, s>>>(a, nsteps, it0, norm_ring, r0_ring);
    fin_hist_kernel<<<(nsteps + 127) / 128, 128, 0, s>>>(a, df, nsteps, it0, pos0, norm_ring, r0_ring);
}

// AoS unk[(j*ncx+i)*4+k] (FV/Indexing.h:11) -> padded planes; applies the T clamp that
// State::UpdateState performs on the initial state (main.cpp:226-238).
__global__ void aos_to_planes_kernel(BlockDev b, int buf, const double* __restrict__ aos) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= (long long)b.ncx * b.ncy) return;
    const int i = (int)(c % b.ncx), j = (int)(c / b.ncx);
    const double2 lo = reinterpret_cast<const double2*>(aos)[2 * c], hi = reinterpret_cast<const double2*>(aos)[2 * c + 1];
    double* S = b.state[buf];
    const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
    S[o] = lo.x; S[b.plane + o] = lo.y; S[2 * b.plane + o] = hi.x; S[3 * b.plane + o] = fmax(hi.y, 201.0);
}
__global__ void planes_to_aos_kernel(BlockDev b, int buf, double* __restrict__ aos) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= (long long)b.ncx * b.ncy) return;
    const int i = (int)(c % b.ncx), j = (int)(c / b.ncx);
    const double* S = b.state[buf];
    const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
    reinterpret_cast<double2*>(aos)[2 * c] = make_double2(S[o], S[b.plane + o]);
    reinterpret_cast<double2*>(aos)[2 * c + 1] = make_double2(S[2 * b.plane + o], S[3 * b.plane + o]);
}
__global__ void fill_uniform_kernel(BlockDev b, int buf, double r, double u, double v, double T) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= (long long)b.ncx * b.ncy) return;
    const int i = (int)(c % b.ncx), j = (int)(c / b.ncx);
    double* S = b.state[buf];
    const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
    S[o] = r; S[b.plane + o] = u; S[2 * b.plane + o] = v; S[3 * b.plane + o] = fmax(T, 201.0);
}
// slopes of the second-order branch: the reference's AoS ux[(j*ncx+i)*4+k] (main.cpp:169-172) <-> SoA planes [k][j*ncx+i]
__global__ void slopes_transpose_kernel(const double* __restrict__ in, double* __restrict__ out, long long ncell, int to_soa) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ncell) return;
    if (to_soa) {
        const double2 lo = reinterpret_cast<const double2*>(in)[2 * e], hi = reinterpret_cast<const double2*>(in)[2 * e + 1];
        out[e] = lo.x; out[ncell + e] = lo.y; out[2 * ncell + e] = hi.x; out[3 * ncell + e] = hi.y;
    } else {
        reinterpret_cast<double2*>(out)[2 * e] = make_double2(in[e], in[ncell + e]);
        reinterpret_cast<double2*>(out)[2 * e + 1] = make_double2(in[2 * ncell + e], in[3 * ncell + e]);
    }
}
void launch_slopes_transpose(cudaStream_t s, const double* in, double* out, long long ncell, int to_soa) {
    slopes_transpose_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, s>>>(in, out, ncell, to_soa);
}

// node coordinates x[j*nx+i] -> padded planes (row j+1, column COL0+i)
__global__ void nodes_to_planes_kernel(BlockDev b, const double* __restrict__ x, const double* __restrict__ y) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int nx = b.ncx + 1, ny = b.ncy + 1;
    if (c >= (long long)nx * ny) return;
    const int i = (int)(c % nx), j = (int)(c / nx);
    const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
    b.nodes[o] = x[c];
    b.nodes[b.plane + o] = y[c];
}

// Static cell metric of the fast path, third node plane: 1/(vol*ycc) (axisymmetric: the residual's 1/vol,
// MeshModule.cpp:34-36 + SpatialDiscretization.cpp:980-1011, times the 1/ycc of main.cpp:325-331) or 1/vol
// (planar), of cell (i,j) stored in node (i,j)'s slot.  vol is evaluated in the reference's order with
// individually rounded operations (this TU is compiled with -fmad=false, and cell_volume uses
// never-contracted intrinsics), the division is IEEE.
__global__ void cell_metric_kernel(BlockDev b, int axi) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= (long long)b.ncx * b.ncy) return;
    const int i = (int)(c % b.ncx), j = (int)(c / b.ncx);
    const double* X = b.nodes;
    const double* Y = b.nodes + b.plane;
    const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
    const double x1 = X[o], x2 = X[o + 1], x3 = X[o + b.pitch + 1], x4 = X[o + b.pitch];
    const double y1 = Y[o], y2 = Y[o + 1], y3 = Y[o + b.pitch + 1], y4 = Y[o + b.pitch];
    const double vol = cell_volume(x1, y1, x2, y2, x3, y3, x4, y4);
    const double ycc = 0.25 * (y1 + y2 + y3 + y4);
    b.nodes[2 * b.plane + o] = axi ? 1.0 / (vol * ycc) : 1.0 / vol;
}
void launch_cell_metric(cudaStream_t s, const BlockDev& b, int axi) {
    const long long n = (long long)b.ncx * b.ncy;
    cell_metric_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(b, axi);
}

void launch_aos_to_planes(cudaStream_t s, const BlockDev& b, int buf, const double* aos) {
    const long long n = (long long)b.ncx * b.ncy;
    aos_to_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(b, buf, aos);
}
void launch_planes_to_aos(cudaStream_t s, const BlockDev& b, int buf, double* aos) {
    const long long n = (long long)b.ncx * b.ncy;
    planes_to_aos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(b, buf, aos);
}
void launch_fill_uniform(cudaStream_t s, const BlockDev& b, int buf, const double u[4]) {
    const long long n = (long long)b.ncx * b.ncy;
    fill_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(b, buf, u[0], u[1], u[2], u[3]);
}
void launch_nodes_to_planes(cudaStream_t s, const BlockDev& b, const double* x, const double* y) {
    const long long n = (long long)(b.ncx + 1) * (b.ncy + 1);
    nodes_to_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(b, x, y);
}

// Outward unit normals of the boundary faces, CCW like ibound[] (generate_ghost_cells, SpatialDiscretization.cpp:
// 33-34,51-52,71-72,91-92): bottom +n(i,0), right +n of column 0 (the reference reads geofa[0,j], not column nx-1),
// top -n(i,ny-1), left -n(0,j), n = (dy,-dx)/len as MeshModule.cpp:52-81 - IEEE operations, no contraction (this TU).
__global__ void face_normals_kernel(BlockDev b, double2* out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int ncx = b.ncx, ncy = b.ncy;
    if (f >= 2 * (ncx + ncy)) return;
    const double* X = b.nodes;
    const double* Y = b.nodes + b.plane;
    long long p0, p1;
    double sgn;
    if (f < ncx) { p0 = (long long)b.pitch + COL0 + f; p1 = p0 + 1; sgn = 1.0; }
    else if (f < ncx + ncy) { p0 = (long long)(f - ncx + 1) * b.pitch + COL0; p1 = p0 + b.pitch; sgn = 1.0; }
    else if (f < 2 * ncx + ncy) { const int n = 2 * ncx + ncy - 1 - f; p0 = (long long)(ncy + 1) * b.pitch + COL0 + n; p1 = p0 + 1; sgn = -1.0; }
    else { const int n = 2 * ncx + 2 * ncy - 1 - f; p0 = (long long)(n + 1) * b.pitch + COL0; p1 = p0 + b.pitch; sgn = -1.0; }
    const FaceGeom fg = face_geom<true>(X[p0], Y[p0], X[p1], Y[p1]);
    out[f] = make_double2(sgn * fg.nx, sgn * fg.ny);
}
void launch_face_normals(cudaStream_t s, const BlockDev& b, double2* out) {
    const int n = 2 * (b.ncx + b.ncy);
    face_normals_kernel<<<(n + 255) / 256, 256, 0, s>>>(b, out);
}

// min over cells of the two face lengths leaving node (i,j), seeded with 1.0
// (main.cpp:26-32 over geofa[.,.,0] and geofa[.,.,3], MeshModule.cpp:56,74).  min is
// order independent, so a parallel reduction reproduces the serial value exactly.
__global__ void __launch_bounds__(256) min_face_kernel(BlockDev b, double* out) {
    __shared__ double red[256];
    const double* X = b.nodes;
    const double* Y = b.nodes + b.plane;
    double m = 1.0;
    const long long n = (long long)b.ncx * b.ncy;
    for (long long c = blockIdx.x * 256ll + threadIdx.x; c < n; c += 256ll * gridDim.x) {
        const int i = (int)(c % b.ncx), j = (int)(c / b.ncx);
        const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
        const double x0 = X[o], y0 = Y[o];
        const double xh = X[o + 1], yh = Y[o + 1], xv = X[o + b.pitch], yv = Y[o + b.pitch];
        const double lh = sqrt(((xh - x0) * (xh - x0)) + ((yh - y0) * (yh - y0)));
        const double lv = sqrt((xv - x0) * (xv - x0) + (yv - y0) * (yv - y0));
        m = fmin(m, lh);
        m = fmin(m, lv);
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}
void launch_min_face(cudaStream_t s, const BlockDev& b, double* out, int nparts) {
    min_face_kernel<<<nparts, 256, 0, s>>>(b, out);
}

// fp64 FMA throughput probe: 8 independent chains per thread.
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
void launch_fp64_probe(cudaStream_t s, double* out, int grid, int iters) {
    fp64_probe_kernel<<<grid, 256, 0, s>>>(out, iters, 0.5);
}

void preload_common() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, (const void*)finalize_kernel);
    cudaFuncGetAttributes(&fa, (const void*)fin_norms_kernel);
    cudaFuncGetAttributes(&fa, (const void*)fin_res0_kernel);
    cudaFuncGetAttributes(&fa, (const void*)fin_hist_kernel);
    cudaFuncGetAttributes(&fa, (const void*)aos_to_planes_kernel);
    cudaFuncGetAttributes(&fa, (const void*)planes_to_aos_kernel);
    cudaFuncGetAttributes(&fa, (const void*)fill_uniform_kernel);
    cudaFuncGetAttributes(&fa, (const void*)nodes_to_planes_kernel);
    cudaFuncGetAttributes(&fa, (const void*)cell_metric_kernel);
    cudaFuncGetAttributes(&fa, (const void*)min_face_kernel);
    cudaFuncGetAttributes(&fa, (const void*)face_normals_kernel);
    cudaFuncGetAttributes(&fa, (const void*)slopes_transpose_kernel);
    cudaGetLastError();
}

}  // namespace par2d
