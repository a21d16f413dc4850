// march.cuh -- the production step kernel: register-marching warp strips.
//
// One warp (= one CTA) advances a strip of MCELLS = 32*CPL-2 cell columns by `march_rows` cell rows of
// one block, marching upwards row by row (FV/SpatialDiscretization.cpp:481-1011 + main.cpp:39-57,
// 307-351,394-404 fused):
//
//   lane l holds CPL (= MARCH_CPL, 1 or 2) adjacent cell columns c0-1+CPL*l ..: the first and the last
//   column of the warp are the strip's left/right halo columns (their state is loaded and derived,
//   but never updated), so a row is one contiguous 256*CPL-byte read per plane from the padded HBM
//   planes (ghost ring in place).
//   Rows are staged ahead of their use by TMA (cp.async.bulk.tensor, SASS UTMALDG): lane 0 issues ONE
//   tiled copy per two rows - a box of 32 columns x 2 rows x 7 planes (state + node coordinates + cell
//   metric) - into a per-warp shared-memory ring of MARCH_STAGES stages, each completing on its own
//   mbarrier.  No registers are held by loads in flight, no address
//   arithmetic is spent per plane, and HBM latency is covered by MARCH_STAGES-1 rows of compute.
//   (A tiled copy of 8-byte elements must start on a 16-byte boundary - tools/tma_probe - which is why
//   cell column 0 sits in the odd array column COL0 = 17: every strip's box starts in an even column.)
//   Per row and cell: derive the cell one row up, ONE i-face flux (left face; across lanes the left
//   cell's state comes from lane l-1 by warp shuffle and the right face's flux from lane l+1), ONE
//   j-face flux (top face; the bottom face is last row's top face, parked in shared memory
//   meanwhile), residual, norms, Jacobian solve, update, coalesced (CPL=2: 16-byte) stores.
//   Every state value is read from HBM once per strip (plus the halo column / row re-reads, 2/30 and
//   1/rows, which hit L2) and every face flux is evaluated once per strip.  The sum of squares is
//   accumulated per lane over the rows and reduced by warp shuffles in a fixed order once per strip.
//
// Why one warp per CTA: every loop bound then derives from blockIdx, ptxas proves the shuffles
// convergent and drops the divergence fallbacks (and their register moves).  Why the ring and the
// parked flux: the kernel is bound by warps x ILP under the register file; both take live values out
// of registers (ncu history in DESIGN.md).
//
// Compared with the first version of this step (a TMA-staged 32x8 tile kernel with block barriers, retired:
// profiles/r01_step_v1_tile_kernel_ncu_summary.txt): no block-wide barriers, ~1/6 of the shared-memory traffic,
// a fifth of the index arithmetic.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"
#include "physics.cuh"

namespace par2d {

#ifndef MARCH_CPL
#define MARCH_CPL 1
#endif
#ifndef MARCH_MINB
#define MARCH_MINB (MARCH_CPL == 2 ? 8 : 16)
#endif
#ifndef MARCH_YCC_FROM_FACES
#define MARCH_YCC_FROM_FACES 1
#endif
#ifndef MARCH_VOTE_J
#define MARCH_VOTE_J 1       // the fully-upwind warp vote also on j-faces (0: i-faces only, A/B runs)
#endif
#ifndef MARCH_STAGES
#define MARCH_STAGES 3    // ring stages of two rows each
#endif
constexpr int MARCH_THREADS = 32;                 // one warp strip per CTA

// ------------------------------------------------------------------ PTX helpers (TMA + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a mis-programmed copy must surface as an error, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); spin++) {
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
// descriptors live in global memory: make them visible to the tensor-map proxy before first use
__device__ __forceinline__ void tmap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ FaceState shfl_up1(const FaceState& s) {
    FaceState o;
    o.rho = shfl_up1(s.rho); o.u = shfl_up1(s.u); o.v = shfl_up1(s.v);
    o.p = shfl_up1(s.p); o.a = shfl_up1(s.a); o.h0 = shfl_up1(s.h0);
    return o;
}
__device__ __forceinline__ Flux shfl_down1(const Flux& f) {
    Flux o;
    o.f0 = shfl_down1(f.f0); o.f1 = shfl_down1(f.f1); o.f2 = shfl_down1(f.f2); o.f3 = shfl_down1(f.f3);
    o.parr = shfl_down1(f.parr);
    return o;
}

// per-warp shared memory: the row ring (TMA destination), its barriers and the parked bottom-face flux.
// One ring stage = ONE tiled copy of 2 array rows x 7 planes x WCOLS columns: the block's 11 planes are laid
// out [state A x4][x y metric][state B x4] so that the state being read and the node planes are always 7
// consecutive planes (A+nodes = planes 0..6, nodes+B = planes 4..10).  One copy per two rows instead of
// two copies per row: the lane-0 issue sequence (about 36 issue slots the whole warp pays for) runs a
// quarter as often.
constexpr int CPL = MARCH_CPL;                     // cell columns per lane
constexpr int WCOLS = 32 * CPL;                    // columns staged per row
constexpr int RING_PLANES = 7;                     // rho,u,v,T + x,y,metric in either order (see above)
constexpr int STAGE_ROWS = 2;
constexpr int PLANE_D = STAGE_ROWS * WCOLS;        // doubles between planes of a stage
constexpr int STAGE_D = RING_PLANES * PLANE_D;     // doubles per stage
struct MarchSmem {
    double ring[MARCH_STAGES][RING_PLANES][STAGE_ROWS][WCOLS];   // TMA box layout: [plane][row][column]
    double G[5][WCOLS];                            // j-face flux at the bottom of the current row
    uint64_t bar[MARCH_STAGES];
};
constexpr uint32_t STAGE_BYTES = STAGE_D * sizeof(double);

// What a lane carries in registers from one row to the next, per cell column: the derived state of the
// cell, its metric and the node at its lower-left corner; plus the node right of the last column.
struct RowState {
    FaceState s[CPL];
    double T[CPL];
    double im[CPL];          // 1/(vol*ycc) (axisymmetric) or 1/vol (planar) of cell (c, j): cell_metric_kernel
    double x[CPL + 1], y[CPL + 1];   // nodes (c, j) .. (c+CPL, j)
    double yfb[CPL];         // mid-point radius of the cell's bottom face (lean path: ycc = (bottom + top)/2)
};

struct MarchCtx {
    int k, c, ncx, ncy, pitch, lane;   // c = first cell column of this lane
    long long plane;
    bool writer[CPL];
    double dt;
    double* Sout;
};

// Residual, norms, solve and new state of one cell from its four face fluxes (left, right, bottom, top).
// (xa,ya)..(xd,yd) = nodes (c,j), (c+1,j), (c+1,j+1), (c,j+1)  (MeshModule.cpp:28-40).  Returns the new
// state in out[4] when a.update.
template <bool FAITHFUL, bool TPG, bool AXI, bool HOOKS>
__device__ __forceinline__ void cell_update(const StepArgs& a, const Gas& g, const MarchCtx& m, int c, int j, const FaceState& cur, double T, double im,
                                            double xa, double ya, double xb, double yb, double xc, double yc_, double xd, double yd,
                                            const Flux& Fl, const Flux& Fr, const Flux& Gb, const Flux& Gt,
                                            double out[4], double& sq0, double& sq1, double& sq2, double& sq3, unsigned& flags, bool live = true,
                                            double ycc_lean = 0.0) {
    const bool update = HOOKS ? a.update != 0 : true;   // HOOKS = false: the production step (no residual dump, always advancing)
    // centroid radius (MeshModule.cpp:40).  Lean path: the mean of the bottom and top face mid-points, which the
    // fluxes computed anyway (same value to an ulp, two operations instead of four, two fewer carried coordinates)
    const double ycc = (!FAITHFUL && MARCH_YCC_FROM_FACES) ? ycc_lean : 0.25 * (ya + yb + yc_ + yd);
    double r[4];        // residual = calc_dudt's dudt
    double dv[4];
    if constexpr (FAITHFUL) {
        // accumulation order of SpatialDiscretization.cpp:481-625 (interior faces: left, right,
        // bottom, top) then :643-977 (block-boundary faces: left, right, bottom, top)
        const double vol = cell_volume(xa, ya, xb, yb, xc, yc_, xd, yd);
        const double yc = AXI ? ycc : 0.0;
        const bool bl = c == 0, br = c == m.ncx - 1, bb = j == 0, bt = j == m.ncy - 1;
        const double fl[4] = {Fl.f0, Fl.f1, Fl.f2 + (yc * Fl.parr), Fl.f3};
        const double fr[4] = {Fr.f0, Fr.f1, Fr.f2 + (yc * Fr.parr), Fr.f3};
        const double gb[4] = {Gb.f0, Gb.f1, Gb.f2 + (yc * Gb.parr), Gb.f3};
        const double gt[4] = {Gt.f0, Gt.f1, Gt.f2 + (yc * Gt.parr), Gt.f3};
#pragma unroll
        for (int v = 0; v < 4; v++) {
            double acc = 0.0;
            if (!bl) acc += fl[v];
            if (!br) acc -= fr[v];
            if (!bb) acc -= gb[v];
            if (!bt) acc += gt[v];
            if (bl) acc += fl[v];
            if (br) acc -= fr[v];
            if (bb) acc -= gb[v];
            if (bt) acc += gt[v];
            r[v] = acc / vol;
        }
        if (update) {
            unsigned fl2 = 0;
            const double v2 = cur.u * cur.u + cur.v * cur.v;
            const double e0 = enthalpy<TPG>(g, T, fl2) - g.Rs * T + 0.5 * v2;   // StateVariables.h:62
            const double Cv = heat_capacity<TPG>(g, T, fl2) - g.Rs;
            solve_cell<true, AXI>(m.dt, ycc, cur.rho, cur.u, cur.v, e0, Cv, 0.0, r, dv, flags);
        }
    } else {
        double R[4];    // raw flux sums
        R[0] = (Fl.f0 - Fr.f0) + (Gt.f0 - Gb.f0);
        R[1] = (Fl.f1 - Fr.f1) + (Gt.f1 - Gb.f1);
        R[2] = (Fl.f2 - Fr.f2) + (Gt.f2 - Gb.f2);
        R[3] = (Fl.f3 - Fr.f3) + (Gt.f3 - Gb.f3);
        double ivol;
        if constexpr (AXI) {
            // axisymmetric source: the y-momentum pressure term carries the cell's own radius
            // (SpatialDiscretization.cpp:521-523,595-597); 1/vol and 1/ycc (main.cpp:325-331) come
            // from the stored metric 1/(vol*ycc)
            R[2] = fma(ycc, (Fl.parr - Fr.parr) + (Gt.parr - Gb.parr), R[2]);
            ivol = im * ycc;
        } else {
            ivol = im;
        }
        const double sc = m.dt * im;
#pragma unroll
        for (int v = 0; v < 4; v++) r[v] = R[v] * ivol;
        if (update) {
            unsigned fl2 = 0;
            const double e0 = fma(-g.Rs, T, cur.h0);   // h - Rs T + v2/2 with h0 = h + v2/2 at hand
            double inv_cv = g.inv_cv_cpg;
            if constexpr (TPG) inv_cv = fast_rcp(heat_capacity<TPG>(g, T, fl2) - g.Rs);
            solve_lean(sc, cur.rho, cur.u, cur.v, e0, inv_cv, R, dv);
        }
    }
    if (live) { sq0 = fma(r[0], r[0], sq0); sq1 = fma(r[1], r[1], sq1); sq2 = fma(r[2], r[2], sq2); sq3 = fma(r[3], r[3], sq3); }
    if constexpr (HOOKS) {
        if (a.res_out != nullptr && a.res_block == m.k) {
            double* o = a.res_out + 4 * ((long long)j * m.ncx + c);
            o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
        }
    }
    if (update) {
        // main.cpp:394-404: unk += dv, rho > 0, UpdateState's T clamp
        out[0] = cur.rho + dv[0]; out[1] = cur.u + dv[1]; out[2] = cur.v + dv[2];
        if constexpr (FAITHFUL) out[3] = fmax(T + dv[3], 201.0);
        else out[3] = max_sel(T + dv[3], 201.0);
        if (live && !(out[0] > 0.0)) flags |= FLAG_DENSITY;
        if constexpr (CPL == 1) {
            if (live) {
                double* o = m.Sout + ((long long)(j + 1) * m.pitch + COL0 + c);
                __stcg(o, out[0]); __stcg(o + m.plane, out[1]); __stcg(o + 2 * m.plane, out[2]); __stcg(o + 3 * m.plane, out[3]);
            }
        }
    }
}

// One row of a strip: consumes the carried state `lo` of cell row j and the staged row above it (srow/nrow
// = this lane's first column of the state / node planes of that row in the ring), produces the carried
// state `hi` of row j+1, updates this lane's cells of row j.
template <bool FAITHFUL, bool TPG, bool AXI, bool HOOKS>
__device__ __forceinline__ void march_row(const StepArgs& a, const Gas& g, const MarchCtx& m, MarchSmem& sm, int j, const double* srow, const double* nrow,
                                          const RowState& lo, RowState& hi, double& sq0, double& sq1, double& sq2, double& sq3, unsigned& flags) {
    const int q = CPL * m.lane;                    // this lane's first column in the staged row
    unsigned flu[CPL];
    double raw[RING_PLANES][CPL];                  // 0..3 state, 4..6 x, y, metric
    if constexpr (CPL == 2) {
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const double2 t = *reinterpret_cast<const double2*>(srow + v * PLANE_D);
            raw[v][0] = t.x; raw[v][1] = t.y;
        }
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const double2 t = *reinterpret_cast<const double2*>(nrow + v * PLANE_D);
            raw[4 + v][0] = t.x; raw[4 + v][1] = t.y;
        }
    } else {
#pragma unroll
        for (int v = 0; v < 4; v++) raw[v][0] = srow[v * PLANE_D];
#pragma unroll
        for (int v = 0; v < 3; v++) raw[4 + v][0] = nrow[v * PLANE_D];
    }
#pragma unroll
    for (int e = 0; e < CPL; e++) {
        flu[e] = 0;
        hi.T[e] = raw[3][e];
        hi.s[e] = face_state<FAITHFUL, TPG>(g, raw[0][e], raw[1][e], raw[2][e], raw[3][e], flu[e]);
        hi.x[e] = raw[4][e]; hi.y[e] = raw[5][e];  // node (c+e, j+1)
        hi.im[e] = raw[6][e];
    }
    // node right of this lane's last column = first node of lane+1 (for lane 31 it is never used)
    hi.x[CPL] = shfl_down1(hi.x[0]); hi.y[CPL] = shfl_down1(hi.y[0]);

    // i-faces of row j: face e between cell (c+e-1, j) [left state] and (c+e, j): node (c+e, j) -> (c+e, j+1);
    // the left state of face 0 is the last cell of lane-1, the flux of face CPL is face 0 of lane+1
    Flux F[CPL + 1];
    Flux Gt[CPL], Gb[CPL];
    {
    F[0] = ldfss<FAITHFUL, AXI, !FAITHFUL>(face_geom<FAITHFUL>(lo.x[0], lo.y[0], hi.x[0], hi.y[0]), shfl_up1(lo.s[CPL - 1]), lo.s[0]);
#pragma unroll
    for (int e = 1; e < CPL; e++)
        F[e] = ldfss<FAITHFUL, AXI, !FAITHFUL>(face_geom<FAITHFUL>(lo.x[e], lo.y[e], hi.x[e], hi.y[e]), lo.s[e - 1], lo.s[e]);
    F[CPL] = shfl_down1(F[0]);
    // j-faces: between cell (c+e, j+1) [left state = upper cell] and (c+e, j): node (c+e, j+1) -> (c+e+1, j+1)
#pragma unroll
    for (int e = 0; e < CPL; e++) {
        const FaceGeom gj = face_geom<FAITHFUL>(hi.x[e], hi.y[e], hi.x[e + 1], hi.y[e + 1]);
        hi.yfb[e] = gj.yf;
        Gt[e] = ldfss<FAITHFUL, AXI, !FAITHFUL && MARCH_VOTE_J>(gj, hi.s[e], lo.s[e]);
    }
    }
    // swap the parked bottom fluxes with the new top fluxes
    if constexpr (CPL == 2) {
        double2* G2 = reinterpret_cast<double2*>(&sm.G[0][q]);
        constexpr int stride = WCOLS / 2;
        double2 t;
        t = G2[0 * stride]; Gb[0].f0 = t.x; Gb[1].f0 = t.y; G2[0 * stride] = make_double2(Gt[0].f0, Gt[1].f0);
        t = G2[1 * stride]; Gb[0].f1 = t.x; Gb[1].f1 = t.y; G2[1 * stride] = make_double2(Gt[0].f1, Gt[1].f1);
        t = G2[2 * stride]; Gb[0].f2 = t.x; Gb[1].f2 = t.y; G2[2 * stride] = make_double2(Gt[0].f2, Gt[1].f2);
        t = G2[3 * stride]; Gb[0].f3 = t.x; Gb[1].f3 = t.y; G2[3 * stride] = make_double2(Gt[0].f3, Gt[1].f3);
        t = G2[4 * stride]; Gb[0].parr = t.x; Gb[1].parr = t.y; G2[4 * stride] = make_double2(Gt[0].parr, Gt[1].parr);
    } else {
        Gb[0].f0 = sm.G[0][q]; Gb[0].f1 = sm.G[1][q]; Gb[0].f2 = sm.G[2][q]; Gb[0].f3 = sm.G[3][q]; Gb[0].parr = sm.G[4][q];
        sm.G[0][q] = Gt[0].f0; sm.G[1][q] = Gt[0].f1; sm.G[2][q] = Gt[0].f2; sm.G[3][q] = Gt[0].f3; sm.G[4][q] = Gt[0].parr;
    }

    double out[CPL][4];
#pragma unroll
    for (int e = 0; e < CPL; e++) {
        if (m.writer[e]) {
            flags |= flu[e];   // state checks of the cell above: a real cell or a filled ghost
            cell_update<FAITHFUL, TPG, AXI, HOOKS>(a, g, m, m.c + e, j, lo.s[e], lo.T[e], lo.im[e], lo.x[e], lo.y[e], lo.x[e + 1], lo.y[e + 1],
                                            hi.x[e + 1], hi.y[e + 1], hi.x[e], hi.y[e], F[e], F[e + 1], Gb[e], Gt[e], out[e], sq0, sq1, sq2, sq3, flags, true,
                                            0.5 * (lo.yfb[e] + hi.yfb[e]));
        }
    }
    if constexpr (CPL == 2) {
        if (HOOKS ? a.update != 0 : true) {
            double* o = m.Sout + ((long long)(j + 1) * m.pitch + COL0 + m.c);
            if (m.writer[0] && m.writer[1]) {      // 16-byte stores: COL0 + c is even for every lane
#pragma unroll
                for (int v = 0; v < 4; v++) __stcg(reinterpret_cast<double2*>(o + v * m.plane), make_double2(out[0][v], out[1][v]));
            } else {
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (m.writer[e]) {
#pragma unroll
                        for (int v = 0; v < 4; v++) __stcg(o + v * m.plane + e, out[e][v]);
                    }
            }
        }
    }
}

template <bool FAITHFUL, bool TPG, bool AXI, bool HOOKS>
__global__ void __launch_bounds__(MARCH_THREADS, FAITHFUL ? 1 : MARCH_MINB) march_kernel(const StepArgs a, const Gas g) {
    __shared__ __align__(128) MarchSmem sm;
    if (*a.stop) return;
    const int lane = threadIdx.x;
    int k, strip;
    if (a.tile_list != nullptr) {                  // subset launch (edge strips first, interior strips beside the exchange)
        if ((int)blockIdx.x >= a.tile_count) return;
        const int2 t = __ldg(a.tile_list + blockIdx.x);
        k = t.x; strip = t.y;
    } else {
        k = blockIdx.y;                            // grid.y = owned blocks, grid.x covers the block with most strips
        strip = blockIdx.x;
    }
    const BlockDev& b = a.blocks[k];
    if (strip >= b.tiles_x * b.tiles_y) return;    // CTA-uniform
    MarchCtx m;
    m.k = k; m.ncx = b.ncx; m.ncy = b.ncy; m.pitch = b.pitch; m.plane = b.plane; m.lane = lane;
    const int c0 = (strip % b.tiles_x) * MCELLS;   // first updated cell column
    const int j0 = (strip / b.tiles_x) * a.march_rows;
    const int j1 = min(j0 + a.march_rows, m.ncy);
    m.c = c0 - 1 + CPL * lane;                     // this lane's first cell column (-1 and ncx are the ghost columns)
#pragma unroll
    for (int e = 0; e < CPL; e++) {
        const int w = CPL * lane + e;              // column within the staged row: 0 and WCOLS-1 are the halo columns
        m.writer[e] = w >= 1 && w <= MCELLS && m.c + e < m.ncx;
    }
    m.Sout = b.state[a.cur ^ 1];

    // ---- TMA row ring.  Array row `arow` holds cell row arow-1, node row arow-1 and the metric of cell
    // row arow-1; this strip reads array rows j0 (cell row below) .. j1+1 (cell row above the last).
    // Ring use n holds array rows j0+2n, j0+2n+1 (rows past the array end are zero-filled by TMA).
    const void* tmap = static_cast<const char*>(a.tmaps) + (size_t)k * 128;
    const int plane0 = a.cur ? 4 : 0;              // first of the 7 planes read this step
    const int col0 = COL0 + c0 - 1;                // even: the 16-byte rule of tiled copies
    const int arow_end = j1 + 1;
    auto stage_rows = [&](int n, int slot) {       // lane 0 only
        const int arow = j0 + 2 * n;
        if (arow <= arow_end) {
            mbar_expect_tx(&sm.bar[slot], STAGE_BYTES);
            tma_load_3d(&sm.ring[slot][0][0][0], tmap, col0, arow, plane0, &sm.bar[slot]);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < MARCH_STAGES; s++) mbar_init(&sm.bar[s], 1);
        mbar_init_fence();
        tmap_acquire(tmap);
#pragma unroll
        for (int s = 0; s < MARCH_STAGES; s++) stage_rows(s, s);
    }
    __syncwarp();
    // this lane's first column of the state planes / node planes in stage 0, row 0
    const double* rs = &sm.ring[0][a.cur ? 3 : 0][0][CPL * lane];
    const double* rn = &sm.ring[0][a.cur ? 0 : 4][0][CPL * lane];

    // global dt = max over blocks (main.cpp:270-292); the table was written by the ghost kernel
    double dt = 0.0;
    if (a.dt_ready) {
        dt = __ldcg(a.dt_used);                    // reduced by the ghost kernel
    } else {
        for (int t = lane; t < a.nblocks_global; t += 32) dt = fmax(dt, __ldcg(a.dt_table + t));
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) dt = fmax(dt, __shfl_xor_sync(0xffffffffu, dt, s));
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) *a.dt_used = dt;   // for the step's history record
    }
    m.dt = dt;

    unsigned flags = 0;
    bool ok = true;
    // ---- prologue (ring use 0): cell rows j0-1 (row 0) and j0 (row 1), node row j0, bottom-face fluxes of row j0
    RowState A, B;
    ok &= mbar_wait(&sm.bar[0], 0);
    {
        unsigned fl0[CPL];
        const int q = CPL * lane;
        const double* s1 = rs + WCOLS;             // row 1 of stage 0
        const double* n1 = rn + WCOLS;
#pragma unroll
        for (int e = 0; e < CPL; e++) {
            fl0[e] = 0;
            A.T[e] = s1[3 * PLANE_D + e];
            A.s[e] = face_state<FAITHFUL, TPG>(g, s1[e], s1[PLANE_D + e], s1[2 * PLANE_D + e], A.T[e], fl0[e]);
            A.x[e] = n1[e]; A.y[e] = n1[PLANE_D + e];
            A.im[e] = n1[2 * PLANE_D + e];
        }
        A.x[CPL] = shfl_down1(A.x[0]); A.y[CPL] = shfl_down1(A.y[0]);
#pragma unroll
        for (int e = 0; e < CPL; e++) {
            const FaceState lo = face_state<FAITHFUL, TPG>(g, rs[e], rs[PLANE_D + e], rs[2 * PLANE_D + e], rs[3 * PLANE_D + e], fl0[e]);
            // j-face: left state = upper cell, right state = lower cell (SpatialDiscretization.cpp:556-625)
            const FaceGeom g0 = face_geom<FAITHFUL>(A.x[e], A.y[e], A.x[e + 1], A.y[e + 1]);
            A.yfb[e] = g0.yf;
            const Flux G0 = ldfss<FAITHFUL, AXI, !FAITHFUL && MARCH_VOTE_J>(g0, A.s[e], lo);
            sm.G[0][q + e] = G0.f0; sm.G[1][q + e] = G0.f1; sm.G[2][q + e] = G0.f2; sm.G[3][q + e] = G0.f3; sm.G[4][q + e] = G0.parr;
            if (m.writer[e]) flags |= fl0[e];
        }
    }
    double sq0 = 0.0, sq1 = 0.0, sq2 = 0.0, sq3 = 0.0;
    // stage 0 is consumed: refill it
    __syncwarp();
    if (lane == 0) stage_rows(MARCH_STAGES, 0);

    // ---- march, two rows per trip (the carried state swaps roles, no register moves between rows): rows
    //      j, j+1 consume ring use n = (j-j0)/2 + 1, which is then refilled with use n + MARCH_STAGES
    int n = 1, slot = 1 % MARCH_STAGES;
    uint32_t ph = (1 / MARCH_STAGES) & 1;
    int j = j0;
#pragma unroll 1
    for (; j + 1 < j1; j += 2) {
        ok &= mbar_wait(&sm.bar[slot], ph);
        const double* s0 = rs + slot * STAGE_D;
        const double* n0 = rn + slot * STAGE_D;
        march_row<FAITHFUL, TPG, AXI, HOOKS>(a, g, m, sm, j, s0, n0, A, B, sq0, sq1, sq2, sq3, flags);
        march_row<FAITHFUL, TPG, AXI, HOOKS>(a, g, m, sm, j + 1, s0 + WCOLS, n0 + WCOLS, B, A, sq0, sq1, sq2, sq3, flags);
        __syncwarp();                              // every lane has read this stage
        if (lane == 0) stage_rows(n + MARCH_STAGES, slot);
        n++;
        if (++slot == MARCH_STAGES) { slot = 0; ph ^= 1u; }
    }
    if (j < j1) {
        ok &= mbar_wait(&sm.bar[slot], ph);
        march_row<FAITHFUL, TPG, AXI, HOOKS>(a, g, m, sm, j, rs + slot * STAGE_D, rn + slot * STAGE_D, A, B, sq0, sq1, sq2, sq3, flags);
    }

    // ---- sum of squares of this strip, fixed order: lanes (xor tree)
    double sq[4] = {sq0, sq1, sq2, sq3};
#pragma unroll
    for (int v = 0; v < 4; v++) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sq[v] += __shfl_xor_sync(0xffffffffu, sq[v], s);
    }
    if (isnan(sq[0]) || isnan(sq[1]) || isnan(sq[2]) || isnan(sq[3])) flags |= FLAG_NAN;   // SpatialDiscretization.cpp:1008
    if (!ok) flags |= FLAG_NAN | 0x80000000u;      // a staged row never arrived
    if (lane < 4) {
        const double s = lane == 0 ? sq[0] : lane == 1 ? sq[1] : lane == 2 ? sq[2] : sq[3];
        a.partials[4 * (long long)(b.tile0 + strip) + lane] = s;
    }
    if (flags) atomicOr(a.flags, flags);
}

}  // namespace par2d
