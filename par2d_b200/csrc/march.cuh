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
#include "halo.cuh"
#include "kernels_common.cuh"

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
#define MARCH_VOTE_J 0       // the fully-upwind warp vote on i-faces only.  1: also on j-faces - which follow the walls and the
                             // streamlines in every shipped case and in the benchmark, so their normal Mach number is subsonic,
                             // the vote never passes and its test + branch is all cost: 0.5941/0.5862 -> 0.5735 ms per launch
                             // (profiles/r02_ab_early_vote.txt)
#endif
#ifndef PUSH_EXP
#define PUSH_EXP 31     // diagnostics: bit mask of the fused step's pieces that are compiled in (1 bookkeeping role + stop test, 2 pull,
                        // 4 epilogue, 8 flag wait, 16 proxy fence); anything but 31 gives wrong results and is for timing runs only
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

// PlaneT: the type of the plane stride.  The tuned launch-per-step kernel keeps it in 64 bits (precomputed multiples
// in registers); the fused kernel, which has fewer registers to spare across the row loop, in 32 (one IMAD.WIDE per
// plane; < 2^31 doubles per plane is checked by par2d_create).
template <class PlaneT>
struct MarchCtxT {
    int k, c, ncx, ncy, pitch, lane;   // c = first cell column of this lane
    PlaneT plane;
    bool writer[CPL];
    double dt;
    double* Sout;
};
using MarchCtx = MarchCtxT<long long>;

// Residual, norms, solve and new state of one cell from its four face fluxes (left, right, bottom, top).
// (xa,ya)..(xd,yd) = nodes (c,j), (c+1,j), (c+1,j+1), (c,j+1)  (MeshModule.cpp:28-40).  Returns the new
// state in out[4] when a.update.
template <bool FAITHFUL, bool TPG, bool AXI, bool HOOKS, class CTX>
__device__ __forceinline__ void cell_update(const StepArgs& a, const Gas& g, const CTX& m, int c, int j, const FaceState& cur, double T, double im,
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
template <bool FAITHFUL, bool TPG, bool AXI, bool HOOKS, class CTX>
__device__ __forceinline__ void march_row(const StepArgs& a, const Gas& g, const CTX& m, MarchSmem& sm, int j, const double* srow, const double* nrow,
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

// What changes from one step to the next inside a batch (the persistent launch advances these in the kernel).
struct StepVar {
    int cur;                 // state buffer read
    const double* dt_table;  // per-block dt of the state being read
    double* dt_next;         // per-block dt of the state being written (fused step only)
    double* dt_used;         // the global dt of this step, for the history record
    double* partials;        // this step's strip sums
    int parity;              // peer-mapped mode: parity of the exchange event this step produces
    unsigned long long seq;  //   ... and its number (the step consumes event StepArgs::wait_seq = seq-1)
    int first;               // first step of this CTA in the launch: make the tensor map visible to the TMA unit
    long long* prof;         // diagnostics: this strip's phase time stamps of this step (null in production)
};

// ------------------------------------------------------------------ fused step: the perimeter of a block
// Everything the fused step adds to a strip lives in NON-INLINED functions that take their arguments by value or as
// pointers to global memory: the row loop of march_tile is tuned to the last register and the last branch, and ptxas
// only keeps it that way (convergence of the warp proven, no register moves between the two rows of a trip) while
// the strip's own code path contains no divergent loop and no spill-prone live range - measured: any of these
// pieces inlined costs 8-10 % more instructions in the row loop (and even so the fused kernel is ~10 % slower per
// launch than the launch-per-step kernel below - which is why large meshes keep that one).

// One side of a perimeter strip that borders another RANK, after its rows are done (the strip's new cells sit in Sout,
// visible to the whole warp): the new edge cells n of [n0, n1) straight into the neighbour rank's arena over NVLink, one
// 32-byte AoS record per edge cell (usend[k + NVAR*j], SpatialDiscretization.cpp:236-259).
__device__ __forceinline__ void push_side(const BlockDev& b, int out_buf, int parity, int side, int n0, int n1) {
    if (b.put_dst[side] == nullptr) return;
    const double* Sout = b.state[out_buf];
    for (int n = n0 + (int)threadIdx.x; n < n1; n += 32) {
        int c, r;
        edge_cell_pos(b.ncx, b.ncy, side, n, c, r);
        const long long io = (long long)r * b.pitch + c;
        double2* d2 = reinterpret_cast<double2*>(b.put_dst[side] + parity * b.put_stride[side] + 4 * (long long)n);
        d2[0] = make_double2(__ldcg(Sout + io), __ldcg(Sout + b.plane + io));
        d2[1] = make_double2(__ldcg(Sout + 2 * b.plane + io), __ldcg(Sout + 3 * b.plane + io));
    }
}

// what the epilogue needs of PutArgs, by value
struct PutLite {
    int peer_mode, parity, nranks, rank;
    unsigned long long seq;
    double* const* dt_dst;
    long long dt_stride;
    unsigned long long* const* flag_dst;
    unsigned* ticket;
    unsigned expected;
};

// An L2 load that stays where it is written: the compiler otherwise sinks a load towards its first use, past branches
// that do not need the value - and turns one round trip shared by several loads into one round trip each.
__device__ __forceinline__ double ldcg_here(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// After the rows of a perimeter strip: ghost ring / edge lines of the NEW state, the block's dt, the put ticket.
// Returns error bits.
static __device__ __noinline__ unsigned tile_epilogue(const BlockDev* bk, const Gas* gas, const GhostOp* ops, int op0, int nop, int strip,
                                                     int out_buf, int march_rows, double* dt_next, double gamRs, double cfl, PutLite put, long long* prof) {
    const long long tq = prof ? clock64() : 0;
    const int lane = threadIdx.x;
    unsigned flags = 0;
    __syncwarp();                                  // the strip's new cells are visible to every lane
    if (prof) prof[0] = clock64() - tq;
    // strip 0 holds cell (0,0) of its block: the three values find_dt reads there are requested NOW and used after the
    // records - this code is a chain of L2 round trips, and every one that can overlap another must
    double q1 = 0.0, q2 = 0.0, q3 = 0.0;
    if (strip == 0) {                              // CTA-uniform
        const BlockDev& b = *bk;
        const double* S = b.state[out_buf];
        const long long o = (long long)b.pitch + COL0;
        q1 = ldcg_here(S + b.plane + o); q2 = ldcg_here(S + 2 * b.plane + o); q3 = ldcg_here(S + 3 * b.plane + o);
    }
    // boundary-condition faces and faces shared with a block of this rank: one record per ghost cell.  The loads of a
    // record do not depend on its kind - the whole record in four 16-byte loads, then the edge cell - so lanes holding
    // different kinds wait for ONE round trip to the record and ONE to the cell together (a branch per kind around its
    // own loads made every kind present in the warp pay both round trips in turn: 6300 of the 13500 cycles of a corner
    // strip's step on the shipped cases, profiles/r02_persistent_launch_profile.txt); the kinds then differ in a few
    // arithmetic instructions only.
    if (nop > 0) {
        const double gRs = __ldg(&gas->gamRs);
        const double fs0 = __ldg(&gas->uFS[0]), fs1 = __ldg(&gas->uFS[1]), fs2 = __ldg(&gas->uFS[2]), fs3 = __ldg(&gas->uFS[3]);
#pragma unroll 1
        for (int i = lane; i < nop; i += 32) {     // one trip for all but the corner strips of wide blocks: not unrolled, this code runs cold
            const GhostOp* op = ops + op0 + i;
            const ulonglong2 sp = __ldg(reinterpret_cast<const ulonglong2*>(&op->src[0]));
            const ulonglong2 dp = __ldg(reinterpret_cast<const ulonglong2*>(&op->dst[0]));
            const int4 meta = __ldg(reinterpret_cast<const int4*>(&op->src_plane));     // src_plane, dst_plane, kind, blk
            const double2 nrm = __ldg(reinterpret_cast<const double2*>(&op->nx));
            const double* src = reinterpret_cast<const double*>(out_buf ? sp.y : sp.x);
            double* dst = reinterpret_cast<double*>(out_buf ? dp.y : dp.x);
            const double u0 = ldcg_here(src), u1 = ldcg_here(src + meta.x), u2 = ldcg_here(src + 2 * (long long)meta.x), u3 = ldcg_here(src + 3 * (long long)meta.x);
            const int kind = meta.z;
            // slip wall: V - 2 (V.n) n, BoundaryConditions.cpp:79-88 (never-contracted operations, as bc_ghost_fill)
            const double vn = __dadd_rn(__dmul_rn(u1, nrm.x), __dmul_rn(u2, nrm.y));
            const double t = __dmul_rn(2.0, vn);
            const double w1 = __dsub_rn(u1, __dmul_rn(t, nrm.x)), w2 = __dsub_rn(u2, __dmul_rn(t, nrm.y));
            double g0 = u0, g1 = kind == 1 ? w1 : u1, g2 = kind == 1 ? w2 : u2, g3 = u3;
            bool general = kind == 2;
            if (kind == 3) {
                // far field (code 1, BoundaryConditions.cpp:148-188) where the normal velocity is supersonic - the usual
                // state of these faces: vn/a <= -1 takes the free stream, vn/a >= 1 the interior state.  For a > 0 the
                // quotient's comparisons with -1 and 1 are exactly those of vn with -a and a, so no division is needed.
                const double aint = sqrt(__dmul_rn(gRs, u3));
                if (aint > 0.0 && (vn <= -aint || vn >= aint)) {
                    const bool in = vn <= -aint;
                    g0 = in ? fs0 : u0; g1 = in ? fs1 : u1; g2 = in ? fs2 : u2; g3 = in ? fs3 : u3;
                } else {
                    general = true;
                }
            }
            if (general) {                         // back pressure, Bernoulli inflow, subsonic far field, unknown codes
                const BlockDev& b = *bk;           // a record's block (GhostOp::blk) is the block of the strip that holds it
                double* S = b.state[out_buf];
                const int2 fc = __ldg(reinterpret_cast<const int2*>(&op->f));
                bc_ghost_general(gas, &b, S, fc.y, fc.x, src - S, dst - S);
            } else {
                __stcg(dst, g0); __stcg(dst + meta.y, g1); __stcg(dst + 2 * (long long)meta.y, g2); __stcg(dst + 3 * (long long)meta.y, fmax(g3, 201.0));
            }
        }
    }
    if (prof) prof[1] = clock64() - tq;
    bool put_something = false;
    if (put.peer_mode) {
        // faces shared with a block of another rank: the new edge cells straight into that rank's arena
        const BlockDev& b = *bk;
        const int ex = strip % b.tiles_x, ey = strip / b.tiles_x;
        const bool tl = ex == 0, tr = ex == b.tiles_x - 1, tb = ey == 0, tt = ey == b.tiles_y - 1;
        const int c0 = ex * MCELLS, c1 = min(c0 + MCELLS, b.ncx);
        const int j0 = ey * march_rows, j1 = min(j0 + march_rows, b.ncy);
        const bool rl = tl && b.nb_kind[SIDE_LEFT] == NB_REMOTE, rr = tr && b.nb_kind[SIDE_RIGHT] == NB_REMOTE;
        const bool rb = tb && b.nb_kind[SIDE_BOTTOM] == NB_REMOTE, rt = tt && b.nb_kind[SIDE_TOP] == NB_REMOTE;
        if (rl) push_side(b, out_buf, put.parity, SIDE_LEFT, j0, j1);
        if (rr) push_side(b, out_buf, put.parity, SIDE_RIGHT, j0, j1);
        if (rb) push_side(b, out_buf, put.parity, SIDE_BOTTOM, c0, c1);
        if (rt) push_side(b, out_buf, put.parity, SIDE_TOP, c0, c1);
        put_something = rl || rr || rb || rt || strip == 0;
    }
    if (prof) prof[2] = clock64() - tq;
    if (strip == 0 && lane == 0) {
        // find_dt of the NEW state from cell (0,0) (main.cpp:19-37), for the next step on every rank
        const BlockDev& b = *bk;
        const double dtb = block_dt_math(q1, q2, q3, b.min_face, gamRs, cfl);
        if (isnan(dtb)) flags |= FLAG_NAN;
        if (put.peer_mode) {
            for (int r = 0; r < put.nranks; r++) put.dt_dst[r][put.parity * put.dt_stride + b.gid] = dtb;
        } else {
            dt_next[b.gid] = dtb;
        }
    }
    if (put_something) {
        // all stores of this strip are ordered before its ticket; the last perimeter strip of the launch raises this
        // rank's flag in every peer's arena (release at system scope)
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            if (atomicAdd(put.ticket, 1u) == put.expected - 1) {
                *put.ticket = 0u;
                __threadfence_system();
                for (int r = 0; r < put.nranks; r++)
                    if (r != put.rank) st_release_sys(put.flag_dst[r], put.seq);
            }
        }
    }
    __syncwarp();
    if (prof) prof[3] = clock64() - tq;
    return flags;
}

// Multi-rank step, before a strip's rows: the peers' exchange event (their edge lines and dts of the state being
// read) must have arrived; a strip that borders another rank copies the received lines of its faces into the ghost
// ring of that state, ahead of the TMA copies that stage it.  Returns error bits.
static __device__ __noinline__ unsigned tile_prologue_multi(const BlockDev* blocks, int k, int strip, int in_buf, int march_rows, int pull,
                                                           const unsigned long long* wait_flags, int nranks, int rank, unsigned long long wait_seq,
                                                           long long timeout_ns, long long recv_off) {
    unsigned flags = 0;
    if (!wait_peers_warp(wait_flags, nranks, rank, wait_seq, timeout_ns)) flags |= FLAG_COMM;
    if (pull) {
        const BlockDev& b = blocks[k];
        const int ex = strip % b.tiles_x, ey = strip / b.tiles_x;
        const int c0 = ex * MCELLS, c1 = min(c0 + MCELLS, b.ncx);
        const int j0 = ey * march_rows, j1 = min(j0 + march_rows, b.ncy);
        double* S = b.state[in_buf];
        for (int side = 0; side < 4; side++) {
            const bool touches = side == SIDE_LEFT ? ex == 0 : side == SIDE_RIGHT ? ex == b.tiles_x - 1 : side == SIDE_BOTTOM ? ey == 0 : ey == b.tiles_y - 1;
            if (!touches || b.nb_kind[side] != NB_REMOTE) continue;
            const bool vertical = side == SIDE_LEFT || side == SIDE_RIGHT;
            const int n0 = vertical ? j0 : c0, n1 = vertical ? j1 : c1;
            const double* line0 = b.recv[side] + recv_off;   // parity of the consumed event
            for (int n = n0 + (int)threadIdx.x; n < n1; n += 32) {
                int c, r;
                ghost_cell_pos(b.ncx, b.ncy, side, n, c, r);
                const long long go = (long long)r * b.pitch + c;
                const double* line = line0 + 4 * (long long)n;
                const double2 lo = __ldcg(reinterpret_cast<const double2*>(line)), hi = __ldcg(reinterpret_cast<const double2*>(line) + 1);
                __stcg(S + go, lo.x); __stcg(S + b.plane + go, lo.y); __stcg(S + 2 * b.plane + go, hi.x); __stcg(S + 3 * b.plane + go, fmax(hi.y, 201.0));
            }
        }
        fence_proxy_async_global();
    }
    __syncwarp();
    return flags;
}

// Stop test across ranks (see StopArgs), launch `launch_no`: the launch would overwrite the state left by iteration
// it-1 - it may only run if that iteration's global relres stayed above the tolerance.
__device__ __forceinline__ void stop_test(const StopArgs& sx, long long it_checked, int* stop, unsigned* flags, int launch_no) {
    if (it_checked >= 0 && !*stop) {
        if (!wait_peers(sx.local_flags, sx.nranks, -1, (unsigned long long)(it_checked + 1), sx.timeout_ns, 1)) atomicOr(flags, FLAG_COMM);
        const double* t = sx.res_table + 2 * (long long)((it_checked + 1) & 3) * sx.nblocks_global;
        double tot = 0.0, tot0 = 0.0;
        for (int q = 0; q < sx.nblocks_global; q++) { tot += __ldcg(t + 2 * q); tot0 += __ldcg(t + 2 * q + 1); }   // rank order of main.cpp:381-392
        if (tot / tot0 < sx.tol) { *sx.stop_iter = it_checked; *stop = 1; }
    }
    __threadfence();
    *reinterpret_cast<volatile int*>(sx.go) = launch_no;
}
// ... and what every other CTA of that launch does meanwhile: wait for the decision.  Returns the stop word.
static __device__ __noinline__ int stop_gate(const int* go, int launch_no, const int* stop, unsigned* flags, long long timeout_ns) {
    if (threadIdx.x == 0) {
        const unsigned long long t0 = global_ns();
        while (*reinterpret_cast<const volatile int*>(go) < launch_no) {
            if ((long long)(global_ns() - t0) > timeout_ns) { atomicOr(flags, FLAG_COMM); break; }
            __nanosleep(100);
        }
        __threadfence();
    }
    __syncwarp();
    return *reinterpret_cast<const volatile int*>(stop);
}

// ------------------------------------------------------------------ one strip, one step
// `reload(k, strip)` re-derives the strip's identity after the row loop (from blockIdx or from the tile list), so
// that neither it nor anything computed from it has to stay in a register across the rows.
template <bool FAITHFUL, bool TPG, bool AXI, bool HOOKS, bool PUSH, class Reload>
__device__ __forceinline__ void march_tile(const StepArgs& a, const Gas& g, const PutArgs& put, const StepVar& sv, MarchSmem& sm, int k_in, int strip_in,
                                           const BlockDev* bk_in, Reload reload) {
    const int lane = threadIdx.x;
    const long long te = (PUSH && sv.prof) ? clock64() : 0;
    int k = k_in, strip = strip_in;
    const BlockDev* bk = bk_in;                    // the strip's block descriptor: a.blocks + k, or the persistent launch's copy in shared memory
    const BlockDev& b = *bk;
    MarchCtxT<int> m;
    m.k = k; m.ncx = b.ncx; m.ncy = b.ncy; m.pitch = b.pitch; m.plane = (int)b.plane; m.lane = lane;
    const int tx = strip % b.tiles_x, ty = strip / b.tiles_x;
    const int c0 = tx * MCELLS;                    // first updated cell column
    const int j0 = ty * a.march_rows;
    const int j1 = min(j0 + a.march_rows, m.ncy);
    m.c = c0 - 1 + CPL * lane;                     // this lane's first cell column (-1 and ncx are the ghost columns)
#pragma unroll
    for (int e = 0; e < CPL; e++) {
        const int w = CPL * lane + e;              // column within the staged row: 0 and WCOLS-1 are the halo columns
        m.writer[e] = w >= 1 && w <= MCELLS && m.c + e < m.ncx;
    }
    m.Sout = b.state[sv.cur ^ 1];
    unsigned flags = 0;
    bool ok = true;

    // multi-rank: a strip that borders another rank pulls its ghost cells out of the arena BEFORE staging its rows;
    // every other strip waits for the peers' event (their dts) while its first rows are in flight
    bool pull = false;
    if constexpr (PUSH && (PUSH_EXP & 2)) {
        if (a.wait_flags != nullptr) {
            pull = (tx == 0 && b.nb_kind[SIDE_LEFT] == NB_REMOTE) || (tx == b.tiles_x - 1 && b.nb_kind[SIDE_RIGHT] == NB_REMOTE) ||
                   (ty == 0 && b.nb_kind[SIDE_BOTTOM] == NB_REMOTE) || (ty == b.tiles_y - 1 && b.nb_kind[SIDE_TOP] == NB_REMOTE);
            if (pull) flags |= tile_prologue_multi(a.blocks, k, strip, sv.cur, a.march_rows, 1, a.wait_flags, a.nranks, a.rank, a.wait_seq, a.timeout_ns, a.recv_off);
        }
    }

    // ---- TMA row ring.  Array row `arow` holds cell row arow-1, node row arow-1 and the metric of cell
    // row arow-1; this strip reads array rows j0 (cell row below) .. j1+1 (cell row above the last).
    // Ring use n holds array rows j0+2n, j0+2n+1 (rows past the array end are zero-filled by TMA).
    const void* tmap = static_cast<const char*>(a.tmaps) + (size_t)k * 128;
    const int plane0 = sv.cur ? 4 : 0;             // first of the 7 planes read this step
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
        if (sv.first) tmap_acquire(tmap);
        if constexpr (PUSH && (PUSH_EXP & 16)) fence_proxy_async_global();   // the state was written by other CTAs' generic stores (persistent launch)
#pragma unroll
        for (int s = 0; s < MARCH_STAGES; s++) stage_rows(s, s);
    }
    __syncwarp();
    if (PUSH && sv.prof) sv.prof[6] = clock64() - te;     // entry -> first rows requested
    // this lane's first column of the state planes / node planes in stage 0, row 0
    const double* rs = &sm.ring[0][sv.cur ? 3 : 0][0][CPL * lane];
    const double* rn = &sm.ring[0][sv.cur ? 0 : 4][0][CPL * lane];

    // global dt = max over blocks (main.cpp:270-292); the table was written by the ghost kernel (first step of a
    // batch) or by the strips holding cell (0,0) of every block in the previous step, on any rank
    if constexpr (PUSH && (PUSH_EXP & 8)) {
        if (a.wait_flags != nullptr && !pull)
            flags |= tile_prologue_multi(a.blocks, k, strip, sv.cur, a.march_rows, 0, a.wait_flags, a.nranks, a.rank, a.wait_seq, a.timeout_ns, a.recv_off);
    }
    double dt = 0.0;
    if (a.dt_ready) {
        dt = __ldcg(sv.dt_used);                   // reduced by the ghost kernel
    } else {
        for (int t = lane; t < a.nblocks_global; t += 32) dt = fmax(dt, __ldcg(sv.dt_table + t));
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) dt = fmax(dt, __shfl_xor_sync(0xffffffffu, dt, s));
        if (k == 0 && strip == 0 && lane == 0) *sv.dt_used = dt;   // for the step's history record
    }
    m.dt = dt;

    // ---- prologue (ring use 0): cell rows j0-1 (row 0) and j0 (row 1), node row j0, bottom-face fluxes of row j0
    RowState A, B;
    const long long tp0 = (PUSH && sv.prof) ? clock64() : 0;
    if (PUSH && sv.prof) sv.prof[-1] = tp0 - te;          // entry -> dt known, waiting for the rows
    ok &= mbar_wait(&sm.bar[0], 0);
    if (PUSH && sv.prof) sv.prof[0] = clock64() - tp0;    // wait for the first staged rows
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

    // ---- march, ONE row per trip (ring use n = (j-j0)/2 + 1 holds rows j, j+1 and is refilled with use n + MARCH_STAGES
    //      once both are consumed).  The launch-per-step kernel below unrolls two rows per trip to save the register
    //      moves of the carried state; here the code SIZE counts instead: the fused / persistent step runs everything
    //      once per step, and a per-step code footprint beyond the SM's instruction cache makes every line of it miss
    //      every step (measured on the persistent launch: ~40 cycles per instruction outside the loop).
    int n = 1, slot = 1 % MARCH_STAGES;
    uint32_t ph = (1 / MARCH_STAGES) & 1;
#pragma unroll 1
    for (int j = j0; j < j1; j++) {
        const int r = (j - j0) & 1;                // row within the stage
        if (r == 0) ok &= mbar_wait(&sm.bar[slot], ph);
        march_row<FAITHFUL, TPG, AXI, HOOKS>(a, g, m, sm, j, rs + slot * STAGE_D + r * WCOLS, rn + slot * STAGE_D + r * WCOLS, A, B, sq0, sq1, sq2, sq3, flags);
        A = B;
        if (r == 1) {
            __syncwarp();                          // every lane has read this stage
            if (lane == 0) stage_rows(n + MARCH_STAGES, slot);
            n++;
            if (++slot == MARCH_STAGES) { slot = 0; ph ^= 1u; }
        }
    }

    if (PUSH && sv.prof) sv.prof[1] = clock64() - tp0;    // ... + prologue + rows
    // ---- sum of squares of this strip, fixed order: lanes (xor tree)
    double sq[4] = {sq0, sq1, sq2, sq3};
#pragma unroll
    for (int v = 0; v < 4; v++) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sq[v] += __shfl_xor_sync(0xffffffffu, sq[v], s);
    }
    if (isnan(sq[0]) || isnan(sq[1]) || isnan(sq[2]) || isnan(sq[3])) flags |= FLAG_NAN;   // SpatialDiscretization.cpp:1008
    if (!ok) flags |= FLAG_NAN | 0x80000000u;      // a staged row never arrived
    reload(k, strip, bk);
    const BlockDev& bb = *bk;
    if (lane < 4) {
        const double s = lane == 0 ? sq[0] : lane == 1 ? sq[1] : lane == 2 ? sq[2] : sq[3];
        sv.partials[4 * (long long)(bb.tile0 + strip) + lane] = s;
    }

    if (PUSH && sv.prof) sv.prof[2] = clock64() - tp0;    // ... + norms, strip identity
    if constexpr (PUSH && (PUSH_EXP & 4)) {
        const int2 rg = __ldg(a.op_range + bb.tile0 + strip);       // this strip's ghost-cell records
        bool epi = rg.y > 0 || strip == 0;
        if (put.peer_mode) {
            const int ntx = bb.tiles_x, ex = strip % ntx, ey = strip / ntx;
            epi = epi || ex == 0 || ex == ntx - 1 || ey == 0 || ey == bb.tiles_y - 1;
        }
        if (epi) {                                 // CTA-uniform: the strip touches the block's perimeter
            PutLite pl;
            pl.peer_mode = put.peer_mode; pl.parity = sv.parity; pl.nranks = put.nranks; pl.rank = put.rank; pl.seq = sv.seq;
            pl.dt_dst = put.dt_dst; pl.dt_stride = put.dt_stride; pl.flag_dst = put.flag_dst; pl.ticket = put.ticket; pl.expected = put.expected;
            flags |= tile_epilogue(bk, a.gas_dev, a.ops, rg.x, rg.y, strip, sv.cur ^ 1, a.march_rows, sv.dt_next, g.gamRs, g.cfl, pl,
                                   sv.prof ? sv.prof + 4 : nullptr);
        }
    }
    if (PUSH && sv.prof) sv.prof[3] = clock64() - tp0;    // ... + perimeter epilogue
    if (flags) atomicOr(a.flags, flags);
}

// The finalize role inside the fused step kernel: CTA `k` (one warp) runs the previous step's bookkeeping for
// owned block k.
__device__ __forceinline__ void march_finalize_role(const FinalizeArgs& fin, MarchSmem& sm, int k, long long* fprof = nullptr) {
    (void)sm;
    finalize_block(fin, k, nullptr, fprof);
}

__device__ __forceinline__ int2 tile_from_list(const int2* list, int t) {
    int2 r;   // volatile: a reload after the row loop must not be folded into the load before it
    asm volatile("ld.global.nc.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(list + t));
    return r;
}

// ------------------------------------------------------------------ the launch-per-step kernel (large meshes)
// One warp strip per CTA; the ghost ring and the per-block dt come from the ghost kernel launched before it, the
// bookkeeping rides in the NEXT ghost launch.  Kept apart from march_tile above although the two run the same rows:
// this row loop sits at the edge of the register file (128 registers at 16 warps per SM) and ptxas only finds its
// 785-instruction, move-free schedule while nothing else shares the strip's code path - with ANY piece of the fused
// step compiled in (a flag wait, the perimeter epilogue, even as non-inlined calls) the same PTX loop becomes
// 846-865 instructions with 20-30 register moves, and the fused kernel measures 0.637 ms per launch against 0.576
// here on the benchmark mesh (same box, same sustained conditions; tools/ab_large.sh).  Large meshes are bound by
// this kernel, so they keep it; small meshes are bound by latency and take the fused step.
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


// ------------------------------------------------------------------ the fused step kernel
// One launch = one step: [stop test across ranks] + bookkeeping of the previous step + every strip of every owned
// block, the perimeter strips pushing the new state's ghost ring / edge lines / dt (tile_epilogue).  Single rank:
// grid (most strips of a block, [1 +] owned blocks), row 0 being the bookkeeping row when a step is pending; across
// ranks: grid (nfin + strips), strips in tile-list order - those that border another rank first, so that their puts
// travel while the interior strips run.
template <bool FAITHFUL, bool TPG, bool AXI>
__global__ void __launch_bounds__(MARCH_THREADS, FAITHFUL ? 1 : MARCH_MINB) march_fused_kernel(const StepArgs a, const Gas g, const PutArgs put, const FinalizeArgs fin) {
    __shared__ __align__(128) MarchSmem sm;
    StepVar sv;
    sv.cur = a.cur; sv.dt_table = a.dt_table; sv.dt_next = a.dt_next; sv.dt_used = a.dt_used; sv.partials = a.partials;
    sv.parity = put.parity; sv.seq = put.seq; sv.first = 1; sv.prof = nullptr;
    const StopArgs& sx = fin.stop_x;
    if ((PUSH_EXP & 1) && sx.enabled) {
        // across ranks: nothing of this launch runs before its stop decision (made by the first CTA)
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            if (threadIdx.x == 0) stop_test(sx, a.launch_no >= 3 ? sx.iter0 + a.launch_no - 3 : -1, const_cast<int*>(a.stop), a.flags, a.launch_no);
            __syncwarp();
        } else {
            stop_gate(sx.go, a.launch_no, a.stop, a.flags, sx.timeout_ns);
        }
    }
    if (*reinterpret_cast<volatile const int*>(a.stop)) return;
    if (a.tile_list != nullptr) {
        if ((PUSH_EXP & 1) && (int)blockIdx.x < a.nfin) { march_finalize_role(fin, sm, blockIdx.x); return; }
        const int t = (int)blockIdx.x - a.nfin;
        if (t >= a.tile_count) return;
        const int2 tl = tile_from_list(a.tile_list, t);
        march_tile<FAITHFUL, TPG, AXI, false, true>(a, g, put, sv, sm, tl.x, tl.y, a.blocks + tl.x, [&](int& k, int& strip, const BlockDev*& bk) {
            const int2 r = tile_from_list(a.tile_list, (int)blockIdx.x - a.nfin);
            k = r.x; strip = r.y; bk = a.blocks + r.x;
        });
    } else {
        const int fy = a.nfin ? 1 : 0;
        if ((int)blockIdx.y < fy) {
            if ((PUSH_EXP & 1) && (int)blockIdx.x < a.nfin) march_finalize_role(fin, sm, blockIdx.x);
            return;
        }
        const int k = (int)blockIdx.y - fy;
        if ((int)blockIdx.x >= a.blocks[k].tiles_x * a.blocks[k].tiles_y) return;    // CTA-uniform
        march_tile<FAITHFUL, TPG, AXI, false, true>(a, g, put, sv, sm, k, blockIdx.x, a.blocks + k, [&](int& k2, int& strip2, const BlockDev*& bk) {
            k2 = (int)blockIdx.y - (a.nfin ? 1 : 0); strip2 = blockIdx.x; bk = a.blocks + k2;
        });
    }
}

// ------------------------------------------------------------------ persistent launch (small meshes)
// Grid barrier of a cooperative launch: every CTA is one warp.  Bounded (FLAG_COMM), like every wait here.
//   bar[0]                      arrival counter (+1 per CTA and step; bit 31 = "the batch has converged", raised by a
//                               bookkeeping CTA before it arrives)
//   bar[32 * (1 + r)], r < 16   release words, one 128-byte line each: the number of the last completed step, with the
//                               counter's bit 31 carried over
// The last CTA to arrive (it knows from the value its own atomic returned) publishes the step number in the release
// words; everybody else polls ONE of them, chosen by CTA index: the late arrivals' atomics do not queue behind ~1100
// warps re-reading the counter's line, and the stop word of the batch travels in the same word, so that no CTA needs
// another round trip to L2 to learn that the next step may start.  (Measured on the shipped cases: the poll rate and
// where the pollers look are neutral; what costs is every FENCE, ~1000 cycles each.)
// Returns 0 on a time-out, 1 to go on, 2 when the batch has converged.
#ifndef BAR_RELEASE_ATOM
#define BAR_RELEASE_ATOM 1   // 1: the arrival is ONE atom.release.gpu (SASS: MEMBAR.ALL.GPU + ATOMG); 0: fence.acq_rel.gpu + atomicAdd, whose
                             //    fence also invalidates the SM's L1 (CCTL.IVALL) under the strips that still work there: 7.8 -> 7.5 us per step
#endif
#ifndef BAR_FORMAL
#define BAR_FORMAL 0    // 1: every CTA leaves the barrier through an acquire fence and the last arrival fences before the release,
                        //    as the PTX memory model asks for (A/B builds: +2000 cycles on every step of the shipped cases)
#endif
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bar_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// `acquire`: the caller reads memory other CTAs wrote through L1 after the barrier (the bookkeeping CTAs: plain loads
// of the iteration counters and res0) and needs the acquire fence that invalidates it.  A STRIP does not: everything it
// reads of other CTAs' writes it reads at L2 - its rows by TMA (behind a proxy fence), the dt table and the cells of the
// perimeter epilogue by ld.cg / ld.volatile - and what it reads through L1 (ghost-cell records, strip lists, gas
// constants) never changes; its loads are issued after the flag load they depend on (a branch), and every arrival's
// writes were performed at L2 before its arrival (release fence + atomic at L2).  A fence costs ~1000 cycles here, a
// whole step of a shipped case ~14000: the strips leave on the relaxed load (measured 8.76 -> 7.8 us per step).
__device__ __forceinline__ int grid_barrier(unsigned* bar, unsigned step_no, unsigned nctas, bool raise_stop, bool acquire, long long* bprof = nullptr) {
    unsigned word = 0;
    const long long tb = bprof ? clock64() : 0;    // diagnostics (tools/diag_persist.py); null in production
    __syncwarp();
    if (threadIdx.x == 0) {
        if (raise_stop) atomicOr(bar, 0x80000000u);
        // release: what this CTA wrote is performed before its arrival
#if BAR_RELEASE_ATOM
        unsigned old;
        asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
#else
        bar_fence();
        const unsigned old = atomicAdd(bar, 1u);
#endif
        if (bprof) bprof[2] = clock64() - tb;
        if ((old & 0x7fffffffu) == step_no * nctas - 1u) {
            word = step_no | (old & 0x80000000u);
            if (BAR_FORMAL || acquire) bar_fence();
#pragma unroll
            for (int r = 0; r < BAR_RELEASE_WORDS; r++) *reinterpret_cast<volatile unsigned*>(bar + 32 * (1 + r)) = word;
        } else {
            const unsigned* rel = bar + 32 * (1 + (int)(blockIdx.x % BAR_RELEASE_WORDS));
            word = ld_relaxed_gpu(rel);
            if ((word & 0x7fffffffu) < step_no) {
                const unsigned long long t0 = global_ns();
                unsigned spins = 0;
                while (((word = ld_relaxed_gpu(rel)) & 0x7fffffffu) < step_no) {
                    if ((++spins & 63u) == 0 && (long long)(global_ns() - t0) > 4000000000ll) { word = 0; break; }
                }
            }
            if (BAR_FORMAL || acquire) bar_fence();
        }
    }
    word = __shfl_sync(0xffffffffu, word, 0);
    return word == 0 ? 0 : (word & 0x80000000u) ? 2 : 1;
}

// A batch of a.coop_steps steps of a mesh whose strips are all resident at once (the shipped cases hold ~50 k
// cells: a step is a few microseconds of work and would otherwise be bound by launch latency, main.cpp:263-532
// as ONE launch).  CTA i < nowned: the bookkeeping of block i, one step behind the strips; the other CTAs keep
// one strip each for the whole batch.  Between steps: one grid barrier - the perimeter strips have pushed the
// new state's ghost ring and dt by then.  Same strips, same arithmetic, same summation order as the launch-per-step
// path: the results are bit-identical.
// (12 resident warps per SM suffice for the meshes that fit this launch at all; 168 registers instead of 128 spare
//  the strip code its spills)
template <bool FAITHFUL, bool TPG, bool AXI>
__global__ void __launch_bounds__(MARCH_THREADS, FAITHFUL ? 1 : 12) march_persist_kernel(const StepArgs a, const Gas g, const FinalizeArgs fin,
                                                                                                double* dt_used_alt, double* partials_alt, const DeferArgs df) {
    __shared__ __align__(128) MarchSmem sm;
    // the strip's block descriptor, copied once: after every grid barrier (an acquire) a plain global load goes to L2
    // again, and the descriptor's fields stand at the head of every step's chain (strip position -> first TMA copy)
    __shared__ __align__(16) BlockDev sblk;
    if (*a.stop) return;
    // df.on (no stop test in this batch): no bookkeeping CTAs - the strips keep every step's sums, dt and error bits, the
    // history records are made after the launch (launch_finalize_deferred).  On the shipped cases the bookkeeping CTA of
    // a block was the longest CTA of every step (one warp, code that runs once per step out of a cold instruction cache)
    const int nfin = df.on ? 0 : a.nowned;
    const bool fin_role = (int)blockIdx.x < nfin;
    int k = 0, strip = 0;
    if (!fin_role) {
        const int2 tl = __ldg(a.tile_list + ((int)blockIdx.x - nfin));
        k = tl.x; strip = tl.y;
        static_assert(sizeof(BlockDev) % 8 == 0, "BlockDev is copied in 8-byte words");
        const unsigned long long* from = reinterpret_cast<const unsigned long long*>(a.blocks + k);
        unsigned long long* to = reinterpret_cast<unsigned long long*>(&sblk);
        for (int w = threadIdx.x; w < (int)(sizeof(BlockDev) / 8); w += 32) to[w] = from[w];
        __syncwarp();
    }
    PutArgs put;
    put.peer_mode = 0; put.parity = 0; put.seq = 0; put.nranks = 1; put.rank = 0; put.expected = 0;
    put.dt_dst = nullptr; put.flag_dst = nullptr; put.ticket = nullptr; put.local_flags = nullptr; put.dt_stride = 0; put.timeout_ns = 0; put.dg = 0;
    int done = 0;
    bool stopped = false;
    // a.prof (diagnostics, tools/diag_persist.py): per step, SM clock cycles of the first strip's march, of its wait
    // at the grid barrier, and of the first bookkeeping CTA
    const bool prof = a.prof != nullptr && threadIdx.x == 0 && ((int)blockIdx.x == nfin || blockIdx.x == 0);
    // ... and, for step PROF_ALL_STEP alone, the same two numbers of EVERY CTA (upper half of the buffer)
    constexpr int PROF_STEPS = 2048, PROF_ALL_STEP = 500;
    const bool prof_all = a.prof != nullptr && threadIdx.x == 0 && 2 * gridDim.x <= 12 * PROF_STEPS;
    for (int s = 0; s < a.coop_steps; s++) {
        const int odd = s & 1;
        const long long t0 = (prof || prof_all) ? clock64() : 0;
        if (fin_role) {
            if (s > 0) {
                FinalizeArgs f = fin;              // step s-1
                f.partials = odd ? fin.partials : partials_alt;
                f.dt_used = odd ? fin.dt_used : dt_used_alt;
                // diagnostics: phases of the bookkeeping of blocks 0 and 1, 8 slots per step behind the per-CTA records
                long long* fp = (a.prof != nullptr && threadIdx.x == 0 && blockIdx.x < 2 && s < 1024) ? a.prof + 12 * PROF_STEPS + 8192 + 16 * s + 8 * blockIdx.x : nullptr;
                march_finalize_role(f, sm, blockIdx.x, fp);
            }
        } else {
            StepVar sv;
            sv.cur = a.cur ^ odd;
            sv.dt_table = odd ? a.dt_next : a.dt_table;
            sv.dt_next = odd ? const_cast<double*>(a.dt_table) : a.dt_next;
            sv.dt_used = df.on ? df.dt + s : odd ? dt_used_alt : a.dt_used;
            sv.partials = df.on ? df.partials + s * df.stride : odd ? partials_alt : a.partials;
            // error bits as of the end of the previous step, for its history record (per-step status)
            if (df.on && s > 0 && blockIdx.x == 0 && threadIdx.x == 0) df.flags[s - 1] = __ldcg(a.flags);
            sv.parity = 0; sv.seq = 0; sv.first = s == 0;
            sv.prof = (prof && s < PROF_STEPS) ? a.prof + 12 * s + 3 : nullptr;
            march_tile<FAITHFUL, TPG, AXI, false, true>(a, g, put, sv, sm, k, strip, &sblk, [&](int& k2, int& strip2, const BlockDev*& bk) {
                const int2 r = tile_from_list(a.tile_list, (int)blockIdx.x - (df.on ? 0 : a.nowned));
                k2 = r.x; strip2 = r.y; bk = &sblk;
            });
        }
        done = s + 1;
        const long long t1 = (prof || prof_all) ? clock64() : 0;
        if (prof && s < PROF_STEPS) a.prof[12 * s + (fin_role ? 2 : 0)] = t1 - t0;
        // a bookkeeping CTA that finds the stop word set (by itself or by a CTA before it) says so on its way into the barrier
        const bool raise = fin_role && s > 0 && *reinterpret_cast<volatile const int*>(a.stop) != 0;
        const int go = grid_barrier(a.grid_bar, (unsigned)(s + 1), gridDim.x, raise, fin_role, (prof && !fin_role && s < PROF_STEPS) ? a.prof + 12 * s + 9 : nullptr);
        if (go == 0) { if (threadIdx.x == 0) atomicOr(a.flags, FLAG_COMM); break; }
        if (prof && !fin_role && s < PROF_STEPS) a.prof[12 * s + 1] = clock64() - t1;
        if (prof_all && s == PROF_ALL_STEP) { a.prof[12 * PROF_STEPS + 2 * blockIdx.x] = t1 - t0; a.prof[12 * PROF_STEPS + 2 * blockIdx.x + 1] = clock64() - t1; }
        if (go == 2) { stopped = true; break; }
    }
    // the last step's bookkeeping
    if (fin_role && !stopped && done > 0) {
        const int odd = (done - 1) & 1;
        FinalizeArgs f = fin;
        f.partials = odd ? partials_alt : fin.partials;
        f.dt_used = odd ? dt_used_alt : fin.dt_used;
        march_finalize_role(f, sm, blockIdx.x);
    }
}

}  // namespace par2d
