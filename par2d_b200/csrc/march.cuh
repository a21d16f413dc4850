// march.cuh -- the production step kernel: register-marching warp strips.
//
// One warp advances a strip of MCELLS (30) cell columns by `rows` cell rows of one block, marching
// upwards row by row (FV/SpatialDiscretization.cpp:481-1011 + main.cpp:39-57,307-351,394-404 fused):
//
//   lane l holds cell column c0-1+l: lanes 0 and 31 are the strip's left/right halo columns (their
//   state is loaded and derived, but only lanes 1..30 update a cell), so the 32 lanes of a row are one
//   fully coalesced 256-byte read per plane, straight from the padded HBM planes (ghost ring in place).
//   Per row and lane:  load + derive the cell one row up (software-prefetched one further row ahead),
//   ONE i-face flux (left face; the left cell's state comes from lane l-1 by warp shuffle, the right
//   face's flux from lane l+1), ONE j-face flux (top face; the bottom face is last row's top face,
//   carried in registers), residual, norms, Jacobian solve, update, 4 coalesced stores.
//   Nothing is staged in shared memory: every state value is read from HBM once per strip (plus the
//   halo column / row re-reads, 2/30 and 1/rows, which hit L2) and every face flux is evaluated once
//   per strip.  The sum of squares is accumulated per lane over the rows and reduced by warp shuffles
//   in a fixed order once per strip.
//
// Compared with the TMA-staged 32x8 tile kernel (kernels.cuh::step_kernel, kept as the A/B baseline,
// PAR2D_KERNEL=tile): no block-wide barriers, no smem round trips (~70 LDS/STS per cell there), a
// fifth of the index arithmetic, and global-load latency hidden by the row prefetch instead of by
// occupancy.  See DESIGN.md for the ncu numbers of both.
#pragma once
#include "common.cuh"
#include "physics.cuh"

namespace par2d {

#ifndef MARCH_NWARPS
#define MARCH_NWARPS 1
#endif
constexpr int MARCH_WARPS = MARCH_NWARPS;                    // one strip per CTA: every loop bound derives from blockIdx, so ptxas proves the shuffles convergent
constexpr int MARCH_THREADS = 32 * MARCH_WARPS;
#ifndef MARCH_UNROLL2
#define MARCH_UNROLL2 0
#endif
#ifndef MARCH_MINB
#define MARCH_MINB 12
#endif

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ FaceState shfl_up1(const FaceState& s) {
    FaceState o;
    o.rho = shfl_up1(s.rho); o.u = shfl_up1(s.u); o.v = shfl_up1(s.v);
    o.p = shfl_up1(s.p); o.a = shfl_up1(s.a); o.h0 = shfl_up1(s.h0); o.ip = shfl_up1(s.ip);
    return o;
}
__device__ __forceinline__ Flux shfl_down1(const Flux& f) {
    Flux o;
    o.f0 = shfl_down1(f.f0); o.f1 = shfl_down1(f.f1); o.f2 = shfl_down1(f.f2); o.f3 = shfl_down1(f.f3);
    o.parr = shfl_down1(f.parr);
    return o;
}

struct RawRow {   // one cell row as loaded: state of the cell and the node at its lower-left corner
    double rho, u, v, T, x, y;
};

// What a lane carries from one row to the next: the derived state of its cell, the two lower nodes of
// the cell (own column and, from lane+1, the next column) and the flux through the cell's bottom face.
struct RowState {
    FaceState s;
    double T;
    double xa, ya, xb, yb;   // node (c, j), node (c+1, j)
    Flux G;                  // j-face flux at the bottom of the row
};

struct MarchCtx {
    int k, c, ncx, ncy, pitch;
    long long plane;
    bool writer;
    double dt;
    double* Sout;
};

// One row of a strip: consumes the carried state `lo` of cell row j and the raw row above it, produces
// the carried state `hi` of row j+1, updates cell (c, j).  Called alternately with (A,B) and (B,A) so
// that no register is moved between rows.
template <bool FAITHFUL, bool TPG, bool AXI>
__device__ __forceinline__ void march_row(const StepArgs& a, const Gas& g, const MarchCtx& m, int j, const RowState& lo, const RawRow& ru,
                                          RowState& hi, double& sq0, double& sq1, double& sq2, double& sq3, unsigned& flags) {
    unsigned flu = 0;
    hi.s = face_state<FAITHFUL, TPG>(g, ru.rho, ru.u, ru.v, ru.T, flu);
    hi.T = ru.T;
    hi.xa = ru.x; hi.ya = ru.y;                                  // node (c, j+1)
    hi.xb = shfl_down1(ru.x); hi.yb = shfl_down1(ru.y);          // node (c+1, j+1)

    // i-face between cell (c-1, j) [left state] and (c, j): node (c, j) -> (c, j+1)
    const Flux Fl = ldfss<FAITHFUL, AXI>(face_geom<FAITHFUL>(lo.xa, lo.ya, hi.xa, hi.ya), shfl_up1(lo.s), lo.s);
    // j-face between cell (c, j+1) [left state = upper cell] and (c, j): node (c, j+1) -> (c+1, j+1)
    hi.G = ldfss<FAITHFUL, AXI>(face_geom<FAITHFUL>(hi.xa, hi.ya, hi.xb, hi.yb), hi.s, lo.s);
    const Flux Fr = shfl_down1(Fl);
    const Flux& Gb = lo.G;
    const Flux& Gt = hi.G;

    if (m.writer) {
        flags |= flu;
        // cell (c, j): nodes 1..4 = (c,j), (c+1,j), (c+1,j+1), (c,j+1)  (MeshModule.cpp:28-40)
        const double vol = cell_volume(lo.xa, lo.ya, lo.xb, lo.yb, hi.xb, hi.yb, hi.xa, hi.ya);
        const double ycc = 0.25 * (lo.ya + lo.yb + hi.yb + hi.ya);
        const FaceState& cur = lo.s;
        double r[4];        // residual = calc_dudt's dudt
        double dv[4];
        if constexpr (FAITHFUL) {
            // accumulation order of SpatialDiscretization.cpp:481-625 (interior faces: left, right,
            // bottom, top) then :643-977 (block-boundary faces: left, right, bottom, top)
            const double yc = AXI ? ycc : 0.0;
            const bool bl = m.c == 0, br = m.c == m.ncx - 1, bb = j == 0, bt = j == m.ncy - 1;
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
            if (a.update) {
                unsigned fl2 = 0;
                const double v2 = cur.u * cur.u + cur.v * cur.v;
                const double e0 = enthalpy<TPG>(g, lo.T, fl2) - g.Rs * lo.T + 0.5 * v2;   // StateVariables.h:62
                const double Cv = heat_capacity<TPG>(g, lo.T, fl2) - g.Rs;
                solve_cell<true, AXI>(m.dt, ycc, cur.rho, cur.u, cur.v, e0, Cv, 0.0, r, dv, flags);
            }
        } else {
            double R[4];    // raw flux sums
            R[0] = (Fl.f0 - Fr.f0) + (Gt.f0 - Gb.f0);
            R[1] = (Fl.f1 - Fr.f1) + (Gt.f1 - Gb.f1);
            R[2] = (Fl.f2 - Fr.f2) + (Gt.f2 - Gb.f2);
            R[3] = (Fl.f3 - Fr.f3) + (Gt.f3 - Gb.f3);
            double ivol, sc;
            if constexpr (AXI) {
                // axisymmetric source: the y-momentum pressure term carries the cell's own radius
                // (SpatialDiscretization.cpp:521-523,595-597); 1/vol and 1/ycc (main.cpp:325-331) share
                // one reciprocal
                R[2] = fma(ycc, (Fl.parr - Fr.parr) + (Gt.parr - Gb.parr), R[2]);
                const double im = fast_rcp(vol * ycc);
                ivol = im * ycc;
                sc = m.dt * im;
            } else {
                ivol = fast_rcp(vol);
                sc = m.dt * ivol;
            }
#pragma unroll
            for (int v = 0; v < 4; v++) r[v] = R[v] * ivol;
            if (a.update) {
                unsigned fl2 = 0;
                const double e0 = fma(-g.Rs, lo.T, cur.h0);   // h - Rs T + v2/2 with h0 = h + v2/2 at hand
                double inv_cv = g.inv_cv_cpg;
                if constexpr (TPG) inv_cv = fast_rcp(heat_capacity<TPG>(g, lo.T, fl2) - g.Rs);
                solve_lean(sc, cur.rho, cur.u, cur.v, e0, inv_cv, R, dv);
            }
        }
        sq0 = fma(r[0], r[0], sq0); sq1 = fma(r[1], r[1], sq1); sq2 = fma(r[2], r[2], sq2); sq3 = fma(r[3], r[3], sq3);

        if (a.res_out != nullptr && a.res_block == m.k) {
            double* o = a.res_out + 4 * ((long long)j * m.ncx + m.c);
            o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
        }
        if (a.update) {
            // main.cpp:394-404: unk += dv, rho > 0, UpdateState's T clamp
            const double nrho = cur.rho + dv[0], nu = cur.u + dv[1], nv = cur.v + dv[2];
            const double nT = fmax(lo.T + dv[3], 201.0);
            if (!(nrho > 0.0)) flags |= FLAG_DENSITY;
            double* o = m.Sout + ((long long)(j + 1) * m.pitch + COL0 + m.c);
            __stcg(o, nrho); __stcg(o + m.plane, nu); __stcg(o + 2 * m.plane, nv); __stcg(o + 3 * m.plane, nT);
        }
    }
}

template <bool FAITHFUL, bool TPG, bool AXI>
__global__ void __launch_bounds__(MARCH_THREADS, FAITHFUL ? 1 : MARCH_MINB) march_kernel(const StepArgs a, const Gas g) {
    if (*a.stop) return;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.y;                      // grid.y = owned blocks, grid.x covers the block with most strips
    const BlockDev& b = a.blocks[k];
    const int strip = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (strip >= b.tiles_x * b.tiles_y) return;    // warp-uniform
    MarchCtx m;
    m.k = k; m.ncx = b.ncx; m.ncy = b.ncy; m.pitch = b.pitch; m.plane = b.plane;
    const int c0 = (strip % b.tiles_x) * MCELLS;   // first updated cell column
    const int j0 = (strip / b.tiles_x) * a.march_rows;
    const int j1 = min(j0 + a.march_rows, m.ncy);
    m.c = c0 - 1 + lane;                           // this lane's cell column (-1 and ncx are the ghost columns)
    m.writer = lane >= 1 && lane <= MCELLS && m.c < m.ncx;
    m.Sout = b.state[a.cur ^ 1];
    const double* __restrict__ Sin = b.state[a.cur];
    const double* __restrict__ X = b.nodes;
    const long long plane = m.plane;

    // global dt = max over blocks (main.cpp:270-292); the table was written by the ghost kernel
    double dt = 0.0;
    for (int q = lane; q < a.nblocks_global; q += 32) dt = fmax(dt, a.dt_table[q]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) dt = fmax(dt, __shfl_xor_sync(0xffffffffu, dt, s));
    m.dt = dt;

    // lanes right of the right ghost column re-read that column (never used, never stored)
    const double* src = Sin + (COL0 + min(m.c, m.ncx));
    const double* nsrc = X + (COL0 + min(m.c, m.ncx));
    auto load_row = [&](int arow) {   // array row: cell row j and node row j both live in array row j+1
        RawRow r;
        const long long o = (long long)arow * m.pitch;
        r.rho = __ldg(src + o); r.u = __ldg(src + plane + o); r.v = __ldg(src + 2 * plane + o); r.T = __ldg(src + 3 * plane + o);
        r.x = __ldg(nsrc + o); r.y = __ldg(nsrc + plane + o);
        return r;
    };

    unsigned flags = 0;
    // ---- prologue: cell rows j0-1 (below) and j0, node row j0, bottom-face flux of row j0
    RowState A, B;
    RawRow ra, rb;
    {
        const RawRow below = load_row(j0), rc = load_row(j0 + 1);
        ra = load_row(j0 + 2);
        unsigned fl0 = 0;
        A.s = face_state<FAITHFUL, TPG>(g, rc.rho, rc.u, rc.v, rc.T, fl0);
        A.T = rc.T;
        A.xa = rc.x; A.ya = rc.y;
        A.xb = shfl_down1(rc.x); A.yb = shfl_down1(rc.y);
        const FaceState lo = face_state<FAITHFUL, TPG>(g, below.rho, below.u, below.v, below.T, fl0);
        // j-face: left state = upper cell, right state = lower cell (SpatialDiscretization.cpp:556-625)
        A.G = ldfss<FAITHFUL, AXI>(face_geom<FAITHFUL>(A.xa, A.ya, A.xb, A.yb), A.s, lo);
        if (m.writer) flags |= fl0;
    }
    double sq0 = 0.0, sq1 = 0.0, sq2 = 0.0, sq3 = 0.0;

#if MARCH_UNROLL2
    // ---- two rows per trip, (A -> B) then (B -> A); raw rows prefetched one row ahead of their use
    int j = j0;
#pragma unroll 1
    for (; j + 1 < j1; j += 2) {
        rb = load_row(j + 3);
        march_row<FAITHFUL, TPG, AXI>(a, g, m, j, A, ra, B, sq0, sq1, sq2, sq3, flags);
        if (j + 2 < j1) ra = load_row(j + 4);
        march_row<FAITHFUL, TPG, AXI>(a, g, m, j + 1, B, rb, A, sq0, sq1, sq2, sq3, flags);
    }
    if (j < j1) march_row<FAITHFUL, TPG, AXI>(a, g, m, j, A, ra, B, sq0, sq1, sq2, sq3, flags);
#else
    // ---- one row per trip; the raw row above is prefetched one row ahead of its use
#pragma unroll 1
    for (int j = j0; j < j1; j++) {
        rb = ra;
        if (j + 1 < j1) ra = load_row(j + 3);
        march_row<FAITHFUL, TPG, AXI>(a, g, m, j, A, rb, B, sq0, sq1, sq2, sq3, flags);
        A = B;
    }
#endif

    // ---- sum of squares of this strip, fixed order: lanes (xor tree)
    double sq[4] = {sq0, sq1, sq2, sq3};
#pragma unroll
    for (int v = 0; v < 4; v++) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sq[v] += __shfl_xor_sync(0xffffffffu, sq[v], s);
    }
    if (isnan(sq[0]) || isnan(sq[1]) || isnan(sq[2]) || isnan(sq[3])) flags |= FLAG_NAN;   // SpatialDiscretization.cpp:1008
    if (lane < 4) {
        const double s = lane == 0 ? sq[0] : lane == 1 ? sq[1] : lane == 2 ? sq[2] : sq[3];
        a.partials[4 * (long long)(b.tile0 + strip) + lane] = s;
    }
    if (flags) atomicOr(a.flags, flags);
}

}  // namespace par2d
