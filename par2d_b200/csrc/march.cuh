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

constexpr int MARCH_WARPS = 4;                    // independent strips per CTA
constexpr int MARCH_THREADS = 32 * MARCH_WARPS;

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

template <bool FAITHFUL, bool TPG, bool AXI>
__global__ void __launch_bounds__(MARCH_THREADS, FAITHFUL ? 1 : 3) march_kernel(const StepArgs a, const Gas g) {
    if (*a.stop) return;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.y;                      // grid.y = owned blocks, grid.x covers the block with most strips
    const BlockDev& b = a.blocks[k];
    const int strip = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (strip >= b.tiles_x * b.tiles_y) return;    // warp-uniform
    const int ncx = b.ncx, ncy = b.ncy, pitch = b.pitch;
    const long long plane = b.plane;
    const int c0 = (strip % b.tiles_x) * MCELLS;   // first updated cell column
    const int j0 = (strip / b.tiles_x) * a.march_rows;
    const int j1 = min(j0 + a.march_rows, ncy);
    const int c = c0 - 1 + lane;                   // this lane's cell column (-1 and ncx are the ghost columns)
    const bool writer = lane >= 1 && lane <= MCELLS && c < ncx;
    const double* __restrict__ Sin = b.state[a.cur];
    double* __restrict__ Sout = b.state[a.cur ^ 1];
    const double* __restrict__ X = b.nodes;
    const double* __restrict__ Y = b.nodes + plane;

    // global dt = max over blocks (main.cpp:270-292); the table was written by the ghost kernel
    double dt = 0.0;
    for (int q = lane; q < a.nblocks_global; q += 32) dt = fmax(dt, a.dt_table[q]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) dt = fmax(dt, __shfl_xor_sync(0xffffffffu, dt, s));

    // lanes right of the right ghost column re-read that column (never used, never stored)
    const long long col = COL0 + min(c, ncx);
    auto load_row = [&](int arow) {   // array row: cell row j lives in arow j+1, node row j in arow j+1
        RawRow r;
        const long long o = (long long)arow * pitch + col;
        r.rho = Sin[o]; r.u = Sin[plane + o]; r.v = Sin[2 * plane + o]; r.T = Sin[3 * plane + o];
        r.x = X[o]; r.y = Y[o];
        return r;
    };

    unsigned flags = 0;
    // ---- prologue: cell rows j0-1 (below) and j0, node row j0, bottom-face flux of row j0
    RawRow below = load_row(j0), rc = load_row(j0 + 1), rn = load_row(j0 + 2);
    unsigned fl0 = 0;
    FaceState cur = face_state<FAITHFUL, TPG>(g, rc.rho, rc.u, rc.v, rc.T, fl0);
    double curT = rc.T;
    double x0 = rc.x, y0 = rc.y;                               // node (c, j)
    double x1 = shfl_down1(x0), y1 = shfl_down1(y0);           // node (c+1, j)
    Flux Gb;
    {
        unsigned flb = 0;
        const FaceState lo = face_state<FAITHFUL, TPG>(g, below.rho, below.u, below.v, below.T, flb);
        // j-face: left state = upper cell, right state = lower cell (SpatialDiscretization.cpp:556-625)
        Gb = ldfss<FAITHFUL, AXI>(face_geom<FAITHFUL>(x0, y0, x1, y1), cur, lo);
        if (writer) flags |= flb | fl0;
    }
    double sq0 = 0.0, sq1 = 0.0, sq2 = 0.0, sq3 = 0.0;

#pragma unroll 1
    for (int j = j0; j < j1; j++) {
        // rn = raw row j+1 (cell row above, node row j+1); prefetch the row after it
        const RawRow ru = rn;
        if (j + 1 < j1) rn = load_row(j + 3);
        unsigned flu = 0;
        const FaceState up = face_state<FAITHFUL, TPG>(g, ru.rho, ru.u, ru.v, ru.T, flu);
        const double x3 = ru.x, y3 = ru.y;                      // node (c, j+1)
        const double x2 = shfl_down1(x3), y2 = shfl_down1(y3);  // node (c+1, j+1)

        // i-face between cell (c-1, j) [left state] and (c, j): node (c, j) -> (c, j+1)
        const Flux Fl = ldfss<FAITHFUL, AXI>(face_geom<FAITHFUL>(x0, y0, x3, y3), shfl_up1(cur), cur);
        // j-face between cell (c, j+1) [left state = upper cell] and (c, j): node (c, j+1) -> (c+1, j+1)
        const Flux Gt = ldfss<FAITHFUL, AXI>(face_geom<FAITHFUL>(x3, y3, x2, y2), up, cur);
        const Flux Fr = shfl_down1(Fl);

        if (writer) {
            flags |= flu;
            // cell (c, j): nodes 1..4 = (c,j), (c+1,j), (c+1,j+1), (c,j+1)  (MeshModule.cpp:28-40)
            const double vol = cell_volume(x0, y0, x1, y1, x2, y2, x3, y3);
            const double ycc = 0.25 * (y0 + y1 + y2 + y3);
            double r[4];
            if constexpr (FAITHFUL) {
                // accumulation order of SpatialDiscretization.cpp:481-625 (interior faces: left, right,
                // bottom, top) then :643-977 (block-boundary faces: left, right, bottom, top)
                const double yc = AXI ? ycc : 0.0;
                const bool bl = c == 0, br = c == ncx - 1, bb = j == 0, bt = j == ncy - 1;
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
            } else {
                r[0] = (Fl.f0 - Fr.f0) + (Gt.f0 - Gb.f0);
                r[1] = (Fl.f1 - Fr.f1) + (Gt.f1 - Gb.f1);
                r[2] = (Fl.f2 - Fr.f2) + (Gt.f2 - Gb.f2);
                r[3] = (Fl.f3 - Fr.f3) + (Gt.f3 - Gb.f3);
                if constexpr (AXI) r[2] = fma(ycc, (Fl.parr - Fr.parr) + (Gt.parr - Gb.parr), r[2]);
                const double ivol = fast_rcp(vol);
#pragma unroll
                for (int v = 0; v < 4; v++) r[v] *= ivol;
            }
            sq0 = fma(r[0], r[0], sq0); sq1 = fma(r[1], r[1], sq1); sq2 = fma(r[2], r[2], sq2); sq3 = fma(r[3], r[3], sq3);

            if (a.res_out != nullptr && a.res_block == k) {
                double* o = a.res_out + 4 * ((long long)j * ncx + c);
                o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
            }
            if (a.update) {
                const double T = curT;
                unsigned fl2 = 0;
                // e0 = h - Rs T + v2/2 (StateVariables.h:62) with h0 = h + v2/2 already at hand
                double e0, Cv, inv_cv = g.inv_cv_cpg;
                if constexpr (FAITHFUL) {
                    const double v2 = cur.u * cur.u + cur.v * cur.v;
                    e0 = enthalpy<TPG>(g, T, fl2) - g.Rs * T + 0.5 * v2;
                } else {
                    e0 = fma(-g.Rs, T, cur.h0);
                }
                Cv = heat_capacity<TPG>(g, T, fl2) - g.Rs;
                if constexpr (TPG && !FAITHFUL) inv_cv = fast_rcp(Cv);
                double dv[4];
                solve_cell<FAITHFUL, AXI>(dt, ycc, cur.rho, cur.u, cur.v, e0, Cv, inv_cv, r, dv, flags);
                const double nrho = cur.rho + dv[0], nu = cur.u + dv[1], nv = cur.v + dv[2];
                const double nT = fmax(T + dv[3], 201.0);
                if (!(nrho > 0.0)) flags |= FLAG_DENSITY;
                const long long o = (long long)(j + 1) * pitch + COL0 + c;
                Sout[o] = nrho; Sout[plane + o] = nu; Sout[2 * plane + o] = nv; Sout[3 * plane + o] = nT;
            }
        }
        // rotate: the row above becomes the current row
        cur = up; curT = ru.T;
        x0 = x3; y0 = y3; x1 = x2; y1 = y2;
        Gb = Gt;
    }

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
