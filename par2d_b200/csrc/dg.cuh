// dg.cuh -- the second-order branch of the path (cfg.accur == 1): the reference's DG(P1) discretisation
// (FV/DGP1Tools.cpp, the ACCUR == 1 paths of FV/SpatialDiscretization.cpp:159-1050 and FV/main.cpp:300-463).
//
// Unknowns per cell: the mean (rho,u,v,T) in the padded state planes (common.cuh) plus xi / eta slopes in
// SoA planes ux/uy[k*ncell + j*ncx+i] (the reference's AoS (j*ncx+i)*4+k is transposed at the ABI, so that a warp's
// lanes read consecutive doubles); two Gauss points per face with LDFSS at each; the slopes' right-hand
// sides carry the xi/eta moments of the face fluxes plus a volume integral; the slopes are damped (0.9) and
// limited against the neighbouring means after every update.
//
// Three kernels per step:
//   dg_ghost_kernel : two ghost states per boundary face (DGP1_ghost_cell_generator, blended normals),
//                     or the same-GPU neighbour's reconstruction at its Gauss points (the reference's halo)
//   dg_cell_kernel  : marching warp strips - every interior face flux (two Gauss points) evaluated once per strip -
//                     then per cell the four faces accumulated in the reference's order (so in PAR2D_MODE_FAITHFUL
//                     every cell's right-hand sides are bit-identical to the reference's scatter loops), boundary
//                     + volume integrals, /vol, norms, three solves with one Jacobian, mean update into the other
//                     state buffer, damped slopes into the temporary arrays
//   dg_limit_kernel : the reference's in-place, sequential slope limiter restated per cell: it reads the OLD
//                     mean of the cells after it in memory order and the NEW mean of the cells before it
//                     (both state buffers are at hand), including the row wrap of its flat-index tests
// PAR2D_MODE_FAITHFUL keeps the reference's operation order and divisions; PAR2D_MODE_FAST uses the lean
// LDFSS / state / solve of physics.cuh on interior faces (boundary faces, whose blended normals are not unit
// vectors, keep the general form).  Off-rank neighbours: pack_kernel (kernels.cuh) sends the reconstruction at the two
// Gauss points of every edge face, 8 doubles per edge cell.
#pragma once
#include "common.cuh"
#include "physics.cuh"

namespace par2d {

constexpr double DG_DAMP = 0.9;       // main.cpp:92
constexpr double DG_DUSCALE = 0.5;    // main.cpp:93
constexpr int DG_THREADS = 128;
#ifndef DG_MINB
#define DG_MINB 6      // CTAs per SM the cell kernel is compiled for (register budget 65536 / (32 * DG_WARPS * DG_MINB))
#endif

struct DgArgs {
    double* partials;    // [dg strips][12]: sum of squares of dudt, duxdt, duydt per strip of the cell kernel
    int max_tiles;       // strips of the largest owned block (the cell kernel runs DG_WARPS strips per CTA)
    int max_lim_tiles;   // grid.x of the limiter kernel: ceil(cells of the largest block / DG_THREADS)
    int max_points;      // 2 x longest perimeter (grid.x of the ghost kernel, in threads)
    double* res_out;     // optional dump of dudt | duxdt | duydt (3 x ncell x 4) of block res_block
    int res_block;
    int update;
};

struct DgCell {
    double u[4], ux[4], uy[4], p;
};

__device__ __forceinline__ double dg_gauss() { return 1.0 / sqrt(3.0); }   // folded at compile time, as 1.0/sqrt(3.0) is in the reference

// pull the mean and the slopes of cell (i,j) towards the SM ahead of their use (the strips of the cell kernel run at
// ~8 warps per SM: nothing else hides the latency of a row's first touch)
__device__ __forceinline__ void dg_prefetch_cell(const BlockDev& b, int buf, int i, int j) {
    const double* S = b.state[buf] + (long long)(j + 1) * b.pitch + COL0 + i;
    const long long e = (long long)j * b.ncx + i, nc = (long long)b.ncx * b.ncy;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S + k * b.plane));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(b.ux + k * nc + e));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(b.uy + k * nc + e));
    }
}

// mean + slopes + pressure of cell (i,j) (State::p of ElemVar, StateVariables.h:49-50)
__device__ __forceinline__ void dg_load_cell(const Gas& g, const BlockDev& b, int buf, int i, int j, DgCell& c) {
    const double* S = b.state[buf];
    const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
    const long long e = (long long)j * b.ncx + i, nc = (long long)b.ncx * b.ncy;
#pragma unroll
    for (int k = 0; k < 4; k++) { c.u[k] = S[k * b.plane + o]; c.ux[k] = b.ux[k * nc + e]; c.uy[k] = b.uy[k * nc + e]; }
    c.p = fmax(c.u[0] * g.Rs * c.u[3], 1.0);
}

// get_u_val, DGP1Tools.cpp:10-37: the density is extrapolated through the pressure
template <bool FAITHFUL = true>
__device__ __forceinline__ void dg_u_val(const Gas& g, const DgCell& c, double xsi, double eta, double out[4]) {
    const double dpdxsi = c.ux[0] * g.Rs * c.u[3] + c.u[0] * g.Rs * c.ux[3];
    const double dpdeta = c.uy[0] * g.Rs * c.u[3] + c.u[0] * g.Rs * c.uy[3];
    const double pout = c.p + (xsi * dpdxsi) + (eta * dpdeta);
    if constexpr (FAITHFUL) out[0] = pout / (g.Rs * (c.u[3] + (xsi * c.ux[3]) + (eta * c.uy[3])));
    else out[0] = pout * fast_rcp(g.Rs * (c.u[3] + (xsi * c.ux[3]) + (eta * c.uy[3])));
#pragma unroll
    for (int k = 1; k < 4; k++) out[k] = c.u[k] + (xsi * c.ux[k]) + (eta * c.uy[k]);
}

// the reconstructed state at a Gauss point, UpdateState'd (T clamp, StateVariables.h:38-68)
// FAITHFUL = false: the lean state of physics.cuh (a/2 carried, series sqrt), for the lean LDFSS of interior faces
template <bool FAITHFUL, bool TPG>
__device__ __forceinline__ FaceState dg_point_state(const Gas& g, const DgCell& c, double xsi, double eta, unsigned& flags) {
    double u[4];
    dg_u_val<FAITHFUL>(g, c, xsi, eta, u);
    if constexpr (FAITHFUL) u[3] = fmax(u[3], 201.0);
    else u[3] = max_sel(u[3], 201.0);
    return face_state<FAITHFUL, TPG>(g, u[0], u[1], u[2], u[3], flags);
}

struct DgAcc {
    double r[4], rx[4], ry[4];
};

// the accumulation statements of DGP1Tools.cpp:168-184 (and twins) for ONE of the two cells of a face
template <bool AXI>
__device__ __forceinline__ void dg_accumulate(DgAcc& a, const Flux& F, double yc, bool subtract, double xsi, double eta) {
    const double weight = 0.5;
    double f[4] = {F.f0, F.f1, F.f2, F.f3};
    if constexpr (AXI) f[2] = (F.f2 + (yc * F.parr));
    else f[2] = (F.f2 + (0.0 * F.parr));
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (subtract) {
            a.r[k] -= weight * f[k];
            a.rx[k] -= weight * f[k] * xsi;
            a.ry[k] -= weight * f[k] * eta;
        } else {
            a.r[k] += weight * f[k];
            a.rx[k] += weight * f[k] * xsi;
            a.ry[k] += weight * f[k] * eta;
        }
    }
}

// DGP1_xsi_face_integral (dir 0) / DGP1_eta_face_integral (dir 1), the part that lands in one cell: `me` is the
// face's left cell when mine_is_left, `nb` the cell across.  One body for the four faces of a cell (the caller
// loops over them without unrolling: the kernel is instruction-cache and register bound otherwise).
template <bool FAITHFUL, bool TPG, bool AXI>
__device__ __forceinline__ void dg_face(const Gas& g, const FaceGeom& fg, const DgCell& me, const DgCell& nb, int dir, bool mine_is_left,
                                        double yc, DgAcc& a, unsigned& flags) {
    const double G = dg_gauss();
#pragma unroll 1
    for (int pt = 0; pt < 2; pt++) {
        const double s = pt == 0 ? G : -G;
        double xsiL, etaL, xsiR, etaR;
        if (dir == 0) { xsiL = 1.0; xsiR = -1.0; etaL = etaR = s; }
        else { etaL = -1.0; etaR = 1.0; xsiL = xsiR = s; }
        const double xm = mine_is_left ? xsiL : xsiR, em = mine_is_left ? etaL : etaR;
        const double xo = mine_is_left ? xsiR : xsiL, eo = mine_is_left ? etaR : etaL;
        const FaceState sm = dg_point_state<FAITHFUL, TPG>(g, me, xm, em, flags);
        const FaceState so = dg_point_state<FAITHFUL, TPG>(g, nb, xo, eo, flags);
        FaceState sl, sr;
        sl.rho = mine_is_left ? sm.rho : so.rho; sr.rho = mine_is_left ? so.rho : sm.rho;
        sl.u = mine_is_left ? sm.u : so.u;       sr.u = mine_is_left ? so.u : sm.u;
        sl.v = mine_is_left ? sm.v : so.v;       sr.v = mine_is_left ? so.v : sm.v;
        sl.p = mine_is_left ? sm.p : so.p;       sr.p = mine_is_left ? so.p : sm.p;
        sl.a = mine_is_left ? sm.a : so.a;       sr.a = mine_is_left ? so.a : sm.a;
        sl.h0 = mine_is_left ? sm.h0 : so.h0;    sr.h0 = mine_is_left ? so.h0 : sm.h0;
        const Flux F = ldfss<FAITHFUL, AXI>(fg, sl, sr);
        if (isnan(F.f0 + F.f1 + F.f2 + F.f3)) flags |= FLAG_NAN;
        dg_accumulate<AXI>(a, F, yc, mine_is_left, xm, em);
    }
}

// unit normal + length + mid-point radius of the face node (i0,j0) -> (i1,j1) (MeshModule.cpp:52-85)
__device__ __forceinline__ FaceGeom dg_geom(const BlockDev& b, int i0, int j0, int i1, int j1) {
    const double* X = b.nodes;
    const double* Y = b.nodes + b.plane;
    const long long p0 = (long long)(j0 + 1) * b.pitch + COL0 + i0, p1 = (long long)(j1 + 1) * b.pitch + COL0 + i1;
    return face_geom<true>(X[p0], Y[p0], X[p1], Y[p1]);
}

// DGP1_boundary_face_integral, DGP1Tools.cpp:317-531: type 1 bottom, 2 right, 3 top, 4 left
template <bool TPG, bool AXI>
__device__ __forceinline__ void dg_boundary_face(const Gas& g, const BlockDev& b, int type, int side, int n, const DgCell& in, const FaceGeom& fg,
                                                 double nLx, double nLy, double nRx, double nRy, double yc, DgAcc& a, unsigned& flags) {
    const double G = dg_gauss();
    const double* gh = b.ghost2[side];
#pragma unroll 1
    for (int pt = 0; pt < 2; pt++) {
        const double s = pt == 0 ? G : -G;
        double xsi, eta;
        if (type == 1) { xsi = s; eta = -1.0; }
        else if (type == 2) { xsi = 1.0; eta = s; }
        else if (type == 3) { xsi = s; eta = 1.0; }
        else { xsi = -1.0; eta = s; }
        const double temp = fabs(s);
        FaceGeom f2 = fg;
        f2.nx = (1.0 - 0.5 * temp) * fg.nx + 0.5 * temp * (pt == 0 ? nRx : nLx);
        f2.ny = (1.0 - 0.5 * temp) * fg.ny + 0.5 * temp * (pt == 0 ? nRy : nLy);
        const double* e = gh + 4 * (2 * (long long)n + pt);
        const FaceState ex = face_state<true, TPG>(g, e[0], e[1], e[2], e[3], flags);   // ghosts are stored clamped
        const FaceState si = dg_point_state<true, TPG>(g, in, xsi, eta, flags);   // blended, non-unit normals: the general form in both modes
        Flux F;
        if (type == 1 || type == 2) F = ldfss<true, AXI>(f2, si, ex);
        else F = ldfss<true, AXI>(f2, ex, si);
        if (isnan(F.f0 + F.f1 + F.f2 + F.f3)) flags |= FLAG_NAN;
        // rhs += sideMult * weight * f: sideMult = -1 on the bottom and right faces
        const double weight = 0.5, sm = (type == 1 || type == 2) ? -1.0 : 1.0;
        double f[4] = {F.f0, F.f1, F.f2, F.f3};
        if constexpr (AXI) f[2] = (F.f2 + (yc * F.parr));
        else f[2] = (F.f2 + (0.0 * F.parr));
#pragma unroll
        for (int k = 0; k < 4; k++) {
            a.r[k] += sm * weight * f[k];
            a.rx[k] += sm * weight * f[k] * xsi;
            a.ry[k] += sm * weight * f[k] * eta;
        }
    }
}

// bounded wait for the peers' sequence flags (same protocol as kernels.cuh::wait_peers, which is defined after this header)
__device__ __forceinline__ bool dg_wait_peers(const unsigned long long* flags, int nranks, int rank, unsigned long long need, long long timeout_ns) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int r = 0; r < nranks; r++) {
        if (r == rank) continue;
        for (;;) {
            unsigned long long v, now;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + 16 * r) : "memory");
            if (v >= need) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if ((long long)(now - t0) > timeout_ns) return false;
            __nanosleep(200);
        }
    }
    return true;
}

// ------------------------------------------------------------------ ghost points
template <bool FAITHFUL>
__global__ void __launch_bounds__(256) dg_ghost_kernel(StepArgs a, Gas g, DgArgs d, int write_dt) {
    if (*a.stop) return;
    if (a.wait_flags != nullptr) {
        // peer-mapped mode: wait for the neighbours' lines and dts of this state (see ghost_kernel)
        __shared__ int arrived;
        if (threadIdx.x == 0) {
            arrived = dg_wait_peers(a.wait_flags, a.nranks, a.rank, a.wait_seq, a.timeout_ns) ? 1 : 0;
            if (!arrived) atomicOr(a.flags, FLAG_COMM);
        }
        __syncthreads();
    }
    const BlockDev& b = a.blocks[blockIdx.y];
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ncx = b.ncx, ncy = b.ncy;
    if (t == 0 && write_dt) {   // find_dt of this block from cell (0,0), main.cpp:19-37 (multi-rank: the pack kernel's job)
        const double* S = b.state[a.cur];
        const long long o = (long long)b.pitch + COL0;
        const double u = S[b.plane + o], v = S[2 * b.plane + o], T = S[3 * b.plane + o];
        const double vmax = sqrt(u * u + v * v) + sqrt(g.gamRs * T);
        const double dt = g.cfl * b.min_face / vmax;
        if (isnan(dt)) atomicOr(a.flags, FLAG_NAN);
        a.dt_table[b.gid] = dt;
    }
    if (t >= 4 * (ncx + ncy)) return;
    const int pt = t & 1;
    int n = t >> 1, side;
    if (n < ncx) side = SIDE_BOTTOM;
    else if ((n -= ncx) < ncy) side = SIDE_RIGHT;
    else if ((n -= ncy) < ncx) side = SIDE_TOP;
    else { n -= ncx; side = SIDE_LEFT; }
    const double G = dg_gauss();
    const double s = pt == 0 ? G : -G;
    double* out = b.ghost2[side] + 4 * (2 * (long long)n + pt);
    unsigned fl = 0;
    if (b.nb_kind[side] == NB_REMOTE) {
        // the line an off-rank neighbour packed (pack_kernel): 8 doubles per edge cell, point 1 then point 2
        const double* line = b.recv[side] + a.recv_off + 8 * (long long)n + 4 * pt;
        out[0] = __ldcg(line); out[1] = __ldcg(line + 1); out[2] = __ldcg(line + 2); out[3] = fmax(__ldcg(line + 3), 201.0);
        return;
    }
    if (b.nb_kind[side] == NB_LOCAL) {
        // halo of the second-order branch (SpatialDiscretization.cpp:243-256 and twins): the neighbour's
        // reconstruction at the Gauss points of ITS edge face
        const BlockDev& o = a.blocks[b.nb_block[side]];
        const int os = b.nb_side[side];
        int i, j;
        double xsi, eta;
        if (os == SIDE_LEFT) { i = 0; j = n; xsi = -1.0; eta = s; }
        else if (os == SIDE_RIGHT) { i = o.ncx - 1; j = n; xsi = 1.0; eta = s; }
        else if (os == SIDE_BOTTOM) { i = n; j = 0; xsi = s; eta = -1.0; }
        else { i = n; j = o.ncy - 1; xsi = s; eta = 1.0; }
        DgCell c;
        dg_load_cell(g, o, a.cur, i, j, c);
        double u[4];
        dg_u_val(g, c, xsi, eta, u);
        out[0] = u[0]; out[1] = u[1]; out[2] = u[2]; out[3] = fmax(u[3], 201.0);
        return;
    }
    // DGP1_ghost_cell_generator, DGP1Tools.cpp:548-784
    int i, j, code, i0, j0, i1, j1, ir0, jr0, ir1, jr1, il0, jl0, il1, jl1;
    double sg, xsi, eta;
    const int last = (side == SIDE_BOTTOM || side == SIDE_TOP) ? ncx - 1 : ncy - 1;
    const int nr = n != last ? n + 1 : n, nl = n != 0 ? n - 1 : n;
    if (side == SIDE_BOTTOM) {
        i = n; j = 0; code = b.ibound[n]; sg = 1.0; xsi = s; eta = -1.0;
        i0 = n; j0 = 0; i1 = n + 1; j1 = 0; ir0 = nr; jr0 = 0; ir1 = nr + 1; jr1 = 0; il0 = nl; jl0 = 0; il1 = nl + 1; jl1 = 0;
    } else if (side == SIDE_RIGHT) {   // column nx-1 here (the first-order generator reads column 0)
        i = ncx - 1; j = n; code = b.ibound[ncx + n]; sg = 1.0; xsi = 1.0; eta = s;
        i0 = ncx; j0 = n; i1 = ncx; j1 = n + 1; ir0 = ncx; jr0 = nr; ir1 = ncx; jr1 = nr + 1; il0 = ncx; jl0 = nl; il1 = ncx; jl1 = nl + 1;
    } else if (side == SIDE_TOP) {
        i = n; j = ncy - 1; code = b.ibound[2 * ncx + ncy - 1 - n]; sg = -1.0; xsi = s; eta = 1.0;
        i0 = n; j0 = ncy; i1 = n + 1; j1 = ncy; ir0 = nr; jr0 = ncy; ir1 = nr + 1; jr1 = ncy; il0 = nl; jl0 = ncy; il1 = nl + 1; jl1 = ncy;
    } else {
        i = 0; j = n; code = b.ibound[2 * ncx + 2 * ncy - 1 - n]; sg = -1.0; xsi = -1.0; eta = s;
        i0 = 0; j0 = n; i1 = 0; j1 = n + 1; ir0 = 0; jr0 = nr; ir1 = 0; jr1 = nr + 1; il0 = 0; jl0 = nl; il1 = 0; jl1 = nl + 1;
    }
    const FaceGeom f0 = dg_geom(b, i0, j0, i1, j1);
    const FaceGeom fo = pt == 0 ? dg_geom(b, ir0, jr0, ir1, jr1) : dg_geom(b, il0, jl0, il1, jl1);
    double normx = sg < 0 ? -f0.nx : f0.nx, normy = sg < 0 ? -f0.ny : f0.ny;
    const double ox = sg < 0 ? -fo.nx : fo.nx, oy = sg < 0 ? -fo.ny : fo.ny;
    if (pt == 0) {
        normx = normx * (1.0 - 0.5 * s) + 0.5 * s * ox;
        normy = normy * (1.0 - 0.5 * s) + 0.5 * s * oy;
    } else {
        normx = normx * (1.0 + 0.5 * s) - 0.5 * s * ox;
        normy = normy * (1.0 + 0.5 * s) - 0.5 * s * oy;
    }
    // get_boundary_point, DGP1Tools.cpp:533-546
    DgCell c;
    dg_load_cell(g, b, a.cur, i, j, c);
    double u[4];
    dg_u_val(g, c, xsi, eta, u);
    u[3] = fmax(u[3], 201.0);
    const double v2 = u[1] * u[1] + u[2] * u[2];
    const double p = fmax(u[0] * g.Rs * u[3], 1.0);
    const double aint = sqrt(g.gamRs * u[3]);
    double gh[4] = {out[0], out[1], out[2], out[3]};   // stays as is when the BC leaves it unwritten
    boundary_state(g, code, normx, normy, u, aint, p, v2, gh);
    out[0] = gh[0]; out[1] = gh[1]; out[2] = gh[2]; out[3] = gh[3];
    (void)fl;
}

// ------------------------------------------------------------------ residual + solve + update of the means
// One WARP marches a strip of DG_SW = 31 cell columns (lane 31 is a halo lane: it only evaluates the xi face
// between column 30 of the strip and the next column) upwards over DG_ROWS rows.  Per row a lane evaluates ONE
// xi face (its left one; the right one is the left face of lane+1, fetched by warp shuffles) and ONE eta face
// (its top one; the bottom one is the previous row's top face, carried in registers) - every interior face flux
// once per strip instead of once per adjacent cell - and then accumulates its four faces in the reference's
// order, so that PAR2D_MODE_FAITHFUL stays bit-identical to the reference's scatter loops.
constexpr int DG_SW = 31;
#ifndef DG_PARK
#define DG_PARK 1
#endif
#ifndef DG_PREFETCH
#define DG_PREFETCH 1
#endif
#ifndef DG_ROWS
#define DG_ROWS 32
#endif
#ifndef DG_WARPS
#define DG_WARPS 2          // independent strips per CTA (no shared memory, no barrier)
#endif

struct DgFluxPair {
    Flux F[2];              // the two Gauss points of a face
};

// the two Gauss-point fluxes of the interior face with left cell L and right cell R (DGP1Tools.cpp:150-222)
template <bool FAITHFUL, bool TPG, bool AXI>
__device__ __forceinline__ void dg_face_flux(const Gas& g, const FaceGeom& fg, const DgCell& L, const DgCell& R, int dir, DgFluxPair& out, unsigned& flags) {
    const double G = dg_gauss();
#pragma unroll 1
    for (int pt = 0; pt < 2; pt++) {
        const double s = pt == 0 ? G : -G;
        double xsiL, etaL, xsiR, etaR;
        if (dir == 0) { xsiL = 1.0; xsiR = -1.0; etaL = etaR = s; }
        else { etaL = -1.0; etaR = 1.0; xsiL = xsiR = s; }
        const FaceState sl = dg_point_state<FAITHFUL, TPG>(g, L, xsiL, etaL, flags);
        const FaceState sr = dg_point_state<FAITHFUL, TPG>(g, R, xsiR, etaR, flags);
        const Flux F = ldfss<FAITHFUL, AXI>(fg, sl, sr);
        if (isnan(F.f0 + F.f1 + F.f2 + F.f3)) flags |= FLAG_NAN;
        if (pt == 0) out.F[0] = F; else out.F[1] = F;
    }
}

__device__ __forceinline__ DgFluxPair dg_shfl_down1(const DgFluxPair& p) {
    DgFluxPair o;
#pragma unroll
    for (int pt = 0; pt < 2; pt++) {
        o.F[pt].f0 = __shfl_down_sync(0xffffffffu, p.F[pt].f0, 1); o.F[pt].f1 = __shfl_down_sync(0xffffffffu, p.F[pt].f1, 1);
        o.F[pt].f2 = __shfl_down_sync(0xffffffffu, p.F[pt].f2, 1); o.F[pt].f3 = __shfl_down_sync(0xffffffffu, p.F[pt].f3, 1);
        o.F[pt].parr = __shfl_down_sync(0xffffffffu, p.F[pt].parr, 1);
    }
    return o;
}

template <bool FAITHFUL, bool TPG, bool AXI>
__global__ void __launch_bounds__(32 * DG_WARPS, DG_MINB) dg_cell_kernel(StepArgs a, Gas g, DgArgs d) {
    if (*a.stop) return;
    const int k = blockIdx.y;
    const BlockDev& b = a.blocks[k];
    const int ncx = b.ncx, ncy = b.ncy;
    const int lane = threadIdx.x & 31;
    const int strip = (int)blockIdx.x * DG_WARPS + (threadIdx.x >> 5);
    if (strip >= b.dg_ntiles) return;              // warp-uniform
    const int strips_x = (ncx + DG_SW - 1) / DG_SW;
    const int i = (strip % strips_x) * DG_SW + lane;
    const int j0 = (strip / strips_x) * DG_ROWS, j1 = min(j0 + DG_ROWS, ncy);
    const bool has = i < ncx;                      // a real cell column (the halo lane included)
    const bool cell = has && lane < DG_SW;         // this lane updates the cells of its column
    const long long ncell = (long long)ncx * ncy;
    const double* X = b.nodes;
    const double* Y = b.nodes + b.plane;
    const double G = dg_gauss();
    unsigned flags = 0;
#if DG_PARK
    // the 12 running sums of squares and the carried bottom-face flux pair live in shared memory between rows:
    // 22 doubles per lane that would otherwise sit in registers through the whole row
    __shared__ double park[DG_WARPS][22][32];
    double (*pk)[32] = park[threadIdx.x >> 5];
#pragma unroll
    for (int q = 0; q < 12; q++) pk[q][lane] = 0.0;
#else
    double sq[12];
#pragma unroll
    for (int q = 0; q < 12; q++) sq[q] = 0.0;
#endif

    // global dt = max over blocks (main.cpp:270-292)
    double dt = 0.0;
    for (int t = 0; t < a.nblocks_global; t++) dt = fmax(dt, a.dt_table[t]);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.dt_used = dt;

    // eta face j0 (between row j0, the left = upper cell, and row j0-1): the first row's bottom face
    DgFluxPair Fb = {};
    if (cell && j0 >= 1) {
        DgCell up, lo;
        dg_load_cell(g, b, a.cur, i, j0, up);
        dg_load_cell(g, b, a.cur, i, j0 - 1, lo);
        const long long n00 = (long long)(j0 + 1) * b.pitch + COL0 + i;
        dg_face_flux<FAITHFUL, TPG, AXI>(g, face_geom<FAITHFUL>(X[n00], Y[n00], X[n00 + 1], Y[n00 + 1]), up, lo, 1, Fb, flags);
    }
#if DG_PARK
#define DG_PARK_PUT(P) { _Pragma("unroll") for (int pt = 0; pt < 2; pt++) { pk[12 + 5 * pt][lane] = (P).F[pt].f0; pk[13 + 5 * pt][lane] = (P).F[pt].f1; pk[14 + 5 * pt][lane] = (P).F[pt].f2; pk[15 + 5 * pt][lane] = (P).F[pt].f3; pk[16 + 5 * pt][lane] = (P).F[pt].parr; } }
#define DG_PARK_GET(P) { _Pragma("unroll") for (int pt = 0; pt < 2; pt++) { (P).F[pt].f0 = pk[12 + 5 * pt][lane]; (P).F[pt].f1 = pk[13 + 5 * pt][lane]; (P).F[pt].f2 = pk[14 + 5 * pt][lane]; (P).F[pt].f3 = pk[15 + 5 * pt][lane]; (P).F[pt].parr = pk[16 + 5 * pt][lane]; } }
    DG_PARK_PUT(Fb)
#endif
#pragma unroll 1
    for (int j = j0; j < j1; j++) {
        DgCell me;
        DgFluxPair Fl = {}, Fr;
        const long long n00 = (long long)(j + 1) * b.pitch + COL0 + i;
        double xa = 0.0, ya = 0.0, xd = 0.0, yd = 0.0;
        if (has) {
#if DG_PREFETCH
            if (j + 2 < ncy) dg_prefetch_cell(b, a.cur, i, j + 2);   // the row after next: its first touch is row j+1's top face
#endif
            dg_load_cell(g, b, a.cur, i, j, me);
            xa = X[n00]; ya = Y[n00]; xd = X[n00 + b.pitch]; yd = Y[n00 + b.pitch];
            if (i >= 1) {          // xi face i: node (i,j) -> (i,j+1), left cell (i-1,j), right cell (i,j)
                DgCell le;
                dg_load_cell(g, b, a.cur, i - 1, j, le);
                dg_face_flux<FAITHFUL, TPG, AXI>(g, face_geom<FAITHFUL>(xa, ya, xd, yd), le, me, 0, Fl, flags);
            }
        }
        Fr = dg_shfl_down1(Fl);    // xi face i+1 = the left face of the next lane
        if (!cell) continue;       // the halo lane and lanes past the block only feed the shuffle
        const double xb = X[n00 + 1], yb = Y[n00 + 1], xc = X[n00 + b.pitch + 1], yc_ = Y[n00 + b.pitch + 1];
        {
            const long long e = (long long)j * ncx + i;
            const double ycc = 0.25 * (ya + yb + yc_ + yd);           // geoel[2], MeshModule.cpp:40
            const double vol = cell_volume(xa, ya, xb, yb, xc, yc_, xd, yd);
            const double ycs = AXI ? ycc : 0.0;
            DgAcc acc;
#pragma unroll
            for (int q = 0; q < 4; q++) acc.r[q] = acc.rx[q] = acc.ry[q] = 0.0;
            // the reference's accumulation order: xi face i (this cell is the right one), xi face i+1 (left)
            // (SpatialDiscretization.cpp:481-552), eta face j (left = upper), eta face j+1 (right) (:556-625);
            // per face point 1 then point 2.  Each flux pair is consumed as early as the order allows (registers)
#pragma unroll
            for (int pt = 0; pt < 2; pt++) if (i >= 1) dg_accumulate<AXI>(acc, Fl.F[pt], ycs, false, -1.0, pt == 0 ? G : -G);
#pragma unroll
            for (int pt = 0; pt < 2; pt++) if (i <= ncx - 2) dg_accumulate<AXI>(acc, Fr.F[pt], ycs, true, 1.0, pt == 0 ? G : -G);
#if DG_PARK
            DgFluxPair Fb;
            DG_PARK_GET(Fb)
#endif
#pragma unroll
            for (int pt = 0; pt < 2; pt++) if (j >= 1) dg_accumulate<AXI>(acc, Fb.F[pt], ycs, true, pt == 0 ? G : -G, -1.0);
            if (j <= ncy - 2) {    // eta face j+1: node (i,j+1) -> (i+1,j+1), left = upper cell (i,j+1), right cell (i,j)
                DgCell ab;
                dg_load_cell(g, b, a.cur, i, j + 1, ab);
                dg_face_flux<FAITHFUL, TPG, AXI>(g, face_geom<FAITHFUL>(xd, yd, xc, yc_), ab, me, 1, Fb, flags);   // next row's bottom face
#pragma unroll
                for (int pt = 0; pt < 2; pt++) dg_accumulate<AXI>(acc, Fb.F[pt], ycs, false, pt == 0 ? G : -G, 1.0);
#if DG_PARK
                DG_PARK_PUT(Fb)
#endif
            }
            // boundary faces (:643-977), left, right, bottom, top.  fNormalL / fNormalR are declared once per row
            // (column) and shared by the left and right (bottom and top) face of it: at the ends of the right (top)
            // side they still hold the left (bottom) side's values
            if (i == 0 || i == ncx - 1 || j == 0 || j == ncy - 1) {
    #pragma unroll 1
                for (int sd = 0; sd < 4; sd++) {
                    int type, side, n;
                    bool exists;
                    FaceGeom f, fL, fR;
                    if (sd == 0) {
                        exists = i == 0; type = 4; side = SIDE_LEFT; n = j;
                        if (!exists) continue;
                        f = face_geom<true>(xa, ya, xd, yd);
                        fL = j != 0 ? dg_geom(b, 0, j - 1, 0, j) : f;
                        fR = j != ncy - 1 ? dg_geom(b, 0, j + 1, 0, j + 2) : f;
                    } else if (sd == 1) {
                        exists = i == ncx - 1; type = 2; side = SIDE_RIGHT; n = j;
                        if (!exists) continue;
                        f = face_geom<true>(xb, yb, xc, yc_);
                        fL = j != 0 ? dg_geom(b, ncx, j - 1, ncx, j) : dg_geom(b, 0, j, 0, j + 1);
                        fR = j != ncy - 1 ? dg_geom(b, ncx, j + 1, ncx, j + 2) : dg_geom(b, 0, j, 0, j + 1);
                    } else if (sd == 2) {
                        exists = j == 0; type = 1; side = SIDE_BOTTOM; n = i;
                        if (!exists) continue;
                        f = face_geom<true>(xa, ya, xb, yb);
                        fL = i != 0 ? dg_geom(b, i - 1, 0, i, 0) : f;
                        fR = i != ncx - 1 ? dg_geom(b, i + 1, 0, i + 2, 0) : f;
                    } else {
                        exists = j == ncy - 1; type = 3; side = SIDE_TOP; n = i;
                        if (!exists) continue;
                        f = face_geom<true>(xd, yd, xc, yc_);
                        fL = i != 0 ? dg_geom(b, i - 1, ncy, i, ncy) : dg_geom(b, i, 0, i + 1, 0);
                        fR = i != ncx - 1 ? dg_geom(b, i + 1, ncy, i + 2, ncy) : dg_geom(b, i, 0, i + 1, 0);
                    }
                    dg_boundary_face<TPG, AXI>(g, b, type, side, n, me, f, fL.nx, fL.ny, fR.nx, fR.ny, ycs, acc, flags);
                }
            }
            // :980-1011 division by the volume, then the volume integral of the slopes' equations (DGP1Tools.cpp:52-124)
            double r[4], rx[4], ry[4];
            unsigned fl2 = 0;
            const FaceState sme = face_state<true, TPG>(g, me.u[0], me.u[1], me.u[2], me.u[3], fl2);
            {
                const double ycv = AXI ? ycc : 1.0;
                const double rho = me.u[0];
                double Fx[4], Fy[4];
                Fx[0] = ycv * rho * me.u[1];
                Fx[1] = ycv * (rho * me.u[1] * me.u[1] + sme.p);
                Fx[2] = ycv * rho * me.u[1] * me.u[2];
                Fx[3] = ycv * rho * me.u[1] * sme.h0;
                Fy[0] = ycv * rho * me.u[2];
                Fy[1] = ycv * rho * me.u[1] * me.u[2];
                Fy[2] = ycv * (rho * me.u[2] * me.u[2] + sme.p);
                Fy[3] = ycv * rho * me.u[2] * sme.h0;
                const double xL = 0.5 * (xd + xa), xR = 0.5 * (xc + xb), xD = 0.5 * (xb + xa), xU = 0.5 * (xc + xd);
                const double yL = 0.5 * (yd + ya), yR = 0.5 * (yc_ + yb), yD = 0.5 * (yb + ya), yU = 0.5 * (yc_ + yd);
                const double nxs0 = -0.5 * (yU - yD), nxs1 = 0.5 * (xU - xD);
                const double nxs2 = -0.25 * ((xU - xD) * (yL - yR) - (yU - yD) * (xL - xR));
                double db1dx, db1dy, db2dx, db2dy, ivol = 0.0;
                if constexpr (FAITHFUL) { db1dx = nxs0 / nxs2; db1dy = nxs1 / nxs2; }
                else { const double q2 = fast_rcp(nxs2); db1dx = nxs0 * q2; db1dy = nxs1 * q2; ivol = fast_rcp(vol); }
                const double net0 = -0.5 * (yL - yR), net1 = 0.5 * (xL - xR);
                const double net2 = -0.25 * ((xL - xR) * (yD - yU) - (yL - yR) * (xD - xU));
                if constexpr (FAITHFUL) { db2dx = net0 / net2; db2dy = net1 / net2; }
                else { const double q2 = fast_rcp(net2); db2dx = net0 * q2; db2dy = net1 * q2; }
    #pragma unroll
                for (int q = 0; q < 4; q++) {
                    if constexpr (FAITHFUL) {
                        r[q] = acc.r[q] / vol;
                        rx[q] = 3.0 * acc.rx[q] / vol;
                        ry[q] = 3.0 * acc.ry[q] / vol;
                    } else {
                        r[q] = acc.r[q] * ivol;
                        rx[q] = 3.0 * acc.rx[q] * ivol;
                        ry[q] = 3.0 * acc.ry[q] * ivol;
                    }
                    rx[q] += 3.0 * (Fx[q] * db1dx + Fy[q] * db1dy);
                    ry[q] += 3.0 * (Fx[q] * db2dx + Fy[q] * db2dy);
                }
            }
            if (isnan(r[0])) flags |= FLAG_NAN;
    #pragma unroll
#if DG_PARK
            for (int q = 0; q < 4; q++) { pk[q][lane] += r[q] * r[q]; pk[4 + q][lane] += rx[q] * rx[q]; pk[8 + q][lane] += ry[q] * ry[q]; }
#else
            for (int q = 0; q < 4; q++) { sq[q] += r[q] * r[q]; sq[4 + q] += rx[q] * rx[q]; sq[8 + q] += ry[q] * ry[q]; }
#endif
            if (d.res_out != nullptr && d.res_block == k) {
                double* o = d.res_out + 4 * e;
    #pragma unroll
                for (int q = 0; q < 4; q++) { o[q] = r[q]; o[4 * ncell + q] = rx[q]; o[8 * ncell + q] = ry[q]; }
            }
            if (d.update) {
                // main.cpp:307-351: one Jacobian, three right-hand sides; :394-425 update of the mean, damped slopes
                const double v2 = me.u[1] * me.u[1] + me.u[2] * me.u[2];
                const double e0 = enthalpy<TPG>(g, me.u[3], fl2) - g.Rs * me.u[3] + 0.5 * v2;
                const double Cv = heat_capacity<TPG>(g, me.u[3], fl2) - g.Rs;
                double dv[4], dvx[4], dvy[4];
                if constexpr (FAITHFUL) {
                    solve_cell<true, AXI>(dt, ycc, me.u[0], me.u[1], me.u[2], e0, Cv, 0.0, r, dv, flags);
                    solve_cell<true, AXI>(dt, ycc, me.u[0], me.u[1], me.u[2], e0, Cv, 0.0, rx, dvx, flags);
                    solve_cell<true, AXI>(dt, ycc, me.u[0], me.u[1], me.u[2], e0, Cv, 0.0, ry, dvy, flags);
                } else {
                    // D is lower triangular: closed-form forward substitution, one reciprocal of rho for the three systems
                    const double sc = AXI ? dt * fast_rcp(ycc) : dt;
                    const double inv_cv = fast_rcp(Cv);
                    solve_lean(sc, me.u[0], me.u[1], me.u[2], e0, inv_cv, r, dv);
                    solve_lean(sc, me.u[0], me.u[1], me.u[2], e0, inv_cv, rx, dvx);
                    solve_lean(sc, me.u[0], me.u[1], me.u[2], e0, inv_cv, ry, dvy);
                }
                double* So = b.state[a.cur ^ 1];
                const long long o = (long long)(j + 1) * b.pitch + COL0 + i;
                double nu[4];
    #pragma unroll
                for (int q = 0; q < 4; q++) nu[q] = me.u[q] + dv[q];
                if (!(nu[0] > 0.0)) flags |= FLAG_DENSITY;
                nu[3] = fmax(nu[3], 201.0);
    #pragma unroll
                for (int q = 0; q < 4; q++) {
                    So[q * b.plane + o] = nu[q];
                    b.uxt[q * ((long long)ncx * b.ncy) + e] = (me.ux[q] + dvx[q]) * DG_DAMP;
                    b.uyt[q * ((long long)ncx * b.ncy) + e] = (me.uy[q] + dvy[q]) * DG_DAMP;
                }
            }
            flags |= fl2;
        }
    }
    // sums of squares of the strip: lanes (xor tree), one record per strip
#if DG_PARK
    double sq[12];
#pragma unroll
    for (int q = 0; q < 12; q++) sq[q] = pk[q][lane];
#endif
#pragma unroll
    for (int q = 0; q < 12; q++) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sq[q] += __shfl_xor_sync(0xffffffffu, sq[q], s);
    }
    if (lane < 12) {
        double s = sq[0];
#pragma unroll
        for (int q = 1; q < 12; q++) s = lane == q ? sq[q] : s;
        if (isnan(s)) flags |= FLAG_NAN;
        d.partials[12 * (long long)(b.dg_tile0 + strip) + lane] = s;
    }
    if (flags) atomicOr(a.flags, flags);
}

// ------------------------------------------------------------------ slope limiter (main.cpp:427-461)
__device__ __forceinline__ double dg_sign(double x) { return signbit(x) ? -1.0 : 1.0; }   // Indexing.h:22

template <bool FAITHFUL>
__global__ void __launch_bounds__(DG_THREADS) dg_limit_kernel(StepArgs a, DgArgs d) {
    if (*a.stop) return;
    const BlockDev& b = a.blocks[blockIdx.y];
    const int ncx = b.ncx;
    const long long ncell = (long long)ncx * b.ncy;
    const long long e = (long long)blockIdx.x * DG_THREADS + threadIdx.x;
    if (e >= ncell) return;
    // a.cur was the buffer read by the step: OLD means there, NEW means in the other one.  The reference updates
    // cell after cell in memory order, so a cell sees new means behind it (e-1, e-ncx) and old ones ahead
    const double* Sold = b.state[a.cur];
    const double* Snew = b.state[a.cur ^ 1];
    // flat neighbours e-1, e+1 (they wrap across rows at the ends of a row, as the reference's flat index does) and
    // e-ncx, e+ncx as offsets in the padded planes - 32-bit index arithmetic, one division per thread
    const int ei = (int)e, i = ei % ncx, j = ei / ncx;
    const long long om = (long long)(j + 1) * b.pitch + COL0 + i;
    const long long o_im = i > 0 ? om - 1 : om - b.pitch + (ncx - 1);
    const long long o_ip = i < ncx - 1 ? om + 1 : om + b.pitch - (ncx - 1);
    const long long o_jm = om - b.pitch, o_jp = om + b.pitch;
    const bool has_ip = e + 1 < ncell, has_im = e - 1 >= 0, has_jp = e + ncx < ncell, has_jm = e - ncx >= 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const double me = Snew[q * b.plane + om];
        double ux = b.uxt[q * ncell + e], uy = b.uyt[q * ncell + e];
        double du;
        if (has_ip && has_im) du = DG_DUSCALE * ((Sold[q * b.plane + o_ip] - Snew[q * b.plane + o_im]) / 2.0);
        else if (!has_im) du = DG_DUSCALE * ((Sold[q * b.plane + o_ip] - me) / 1.0);
        else du = DG_DUSCALE * ((me - Snew[q * b.plane + o_im]) / 1.0);
        ux = dg_sign(ux) * fmin(fabs(ux), fabs(du));
        if (has_jp && has_jm) du = DG_DUSCALE * ((Sold[q * b.plane + o_jp] - Snew[q * b.plane + o_jm]) / 2.0);
        else if (!has_jm) du = DG_DUSCALE * ((Sold[q * b.plane + o_jp] - me) / 1.0);
        else du = DG_DUSCALE * ((me - Snew[q * b.plane + o_jm]) / 1.0);
        uy = dg_sign(uy) * fmin(fabs(uy), fabs(du));
        b.ux[q * ncell + e] = ux;
        b.uy[q * ncell + e] = uy;
    }
}

}  // namespace par2d
