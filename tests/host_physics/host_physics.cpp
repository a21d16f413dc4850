// tests/host_physics/host_physics.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
// The point-wise device functions of par2d_b200/csrc/physics.cuh compiled for the HOST (g++, no CUDA),
// so that the formulas of both arithmetic modes can be checked against the oracle without a GPU.
// Built three times by the Makefile next to it: HP_MODE=0 (fast formulas, FMA contraction on, like nvcc's
// default), HP_MODE=1 (reference-ordered formulas, contraction off) and HP_MODE=2 (= 0 with PAR2D_LEAN_BC, the
// far-field ghost code of the fast mode's translation unit).  Nothing in par2d_b200/ loads it.
#include <cstdint>
#include <cmath>
#include "../../include/par2d_gpu.h"
#include "../../par2d_b200/csrc/physics.cuh"

using namespace par2d;

#ifndef HP_MODE
#define HP_MODE 0
#endif
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define HP(name) CAT(CAT(CAT(hp, HP_MODE), _), name)
constexpr bool FAITH = HP_MODE == 1;

static Gas make_gas(const par2d_config& c) {   // same as capi.cu::make_gas
    Gas g{};
    g.Rs = c.Rs; g.igam = c.igam; g.gam = std::fabs(c.igam); g.gamRs = g.gam * g.Rs; g.gamRs4 = 0.25 * g.gamRs;
    g.cpg_den = (1.0 - 1.0 / (c.igam)); g.cp_cpg = g.Rs / (1.0 - 1.0 / (c.igam)); g.inv_cv_cpg = 1.0 / (g.cp_cpg - g.Rs);
    for (int k = 0; k < 7; k++) { g.hi[k] = c.thermo_hi[k]; g.lo[k] = c.thermo_lo[k]; }
    for (int k = 0; k < 4; k++) g.uFS[k] = c.uFS[k];
    g.p_back = c.p_back; g.cfl = c.cfl;
    return g;
}

template <bool TPG, bool AXI>
static void ldfss_n(const Gas& g, int64_t n, const double* face, const double* uL, const double* uR, double* flux, double* parr) {
    for (int64_t i = 0; i < n; i++) {
        unsigned fl = 0;
        const double nx = face[4 * i], ny = face[4 * i + 1], len = face[4 * i + 2], yf = face[4 * i + 3];
        FaceGeom fg;
        fg.nx = nx; fg.ny = ny; fg.len = len; fg.yf = yf; fg.dx = -ny * len; fg.dy = nx * len; fg.rl = 0.0;
        const FaceState L = face_state<FAITH, TPG>(g, uL[4 * i], uL[4 * i + 1], uL[4 * i + 2], fmax(uL[4 * i + 3], 201.0), fl);
        const FaceState R = face_state<FAITH, TPG>(g, uR[4 * i], uR[4 * i + 1], uR[4 * i + 2], fmax(uR[4 * i + 3], 201.0), fl);
        const Flux f = ldfss<FAITH, AXI>(fg, L, R);
        flux[4 * i] = f.f0; flux[4 * i + 1] = f.f1; flux[4 * i + 2] = f.f2; flux[4 * i + 3] = f.f3;
        parr[i] = f.parr;
    }
}

template <bool TPG, bool AXI>
static void solve_n(const Gas& g, int64_t n, const double* dt_ycc, const double* unk, const double* res, double* dv) {
    for (int64_t i = 0; i < n; i++) {
        unsigned fl = 0;
        const double rho = unk[4 * i], u = unk[4 * i + 1], v = unk[4 * i + 2], T = fmax(unk[4 * i + 3], 201.0);
        const double v2 = u * u + v * v;
        const double h = enthalpy<TPG>(g, T, fl);
        const double Cv = heat_capacity<TPG>(g, T, fl) - g.Rs;
        const double e0 = h - g.Rs * T + 0.5 * v2;
        double inv_cv = g.inv_cv_cpg;
        if (TPG && !FAITH) inv_cv = fast_rcp(Cv);
        const double r[4] = {res[4 * i], res[4 * i + 1], res[4 * i + 2], res[4 * i + 3]};
        double d[4];
        solve_cell<FAITH, AXI>(dt_ycc[2 * i], dt_ycc[2 * i + 1], rho, u, v, e0, Cv, inv_cv, r, d, fl);
        dv[4 * i] = d[0]; dv[4 * i + 1] = d[1]; dv[4 * i + 2] = d[2]; dv[4 * i + 3] = d[3];
    }
}

extern "C" {

int HP(test_ldfss)(const par2d_config* cfg, int64_t n, const double* face, const double* uL, const double* uR, double* flux, double* parr) {
    const Gas g = make_gas(*cfg);
    const bool tpg = !(cfg->igam > 0), axi = cfg->iaxi == 1;
    if (tpg) { if (axi) ldfss_n<true, true>(g, n, face, uL, uR, flux, parr); else ldfss_n<true, false>(g, n, face, uL, uR, flux, parr); }
    else { if (axi) ldfss_n<false, true>(g, n, face, uL, uR, flux, parr); else ldfss_n<false, false>(g, n, face, uL, uR, flux, parr); }
    return 0;
}

int HP(test_boundary_state)(const par2d_config* cfg, int64_t n, const int32_t* btype, const double* norm, const double* uInt, double* uGhost) {
    const Gas g = make_gas(*cfg);
    for (int64_t i = 0; i < n; i++) {
        double ui[4] = {uInt[4 * i], uInt[4 * i + 1], uInt[4 * i + 2], fmax(uInt[4 * i + 3], 201.0)};
        const double v2 = ui[1] * ui[1] + ui[2] * ui[2];
        const double p = fmax(ui[0] * g.Rs * ui[3], 1.0);
#if defined(PAR2D_LEAN_BC)
        const double aint = fast_sqrt(g.gamRs * ui[3]);   // as bc_ghost_general does in that translation unit (halo.cuh)
#else
        const double aint = sqrt(g.gamRs * ui[3]);
#endif
        boundary_state(g, btype[i], norm[2 * i], norm[2 * i + 1], ui, aint, p, v2, uGhost + 4 * i);
    }
    return 0;
}

int HP(test_update_state)(const par2d_config* cfg, int64_t n, double* unk, double* derived) {
    const Gas g = make_gas(*cfg);
    const bool tpg = !(cfg->igam > 0);
    for (int64_t i = 0; i < n; i++) {
        unsigned fl = 0;
        const double T = fmax(unk[4 * i + 3], 201.0);
        unk[4 * i + 3] = T;
        const double rho = unk[4 * i], u = unk[4 * i + 1], v = unk[4 * i + 2];
        const FaceState s = tpg ? face_state<FAITH, true>(g, rho, u, v, T, fl) : face_state<FAITH, false>(g, rho, u, v, T, fl);
        const double v2 = u * u + v * v;
        const double h = tpg ? enthalpy<true>(g, T, fl) : enthalpy<false>(g, T, fl);
        const double Cp = tpg ? heat_capacity<true>(g, T, fl) : heat_capacity<false>(g, T, fl);
        double* o = derived + 8 * i;
        o[0] = s.p; o[1] = FAITH ? s.a : 2.0 * s.a; o[2] = h; o[3] = s.h0; o[4] = h - g.Rs * T + 0.5 * v2; o[5] = v2; o[6] = Cp; o[7] = Cp - g.Rs;
    }
    return 0;
}

int HP(test_solve)(const par2d_config* cfg, int64_t n, const double* dt_ycc, const double* unk, const double* res, double* dv) {
    const Gas g = make_gas(*cfg);
    const bool tpg = !(cfg->igam > 0), axi = cfg->iaxi == 1;
    if (tpg) { if (axi) solve_n<true, true>(g, n, dt_ycc, unk, res, dv); else solve_n<true, false>(g, n, dt_ycc, unk, res, dv); }
    else { if (axi) solve_n<false, true>(g, n, dt_ycc, unk, res, dv); else solve_n<false, false>(g, n, dt_ycc, unk, res, dv); }
    return 0;
}

}  // extern "C"

#if HP_MODE == 0
// The two forms of the lean LDFSS the device chooses between by warp vote (physics.cuh: ldfss_early): per face the test
// `up`, the fully-upwind flux (ldfss_upwind_left, evaluated whether or not the test holds) and the general flux computed
// from the same early quantities (ldfss_pre_from_early + ldfss_lean_post).  out[i*11 + {f0..f3, parr | f0..f3, parr | up}]
template <bool TPG, bool AXI>
static void upwind_n(const Gas& g, int64_t n, const double* face, const double* uL, const double* uR, double* out) {
    for (int64_t i = 0; i < n; i++) {
        unsigned fl = 0;
        const double nx = face[4 * i], ny = face[4 * i + 1], len = face[4 * i + 2], yf = face[4 * i + 3];
        FaceGeom fg;
        fg.nx = nx; fg.ny = ny; fg.len = len; fg.yf = yf; fg.dx = -ny * len; fg.dy = nx * len; fg.rl = 0.0;
        const FaceState L = face_state<false, TPG>(g, uL[4 * i], uL[4 * i + 1], uL[4 * i + 2], fmax(uL[4 * i + 3], 201.0), fl);
        const FaceState R = face_state<false, TPG>(g, uR[4 * i], uR[4 * i + 1], uR[4 * i + 2], fmax(uR[4 * i + 3], 201.0), fl);
        const LdfssEarly e = ldfss_early(fg, L, R);
        const Flux a = ldfss_upwind_left<AXI>(fg, L, e.ml);
        const Flux b = ldfss_lean_post<AXI, false>(fg, L, R, ldfss_pre_from_early(e));
        double* o = out + 11 * i;
        o[0] = a.f0; o[1] = a.f1; o[2] = a.f2; o[3] = a.f3; o[4] = a.parr;
        o[5] = b.f0; o[6] = b.f1; o[7] = b.f2; o[8] = b.f3; o[9] = b.parr;
        o[10] = e.up ? 1.0 : 0.0;
    }
}
extern "C" int hp0_test_ldfss_upwind(const par2d_config* cfg, int64_t n, const double* face, const double* uL, const double* uR, double* out) {
    const Gas g = make_gas(*cfg);
    const bool tpg = !(cfg->igam > 0), axi = cfg->iaxi == 1;
    if (tpg) { if (axi) upwind_n<true, true>(g, n, face, uL, uR, out); else upwind_n<true, false>(g, n, face, uL, uR, out); }
    else { if (axi) upwind_n<false, true>(g, n, face, uL, uR, out); else upwind_n<false, false>(g, n, face, uL, uR, out); }
    return 0;
}
#endif

extern "C" {
// cell_volume over n quads: xy[i*8 + {x1,y1,...,x4,y4}]
int HP(test_volume)(int64_t n, const double* xy, double* vol) {
    for (int64_t i = 0; i < n; i++) {
        const double* q = xy + 8 * i;
        vol[i] = cell_volume(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
    }
    return 0;
}

// accuracy probe of the fast reciprocal / rsqrt / sqrt forms on the emulated seeds
int HP(test_fastmath)(int64_t n, const double* x, double* out3) {
    for (int64_t i = 0; i < n; i++) { out3[3 * i] = fast_rcp(x[i]); out3[3 * i + 1] = fast_rsqrt(x[i]); out3[3 * i + 2] = fast_sqrt(x[i]); }
    return 0;
}

}  // extern "C"
