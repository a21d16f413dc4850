// physics.cuh -- point-wise device functions of the PAR2D hot path (sm_100a, fp64).
//
// Everything is templated on
//   FAITHFUL : true  -> the reference's operation order, IEEE division/sqrt, pivoted
//                       LU; the including TU is compiled with -fmad=false so nothing is
//                       contracted.  Bit-identical to the reference except through pow().
//              false -> FMA contraction, Newton reciprocals shared between uses,
//                       closed-form triangular solve.
//   TPG      : NASA-7 curve-fit thermodynamics (igam <= 0) instead of the calorically
//              perfect shortcut (Thermo.cpp:14,34).
//   AXI      : axisymmetric (IAXI == 1) or planar.
// Reference lines are cited per function.
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define PAR2D_FN __device__ __forceinline__
#else
// Host build of the same formulas, used ONLY by tests/host_physics (CPU-side check of the point-wise
// functions against the oracle without a GPU).  The MUFU seeds are emulated at their measured accuracy
// (2^-20) and the never-contracted intrinsics by plain operations (that TU is built with
// -ffp-contract=off or, for the fast mode, with FMA contraction like nvcc's default).
#include <cmath>
#include <cstring>
#define PAR2D_FN inline
namespace par2d {
using std::isnan;
using std::signbit;
inline double par2d_trunc20(double v) { uint64_t b; memcpy(&b, &v, 8); b &= 0xffffffff00000000ull; memcpy(&v, &b, 8); return v; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
}  // namespace par2d
#endif

#ifndef LDFSS_UPWIND_VOTE
#define LDFSS_UPWIND_VOTE 1
#endif
#ifndef LDFSS_VOTE_MIRROR
#define LDFSS_VOTE_MIRROR 0   // 1: a second vote for 'every face supersonic from the right', 2: one vote for both
                              // directions with the upwind side picked by selects.  Measured on B200 (march kernel,
                              // +x flow): 0 -> 0.5195 ms, 1 -> 0.5276 ms, 2 -> 0.5341 ms: every extra vote/branch on the
                              // general path costs more than the shortcut saves, so only the left-upwind case is voted
#endif
#ifndef LDFSS_VOTE_EARLY
#define LDFSS_VOTE_EARLY 1    // the fully-upwind vote is taken on the un-normalised normal velocities, before the rsqrt (see ldfss)
#endif
#ifndef LDFSS_VOTE_MIN
#define LDFSS_VOTE_MIN 1      // early vote on min(ml, mr): one square and one compare instead of two (0.5727 -> 0.5701 ms, two runs each)
#endif
#ifndef LDFSS_BRANCHFREE
#define LDFSS_BRANCHFREE 1    // the subsonic diffusion term is evaluated without a branch and switched off by a select: measured
                              // +1.3% on B200 together with the warp vote (fully supersonic warps skip it anyway; fewer basic
                              // blocks for the scheduler elsewhere); 0: branch around it (A/B runs)
#endif

namespace par2d {

// seeds of the reciprocal / reciprocal square root (MUFU.RCP64H, MUFU.RSQ64H: 2^-20 relative)
PAR2D_FN double seed_rcp(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#elif !defined(__CUDACC__)
    return par2d_trunc20(1.0 / par2d_trunc20(x));
#else
    return 1.0 / x;   // nvcc's host pass only parses this __device__ function
#endif
}
PAR2D_FN double seed_rsqrt(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#elif !defined(__CUDACC__)
    return par2d_trunc20(1.0 / sqrt(par2d_trunc20(x)));
#else
    return 1.0 / sqrt(x);
#endif
}

// Physics constants of one run; lives in kernel parameter space.
struct Gas {
    double Rs;         // Thermo.h:31
    double gam;        // fabs(IGAM), Thermo.h:47
    double igam;       // signed IGAM
    double gamRs;      // gam*Rs (same association as StateVariables.h:51: (gam*Rs)*T)
    double gamRs4;     // gam*Rs/4: sqrt(gamRs4*T) = a/2 (fast path only)
    double cp_cpg;     // Rs/(1-1/IGAM)             Thermo.cpp:34
    double cpg_den;    // (1-1/IGAM)                Thermo.cpp:14
    double inv_cv_cpg; // 1/(cp_cpg-Rs)   (fast path only)
    double hi[7], lo[7];
    double uFS[4];
    double p_back;
    double cfl;
};

// error bits raised on the device (mapped to PAR2D_ERR_* on the host)
enum : unsigned {
    FLAG_NAN = 1u, FLAG_DENSITY = 2u, FLAG_STATE = 4u, FLAG_SINGULAR = 8u, FLAG_THERMO = 16u, FLAG_COMM = 32u
};

// max(x, lo) for a non-NaN bound lo, as compare + select (3 instructions).  sm_100 has no fp64 max
// instruction: fmax() expands to DSETP.MAX + selects + NaN fix-up (7 instructions).  Same result as
// fmax(x, lo) for every x including NaN (-> lo).
PAR2D_FN double max_sel(double x, double lo) {
#ifdef __CUDA_ARCH__
    double r;   // spelled in PTX: nvcc canonicalises the C++ ternary back into fmax()
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(x), "d"(lo));
    return r;
#else
    return x > lo ? x : lo;
#endif
}

// ---------------------------------------------------------------- fast math
// MUFU.RCP64H / MUFU.RSQ64H seed 2^-20 relative (measured on B200, profiles/r01_hw_probe_*.txt).
// One third-order step in the seed's residual e (|e| <= 2^-19) leaves a truncation error of e^3 <= 2^-57,
// a sixteenth of the rounding error of the last operation; shorter (in operations and in dependency
// chain) than two Newton steps.  PAR2D_SERIES4 = 1 keeps one more term (e^4 = 2^-76), for A/B runs.
#ifndef PAR2D_SERIES4
#define PAR2D_SERIES4 0
#endif
PAR2D_FN double fast_rcp(double x) {
    const double y = seed_rcp(x);
    const double e = fma(-x, y, 1.0);          // 1 - x*y0
#if PAR2D_SERIES4
    const double y1 = fma(y, e, y);            // y0 (1 + e)
    const double e2 = e * e;
    return fma(y1, e2, y1);                    // y0 (1 + e)(1 + e^2) = y0 (1 + e + e^2 + e^3)
#else
    return fma(y, fma(e, e, e), y);            // y0 (1 + e + e^2)
#endif
}
// 1/sqrt(x): y0 (1 + e/2 + 3e^2/8 [+ 5e^3/16]), e = 1 - x*y0^2.
PAR2D_FN double fast_rsqrt(double x) {
    const double y = seed_rsqrt(x);
    const double e = fma(-x, y * y, 1.0);
#if PAR2D_SERIES4
    const double p = fma(fma(e, 0.3125, 0.375), e, 0.5);
#else
    const double p = fma(e, 0.375, 0.5);
#endif
    return fma(y * e, p, y);
}
// sqrt(x), x > 0 normal: g = x*y0 ~ sqrt(x), same series in e = 1 - g*y0.
PAR2D_FN double fast_sqrt(double x) {
    const double y = seed_rsqrt(x);
    const double gq = x * y;
    const double e = fma(-gq, y, 1.0);
#if PAR2D_SERIES4
    const double p = fma(fma(e, 0.3125, 0.375), e, 0.5);
#else
    const double p = fma(e, 0.375, 0.5);
#endif
    return fma(gq * e, p, gq);
}

// ---------------------------------------------------------------- thermo
// Thermo::CalcEnthalpy (Thermo.cpp:7-30) and CalcCp (Thermo.cpp:32-48).
template <bool TPG>
PAR2D_FN double enthalpy(const Gas& g, double T, unsigned& flags) {
    if constexpr (!TPG) {
        return (T * g.Rs / g.cpg_den + g.hi[5]);
    } else {
        double T2 = T * T, T3 = T * T2, T4 = T * T3, T5 = T * T4;
        if (!(T > 200.0) || !(T < 10000.0)) flags |= FLAG_THERMO;
        const double* q = (T > 1000.0) ? g.hi : g.lo;
        return (g.Rs) * (q[0] * T + q[1] * T2 * 0.5 + q[2] * T3 / 3.0 + q[3] * T4 * 0.25 + q[4] * T5 * 0.2 + q[5]);
    }
}
template <bool TPG>
PAR2D_FN double heat_capacity(const Gas& g, double T, unsigned& flags) {
    if constexpr (!TPG) {
        return g.cp_cpg;
    } else {
        double T2 = T * T, T3 = T * T2, T4 = T * T3;
        if (!(T > 200.0) || !(T < 10000.0)) flags |= FLAG_THERMO;
        const double* q = (T >= 1000.0) ? g.hi : g.lo;
        return (g.Rs) * (q[0] + q[1] * T + q[2] * T2 + q[3] * T3 + q[4] * T4);
    }
}

// What a face needs from a cell: State::UpdateState (StateVariables.h:38-68) minus
// the members the path never reads.  T must already be clamped (>= 201).
struct FaceState {
    double rho, u, v, p;
    double a;          // speed of sound; the lean path (FAITHFUL = false) carries a/2, see ldfss
    double h0;
};

template <bool FAITHFUL, bool TPG>
PAR2D_FN FaceState face_state(const Gas& g, double rho, double u, double v, double T, unsigned& flags) {
    FaceState s;
    s.rho = rho; s.u = u; s.v = v;
    double v2 = u * u + v * v;
    if constexpr (FAITHFUL) s.p = fmax(rho * g.Rs * T, 1.0);
    else s.p = max_sel(rho * g.Rs * T, 1.0);
    if constexpr (FAITHFUL) {
        s.a = sqrt(g.gamRs * T);
        s.h0 = enthalpy<TPG>(g, T, flags) + 0.5 * v2;
        if (!(s.a > 0.0) || !(s.p > 0.0) || isnan(s.p * s.a * T)) flags |= FLAG_STATE;
    } else {
        // T >= 201 and p >= 1 by the clamps, so a > 0 and p > 0 hold by construction; NaNs surface in
        // the residual norm (FLAG_NAN) and in the density check of the update
        s.a = fast_sqrt(g.gamRs4 * T);             // a/2: ldfss only ever needs (aL + aR)/2
        if constexpr (TPG) s.h0 = fma(0.5, v2, enthalpy<TPG>(g, T, flags));
        else s.h0 = fma(0.5, v2, fma(g.cp_cpg, T, g.hi[5]));   // Thermo.cpp:14 with Rs/(1-1/IGAM) folded
    }
    return s;
}

// ---------------------------------------------------------------- LDFSS
// EulerFlux.cpp:121-192.  The face is given by its node difference (dx,dy) =
// (x1-x0, y1-y0); MeshModule.cpp:52-81 defines len = |d|, n = (dy,-dx)/len.
struct FaceGeom {
    double dx, dy;    // node difference
    double len, rl;   // |d| and 1/|d|
    double nx, ny;    // unit normal (faithful path: the reference's divided values)
    double yf;        // face mid-point radius (MeshModule.cpp:66,84)
};

template <bool FAITHFUL>
PAR2D_FN FaceGeom face_geom(double x0, double y0, double x1, double y1) {
    FaceGeom f;
    f.dx = x1 - x0;
    f.dy = y1 - y0;
    if constexpr (FAITHFUL) {
        f.len = sqrt((f.dx * f.dx) + (f.dy * f.dy));
        f.nx = f.dy / f.len;
        f.ny = -f.dx / f.len;
        f.rl = 0.0;
    } else {
        // the lean LDFSS works on the node difference itself: len*n = (dy,-dx), and |d| only enters
        // through len*ahalf, which it gets from one rsqrt (see ldfss)
        f.len = f.rl = f.nx = f.ny = 0.0;
    }
    f.yf = 0.5 * (y1 + y0);
    return f;
}

// Cell volume, MeshModule.cpp:34-40.  The shoelace sum cancels to ~1e-9 relative on fine meshes, so
// its ROUNDING is part of the reference's result: both modes evaluate it in the reference's order with
// individually rounded products and sums (the intrinsics are never contracted into FMAs).
PAR2D_FN double cell_volume(double x1, double y1, double x2, double y2, double x3, double y3, double x4, double y4) {
    double s = __dmul_rn(x1, y2);
    s = __dadd_rn(s, __dmul_rn(x2, y3));
    s = __dadd_rn(s, __dmul_rn(x3, y4));
    s = __dadd_rn(s, __dmul_rn(x4, y1));
    s = __dsub_rn(s, __dmul_rn(x2, y1));
    s = __dsub_rn(s, __dmul_rn(x3, y2));
    s = __dsub_rn(s, __dmul_rn(x4, y3));
    s = __dsub_rn(s, __dmul_rn(x1, y4));
    return fabs(__dmul_rn(0.5, s));
}

struct Flux {
    double f0, f1, f2, f3, parr;
};

PAR2D_FN double sgn_bit(double x) { return signbit(x) ? -1.0 : 1.0; }  // Indexing.h:22

// The lean LDFSS in two parts, so that a caller can evaluate the Mach-number parts of several faces in one block
// before any of them reaches the warp vote (march.cuh): ldfss_lean_pre = len*ahalf and the two face Mach numbers
// from ONE rsqrt; ldfss_lean_post = the splitting and the flux.
struct LdfssPre {
    double la, xml, xmr;
};
PAR2D_FN LdfssPre ldfss_lean_pre(const FaceGeom& fg, const FaceState& L, const FaceState& R) {
    LdfssPre p;
    const double ahalf = L.a + R.a;          // the lean FaceState carries a/2
    const double w = fma(fg.dx, fg.dx, fg.dy * fg.dy) * (ahalf * ahalf);
    const double q = fast_rsqrt(w);          // 1 / (len * ahalf)
    p.la = w * q;                            // len * ahalf
    p.xml = fma(L.u, fg.dy, -L.v * fg.dx) * q;
    p.xmr = fma(R.u, fg.dy, -R.v * fg.dx) * q;
    return p;
}
template <bool AXI, bool WARP_VOTE>
PAR2D_FN Flux ldfss_lean_post(const FaceGeom& fg, const FaceState& L, const FaceState& R, const LdfssPre& pre) {
    Flux o;
    // lean evaluation: one rsqrt of (|d| ahalf)^2 yields both len*ahalf and its reciprocal, and
    // the pressure-weighted diffusion term (with its three divisions) is
    // evaluated only where both sides are subsonic (it is multiplied by btl*btr = 0 otherwise).
    const double la = pre.la, xml = pre.xml, xmr = pre.xmr;
#if defined(__CUDA_ARCH__) && LDFSS_UPWIND_VOTE
    if constexpr (WARP_VOTE) {
#if LDFSS_VOTE_MIRROR == 2
        // ONE vote for both directions: every face of the warp supersonic, each from its own upwind side
        // (all/alr = 1/0 or 0/1, btl = btr = 0); the upwind state is picked by selects
        const bool fromL = xml >= 1.0 && xmr >= 1.0, fromR = xml <= -1.0 && xmr <= -1.0;
        if (__all_sync(0xffffffffu, fromL || fromR)) {
            const double rho = fromL ? L.rho : R.rho, uu = fromL ? L.u : R.u, vv = fromL ? L.v : R.v;
            const double pp = fromL ? L.p : R.p, hh = fromL ? L.h0 : R.h0, mm = fromL ? xml : xmr;
            if constexpr (AXI) {
                const double fm = ((la * fg.yf) * rho) * mm;
                o.parr = -pp * fg.dx;
                o.f0 = fm;
                o.f1 = fma(fm, uu, (pp * fg.yf) * fg.dy);
                o.f2 = fm * vv;
                o.f3 = fm * hh;
            } else {
                const double fm = (la * rho) * mm;
                o.parr = 0.0;
                o.f0 = fm;
                o.f1 = fma(fm, uu, pp * fg.dy);
                o.f2 = fma(fm, vv, -pp * fg.dx);
                o.f3 = fm * hh;
            }
            return o;
        }
#else
        if (__all_sync(0xffffffffu, xml >= 1.0 && xmr >= 1.0)) {          // all/alr = 1/0, btl = btr = 0
            if constexpr (AXI) {
                const double fml = ((la * fg.yf) * L.rho) * xml;
                o.parr = -L.p * fg.dx;
                o.f0 = fml;
                o.f1 = fma(fml, L.u, (L.p * fg.yf) * fg.dy);
                o.f2 = fml * L.v;
                o.f3 = fml * L.h0;
            } else {
                const double fml = (la * L.rho) * xml;
                o.parr = 0.0;
                o.f0 = fml;
                o.f1 = fma(fml, L.u, L.p * fg.dy);
                o.f2 = fma(fml, L.v, -L.p * fg.dx);
                o.f3 = fml * L.h0;
            }
            return o;
        }
        if (LDFSS_VOTE_MIRROR && __all_sync(0xffffffffu, xml <= -1.0 && xmr <= -1.0)) {        // all/alr = 0/1
            if constexpr (AXI) {
                const double fmr = ((la * fg.yf) * R.rho) * xmr;
                o.parr = -R.p * fg.dx;
                o.f0 = fmr;
                o.f1 = fma(fmr, R.u, (R.p * fg.yf) * fg.dy);
                o.f2 = fmr * R.v;
                o.f3 = fmr * R.h0;
            } else {
                const double fmr = (la * R.rho) * xmr;
                o.parr = 0.0;
                o.f0 = fmr;
                o.f1 = fma(fmr, R.u, R.p * fg.dy);
                o.f2 = fma(fmr, R.v, -R.p * fg.dx);
                o.f3 = fmr * R.h0;
            }
            return o;
        }
#endif
    }
#endif
    const bool subL = fabs(xml) < 1.0, subR = fabs(xmr) < 1.0;
    // split Mach numbers and pressure weights (EulerFlux.cpp:146-175 with all/alr in {0,1} and
    // btl/btr in {-1,0} substituted).  Selects, not branches: skipping the polynomials on
    // supersonic faces was measured 1.6% SLOWER on B200 (the kernel is issue-bound, a branch
    // costs more slots than the eight operations it skips).
    const double hl = fma(0.5, xml, 0.5), hr = fma(0.5, xmr, -0.5);
    const double xmml = hl * hl, xmmr = -(hr * hr);
    double cep = subL ? xmml : (signbit(xml) ? 0.0 : xml);
    double cem = subR ? xmmr : (signbit(xmr) ? xmr : 0.0);
    const double wl = subL ? xmml * (2.0 - xml) : (signbit(xml) ? 0.0 : 1.0);
    const double wr = subR ? -xmmr * (2.0 + xmr) : (signbit(xmr) ? 1.0 : 0.0);
    {
        // subsonic diffusion term, EulerFlux.cpp:152-163.
        // xmc = (xmhalf-1)^2/4 with xmhalf = sqrt((xml^2+xmr^2)/2) (:152-154).  Here xs = xmhalf/sqrt(2) and
        // d = xs - 1/sqrt(2), so d*d = 2 xmc.  The 1e-199 keeps the argument of rsqrt.approx (which flushes
        // subnormals) normal: it vanishes in the sum unless both Mach numbers are below 1e-92, where xs is 0 to 90
        // digits of d either way.  No compare, no branch.
        // Pressure switches (:157-160) over their common denominator: with delp = pL - pR, psum = pL + pR and
        // N = pL pR - |delp| psum,
        //   1 - (delp/psum + 2|delp|/pL) = 2 N / (pL psum),   1 + (delp/psum - 2|delp|/pR) = 2 N / (pR psum)
        // (psum -+ delp = 2 pR, 2 pL), so ONE reciprocal w3 = 1/(pL pR psum) and ONE max(.,0) serve both:
        //   xmcp = xmc max(0, 2 N pR w3) = (2 xmc) max(0,N) w3 pR,   xmcm = (2 xmc) max(0,N) w3 pL
#if LDFSS_BRANCHFREE
        const bool both = true;        // evaluated on every face, switched off by the select below (one basic block)
#else
        const bool both = subL && subR;
#endif
        if (both) {
            const double t = 0.25 * fma(xml, xml, fma(xmr, xmr, 1e-199));
            const double xs = t * fast_rsqrt(t);
            const double d = xs - 0.70710678118654752440;
            const double delp = L.p - R.p, psum = L.p + R.p;
            const double plr = L.p * R.p;
            const double w3 = fast_rcp(plr * psum);
            const double N = max_sel(fma(-fabs(delp), psum, plr), 0.0);
            const double k = (subL && subR) ? ((d * d) * w3) * N : 0.0;
            cep = fma(-k, R.p, cep);
            cem = fma(k, L.p, cem);
        }
    }
    const double pnet = fma(wl, L.p, wr * R.p);
    if constexpr (AXI) {
        // every flux component carries yface (EulerFlux.cpp:178-186): fold it into len*ahalf and pnet
        const double lay = la * fg.yf;
        const double fml = lay * L.rho * cep;
        const double fmr = lay * R.rho * cem;
        const double px = (pnet * fg.yf) * fg.dy;   // yface*len*pnet*normx
        o.parr = -pnet * fg.dx;                     // len*pnet*normy, without yface (:181)
        o.f0 = fml + fmr;
        o.f1 = fma(fml, L.u, fma(fmr, R.u, px));
        o.f2 = fma(fml, L.v, fmr * R.v);
        o.f3 = fma(fml, L.h0, fmr * R.h0);
    } else {
        const double fml = la * L.rho * cep;
        const double fmr = la * R.rho * cem;
        o.parr = 0.0;
        o.f0 = fml + fmr;
        o.f1 = fma(fml, L.u, fma(fmr, R.u, pnet * fg.dy));
        o.f2 = fma(fml, L.v, fma(fmr, R.v, -pnet * fg.dx));
        o.f3 = fma(fml, L.h0, fmr * R.h0);
    }
    return o;
}

// The early form of the fully-upwind test.  xml >= 1  <=>  ml = len*(uL.n) >= len*ahalf  <=>  ml > 0 and ml^2 >= w with
// w = (len*ahalf)^2, and a fully upwind flux needs neither the Mach numbers nor len*ahalf - la*xml IS ml = uL dy - vL dx.
// So the test needs no rsqrt, and a face that passes it none either (ldfss_upwind_left).  Equal to the general path to
// rounding (2 ulp of the mass flux), not bit for bit: on the device a face is evaluated either way depending on the
// other 31 faces of its warp, so fast-mode results are decomposition-invariant to ~1e-15 relative, not bitwise (the
// faithful mode never comes here).  tests/test_host_physics_cpu.py checks both forms against each other on the host.
struct LdfssEarly {
    double w, ml, mr;        // (len*ahalf)^2 and the un-normalised normal velocities len*(u.n) of both sides
    bool up;                 // both sides supersonic from the left
};
PAR2D_FN LdfssEarly ldfss_early(const FaceGeom& fg, const FaceState& L, const FaceState& R) {
    LdfssEarly e;
    const double ahalf = L.a + R.a;          // the lean FaceState carries a/2
    e.w = fma(fg.dx, fg.dx, fg.dy * fg.dy) * (ahalf * ahalf);
    e.ml = fma(L.u, fg.dy, -L.v * fg.dx);
    e.mr = fma(R.u, fg.dy, -R.v * fg.dx);
#if LDFSS_VOTE_MIN
    const double mm = e.ml < e.mr ? e.ml : e.mr;
    e.up = mm > 0.0 && mm * mm >= e.w;
#else
    e.up = e.ml > 0.0 && e.mr > 0.0 && e.ml * e.ml >= e.w && e.mr * e.mr >= e.w;
#endif
    return e;
}
// the flux of a face that is supersonic from the left on both sides: all/alr = 1/0, btl = btr = 0 in EulerFlux.cpp:146-186
template <bool AXI>
PAR2D_FN Flux ldfss_upwind_left(const FaceGeom& fg, const FaceState& L, double ml) {
    Flux o;
    if constexpr (AXI) {
        const double fml = (fg.yf * L.rho) * ml;
        o.parr = -L.p * fg.dx;
        o.f0 = fml;
        o.f1 = fma(fml, L.u, (L.p * fg.yf) * fg.dy);
        o.f2 = fml * L.v;
        o.f3 = fml * L.h0;
    } else {
        const double fml = L.rho * ml;
        o.parr = 0.0;
        o.f0 = fml;
        o.f1 = fma(fml, L.u, L.p * fg.dy);
        o.f2 = fma(fml, L.v, -L.p * fg.dx);
        o.f3 = fml * L.h0;
    }
    return o;
}
// ... and the general path's inputs when the test fails: the rsqrt after all
PAR2D_FN LdfssPre ldfss_pre_from_early(const LdfssEarly& e) {
    LdfssPre pre;
    const double q = fast_rsqrt(e.w);
    pre.la = e.w * q;
    pre.xml = e.ml * q;
    pre.xmr = e.mr * q;
    return pre;
}

// WARP_VOTE (device, lean path, all 32 lanes of the warp calling together): when EVERY face of the warp is
// supersonic in the same direction the flux is fully upwind - cep = xml, cem = 0, pnet = pL (or the mirror
// image) - and the split polynomials, their selects and the dead side's terms are skipped.  The result is
// bit-identical to the general path (the skipped terms are exact zeros there).
template <bool FAITHFUL, bool AXI, bool WARP_VOTE = false>
PAR2D_FN Flux ldfss(const FaceGeom& fg, const FaceState& L, const FaceState& R) {
    Flux o;
    if constexpr (FAITHFUL) {
        const double normx = fg.nx, normy = fg.ny, len = fg.len;
        double yface = fg.yf;
        double ahalf = 0.5 * (L.a + R.a);
        double xml = (L.u * normx + L.v * normy) / ahalf;
        double xmr = (R.u * normx + R.v * normy) / ahalf;
        double all = 0.5 * (1.0 + sgn_bit(xml));
        double alr = 0.5 * (1.0 - sgn_bit(xmr));
        double btl = -fmax(0.0, 1.0 - (double)((int)(fabs(xml))));
        double btr = -fmax(0.0, 1.0 - (double)((int)(fabs(xmr))));
        double xmml = 0.25 * (xml + 1.0) * (xml + 1.0);
        double xmmr = -0.25 * (xmr - 1.0) * (xmr - 1.0);
        double xmhalf = sqrt(0.5 * (xml * xml + xmr * xmr));
        double xmc = 0.25 * btl * btr * (xmhalf - 1.0) * (xmhalf - 1.0);
        double delp = L.p - R.p;
        double psum = L.p + R.p;
        double xmcp = xmc * fmax(0.0, (1.0 - (delp / psum + 2.0 * fabs(delp) / L.p)));
        double xmcm = xmc * fmax(0.0, (1.0 + (delp / psum - 2.0 * fabs(delp) / R.p)));
        double cvlp = all * (1.0 + btl) * xml - btl * xmml;
        double cvlm = alr * (1.0 + btr) * xmr - btr * xmmr;
        double cep = cvlp - xmcp;
        double cem = cvlm + xmcm;
        double fml = len * L.rho * ahalf * cep;
        double fmr = len * R.rho * ahalf * cem;
        double ppl = 0.25 * (xml + 1.0) * (xml + 1.0) * (2.0 - xml);
        double ppr = 0.25 * (xmr - 1.0) * (xmr - 1.0) * (2.0 + xmr);
        double pnet = (all * (1.0 + btl) - btl * ppl) * L.p + (alr * (1.0 + btr) - btr * ppr) * R.p;
        double py_planar;
        if constexpr (AXI) {
            o.parr = len * pnet * normy;
            py_planar = 0.0;
        } else {
            yface = 1.0;
            o.parr = 0.0;
            py_planar = len * pnet * normy;
        }
        o.f0 = yface * (fml + fmr);
        o.f1 = yface * (fml * L.u + fmr * R.u + len * pnet * normx);
        o.f2 = yface * (fml * L.v + fmr * R.v + py_planar);
        o.f3 = yface * (fml * L.h0 + fmr * (R.h0));
    } else {
#if defined(__CUDA_ARCH__) && LDFSS_UPWIND_VOTE && LDFSS_VOTE_EARLY
        if constexpr (WARP_VOTE) {
            // the vote BEFORE the rsqrt (see ldfss_early): a warp whose faces are all supersonic from the left skips the
            // rsqrt series, the three products behind it and the split; the others pay one square and a compare
            const LdfssEarly e = ldfss_early(fg, L, R);
            if (__all_sync(0xffffffffu, e.up)) return ldfss_upwind_left<AXI>(fg, L, e.ml);
            return ldfss_lean_post<AXI, false>(fg, L, R, ldfss_pre_from_early(e));
        }
#endif
        o = ldfss_lean_post<AXI, WARP_VOTE>(fg, L, R, ldfss_lean_pre(fg, L, R));
    }
    return o;
}

// ---------------------------------------------------------------- boundary conditions
// boundary_state + SubsonInflo/SubsonOutfl (BoundaryConditions.cpp:12-190), followed by
// the ghost's UpdateState temperature clamp (SpatialDiscretization.cpp:39).  O(perimeter)
// work, so a single (reference-ordered) formulation serves both modes.  Returns false when
// the ghost must stay unwritten (NaN normal Mach number).
// The isentropic pair of the subsonic far field (BoundaryConditions.cpp:163-187) in the reference's order, two pow():
// rho = ref X^(gam-1), pg = p (rho/ref)^gam.
PAR2D_FN void isentropic_ghost(double ref, double X, double gm1, double gam, double p_int, double& rho, double& pg) {
    rho = ref * pow(X, gm1);
    pg = p_int * pow(rho / ref, gam);
}

#if defined(PAR2D_LEAN_BC)
// The far field (code 1) of the fast mode's translation unit (PAR2D_LEAN_BC, kernels_fast.cu): the four cases of
// BoundaryConditions.cpp:150-190 with the two subsonic ones merged, series reciprocals instead of divisions, and the
// isentropic pair from ONE logarithm - rho = ref exp((gam-1) ln X), pg = p exp(gam (gam-1) ln X) - : ~150 instructions
// instead of ~500, on code that runs once per step and strip out of a cold instruction cache (the persistent launch of
// the shipped nozzle / aerospike cases).  Equal to the reference-ordered form to ~1e-15 relative
// (tests/test_host_physics_cpu.py on the host, test_gpu_parity.py on the device).
PAR2D_FN bool farfield_lean(const Gas& g, double normx, double normy, const double ui[4], double a_int, double p_int, double out[4]) {
    const double vin = ui[1] * normx + ui[2] * normy;
    const double ia = fast_rcp(a_int);
    const double mn = vin * ia;
    if (mn <= -1) {
        out[0] = g.uFS[0]; out[1] = g.uFS[1]; out[2] = g.uFS[2]; out[3] = g.uFS[3];
    } else if (mn >= 1) {
        out[0] = ui[0]; out[1] = ui[1]; out[2] = ui[2]; out[3] = ui[3];
    } else if (mn > -1 && mn < 1) {
        const bool in = mn <= 0;                   // inflow: tangential velocity and reference density of the free stream
        const double gm1 = g.gam - 1;
        const double vfn = g.uFS[1] * normx + g.uFS[2] * normy;
        const double bu = in ? g.uFS[1] : ui[1], bv = in ? g.uFS[2] : ui[2], bn = in ? vfn : vin;
        const double vgn = 0.5 * (vin + vfn);
        const double r = (a_int + 0.5 * gm1 * (vgn - vfn)) * ia;
        const double lx = log(r * r);
        const double rho = (in ? g.uFS[0] : ui[0]) * exp(gm1 * lx);
        out[0] = rho;
        out[1] = bu + (vgn - bn) * normx;
        out[2] = bv + (vgn - bn) * normy;
        out[3] = p_int * exp((g.gam * gm1) * lx) * fast_rcp(rho * g.Rs);
    } else {
        return false;                              // NaN normal Mach number
    }
    return true;
}
#endif

PAR2D_FN bool boundary_state(const Gas& g, int btype, double normx, double normy,
                                               const double ui[4], double a_int, double p_int, double v2_int,
                                               double out[4]) {
    const double uL = ui[1], vL = ui[2];
    const double vn = uL * normx + vL * normy;
    if (btype == -1) btype = 3;
    if (btype == 0) btype = 4;
    bool wrote = true;
    if (btype == 4) {
        out[0] = ui[0];
        out[3] = ui[3];
        out[1] = uL - 2 * vn * normx;
        out[2] = vL - 2 * vn * normy;
    } else if (btype == 2) {
        out[0] = g.p_back / (g.Rs * ui[3]);
        out[1] = fmax(0.0, ui[1]);
        out[2] = ui[2];
        out[3] = ui[3];
    } else if (btype == 3) {
        out[0] = ui[0]; out[1] = ui[1]; out[2] = ui[2]; out[3] = ui[3];
    } else if (btype == 5) {
        double pFS = g.uFS[0] * g.Rs * g.uFS[3];
        double pB = pFS - 0.5 * ui[0] * v2_int;
        if (ui[1] <= 0.0) pB = pFS;
        out[0] = pB / (g.Rs * g.uFS[3]);
        out[1] = fmax(ui[1], 0.0);
        out[2] = ui[2];
        out[3] = ui[3];
#if defined(PAR2D_LEAN_BC)
    } else if (btype == 1) {
        wrote = farfield_lean(g, normx, normy, ui, a_int, p_int, out);
#else
    } else if (btype == 1) {
        const double gm1 = g.gam - 1;
        const double mn = vn / a_int;      // only the far field reads the normal Mach number
        if (mn <= -1) {
            out[0] = g.uFS[0]; out[1] = g.uFS[1]; out[2] = g.uFS[2]; out[3] = g.uFS[3];
        } else if (mn <= 0 && mn > -1) {
            // far state = free stream, var0 = interior state (BoundaryConditions.cpp:163)
            double vin = (uL * normx) + (vL * normy);
            double vfn = (g.uFS[1] * normx) + (g.uFS[2] * normy);
            double vftx = g.uFS[1] - vfn * normx;
            double vfty = g.uFS[2] - vfn * normy;
            double vgn = 0.5 * (vin + vfn) + (1 / gm1) * (a_int - a_int);
            double cg = a_int + 0.5 * gm1 * (vgn - vfn);
            double pg;
            isentropic_ghost(g.uFS[0], cg * cg / (a_int * a_int), gm1, g.gam, p_int, out[0], pg);
            out[1] = (vftx + vgn * normx);
            out[2] = (vfty + vgn * normy);
            out[3] = pg / (out[0] * g.Rs);
        } else if (mn >= 1) {
            out[0] = ui[0]; out[1] = ui[1]; out[2] = ui[2]; out[3] = ui[3];
        } else if (mn > 0 && mn < 1) {
            double vin = (uL * normx) + (vL * normy);
            double vfn = (g.uFS[1] * normx) + (g.uFS[2] * normy);
            double vitx = uL - vin * normx;
            double vity = vL - vin * normy;
            double vgn = 0.5 * (vin + vfn) + (1 / gm1) * (a_int - a_int);
            double cg = a_int + 0.5 * gm1 * (vgn - vfn);
            double pg;
            isentropic_ghost(ui[0], cg * cg / (a_int * a_int), gm1, g.gam, p_int, out[0], pg);
            out[1] = (vitx + vgn * normx);
            out[2] = (vity + vgn * normy);
            out[3] = pg / (out[0] * g.Rs);
        } else {
            wrote = false;  // NaN normal Mach number
        }
#endif
    } else {
        wrote = false;  // unknown code: the reference falls through without writing
    }
    if (wrote) out[3] = fmax(out[3], 201.0);
    return wrote;
}

// Closed-form forward substitution of D dv = res (D lower triangular, Jacobian.cpp:7-47) with every
// scalar factor folded into one: sc = dt/ycc for a residual r = R/vol, or dt/(vol*ycc) when R is the
// raw flux sum (planar: without ycc).
//   drho = sc R0;  du = sc (R1 - u R0)/rho;  dv = sc (R2 - v R0)/rho;
//   dT   = sc (R3 - e0 R0 - u (R1 - u R0) - v (R2 - v R0)) / (rho Cv)
PAR2D_FN void solve_lean(double sc, double rho, double u, double v, double e0, double inv_cv,
                                           const double R[4], double dv[4]) {
    const double sr = sc * fast_rcp(rho);
    const double m1 = fma(-u, R[0], R[1]);
    const double m2 = fma(-v, R[0], R[2]);
    const double m3 = fma(-v, m2, fma(-u, m1, fma(-e0, R[0], R[3])));
    dv[0] = sc * R[0];
    dv[1] = sr * m1;
    dv[2] = sr * m2;
    dv[3] = (sr * inv_cv) * m3;
}

// ---------------------------------------------------------------- per-cell solve
// dv = D^-1 res (Jacobian.cpp:7-47, LUtools.cpp:22-90) and the axisymmetric 1/ycc
// (main.cpp:325-331).  scale = 1/ycc (AXI) or unused.
template <bool FAITHFUL, bool AXI>
PAR2D_FN void solve_cell(double dt, double ycc, double rho, double u, double v, double e0, double Cv,
                                           double inv_cv, const double r[4], double dv[4], unsigned& flags) {
    if constexpr (FAITHFUL) {
        double A[4][4];
        const double dti = 1.0 / dt;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) A[i][j] = 0.0;
        A[0][0] = dti * 1.0;
        A[1][0] = dti * u;
        A[2][0] = dti * v;
        A[1][1] = dti * rho;
        A[2][2] = dti * rho;
        A[3][0] = dti * e0;
        A[3][1] = dti * rho * u;
        A[3][2] = dti * rho * v;
        A[3][3] = dti * rho * Cv;
        double b[4] = {r[0], r[1], r[2], r[3]};
        // partial pivoting with physical row swaps (fully unrolled so A stays in registers)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double big = 0.0;
            int imax = i;
#pragma unroll
            for (int k = i; k < 4; k++) {
                double val = fabs(A[k][i]);
                if (val > big) { big = val; imax = k; }
            }
            if (big < 1e-16) flags |= FLAG_SINGULAR;
#pragma unroll
            for (int k = i + 1; k < 4; k++) {
                if (imax == k) {
#pragma unroll
                    for (int c = 0; c < 4; c++) { double t = A[i][c]; A[i][c] = A[k][c]; A[k][c] = t; }
                    double t = b[i]; b[i] = b[k]; b[k] = t;
                }
            }
#pragma unroll
            for (int j = i + 1; j < 4; j++) {
                A[j][i] /= A[i][i];
#pragma unroll
                for (int k = i + 1; k < 4; k++) A[j][k] -= A[j][i] * A[i][k];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            dv[i] = b[i];
#pragma unroll
            for (int k = 0; k < i; k++) dv[i] -= A[i][k] * dv[k];
        }
#pragma unroll
        for (int i = 3; i >= 0; i--) {
#pragma unroll
            for (int k = i + 1; k < 4; k++) dv[i] -= A[i][k] * dv[k];
            dv[i] /= A[i][i];
        }
        if constexpr (AXI) {
#pragma unroll
            for (int k = 0; k < 4; k++) dv[k] *= (1.0 / ycc);
        }
    } else {
        // D is lower triangular: forward substitution in closed form
        double s = dt;
        if constexpr (AXI) s = dt * fast_rcp(ycc);
        solve_lean(s, rho, u, v, e0, inv_cv, r, dv);
        (void)Cv;
    }
}

}  // namespace par2d
