/* oracle/par2d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C restatement of the reference's explicit finite-volume residual +
 * update path (MrPezser/PAR2D, FV/), all blocks in one process, written
 * against the same structs as the product C ABI (include/par2d_gpu.h) so the
 * parity tests can drive both sides with identical inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.  The product path
 * (par2d_b200/csrc) never links, imports or calls it.
 *
 * Parity pin: tests/test_oracle_cpu.py checks this file BITWISE against
 *   (1) the unmodified reference compiled by oracle/build_ref.sh
 *       (oracle/_ref/MPI2DFV, libpar2d_ref.so), which itself reproduces the
 *       reference's committed golden FV/Outputs/res.tec, and
 *   (2) the committed fixtures under tests/golden/ generated from (1).
 * Every floating-point expression below keeps the reference's association and
 * evaluation order; build with -ffp-contract=off (see oracle/Makefile).
 *
 * Each function cites the reference lines it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/par2d_gpu.h"

#define NV 4

/* ------------------------------------------------------------------ types */

typedef struct {
    double p, a, h, h0, e0, v2, Cp, Cv;
} orc_derived; /* the fields of class State that the path reads (StateVariables.h:28) */

typedef struct {
    int gid;            /* global block id */
    int nx, ny;         /* node counts */
    int ncx, ncy;       /* cell counts */
    double *x, *y;      /* nodes, j outer */
    int* ibound;        /* BC codes CCW */
    int bid[4];
    /* metrics in the reference's layouts (MeshModule.cpp:16-19) */
    double *geoel, *geofa, *yfa;
    double min_face;    /* min over cells of the two face lengths, start 1.0 (main.cpp:26-32) */
    double *unk, *res, *dv;
    orc_derived* der;   /* per cell, refreshed after every state write (ElemVar[]) */
    double* ghost[4];   /* ghost lines per side, [n*4+k] (SpatialDiscretization.cpp:182-185) */
    orc_derived* gder[4];
    double res0[NV];    /* iteration-0 norms (main.cpp:260,365-369) */
    double dt_b;
    /* second-order branch (cfg.accur == 1, par2d_oracle_dg.inc) */
    double* xfa;        /* face mid-point x (MeshModule.cpp:67,85) */
    double *ux, *uy;    /* xi / eta slopes (main.cpp:169-172) */
    double *resx, *resy, *dvx, *dvy;
    double* ghost2[4];  /* two Gauss-point ghost states per boundary face, [(2n+pt)*4+k] */
    orc_derived* gder2[4];
} orc_block;

typedef struct {
    int peer;
    int64_t n_send, n_recv;
    double *send, *recv;
} orc_peer;

typedef struct orc_ctx {
    par2d_config cfg;
    int nblocks;                 /* global */
    par2d_block_desc* table;     /* copy of the global table (no coordinates) */
    int nowned;
    orc_block* blk;              /* owned blocks in ascending global id */
    int* owned_index;            /* global id -> index into blk[] or -1 */
    orc_peer* peers;             /* one per rank (self unused) */
    double* dt_table;            /* nblocks */
    double dt;                   /* global step of the iteration in flight */
    int64_t iter;
    int error;
    /* history of the last batch */
    int hist_cap, hist_len;
    par2d_step_info* hist;
    double* blk_hist;            /* [step][owned][2] */
} orc_ctx;

static char g_err[256] = "";
static int fail(int code, const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }
const char* orc_last_error(void) { return g_err; }

/* -------------------------------------------------- shipped configuration */

/* FV/Indexing.h:12-20, FV/main.cpp:77-94, Thermo.h:25,46; coefficients are the
 * AIR record of FV/therm_air.dat in the order Thermo.h:27-29 reads them. */
void orc_default_config(par2d_config* c) {
    static const double hi[7] = {3.08792717E+00, 1.24597184E-03, -4.23718945E-07, 6.74774789E-11,
                                 -3.97076972E-15, -9.95262755E+02, 5.95960930E+00};
    static const double lo[7] = {3.56839620E+00, -6.78729429E-04, 1.55371476E-06, -3.29937060E-12,
                                 -4.66395387E-13, -1.06234659E+03, 3.71582965E+00};
    memset(c, 0, sizeof *c);
    c->iaxi = 1;
    c->mode = PAR2D_MODE_FAST;
    c->igam = 1.4;
    c->cfl = 0.01;
    c->tol = 1e-6;
    c->Rs = 8314.34 / 28.9;
    c->uFS[0] = 1000.0 / (c->Rs * 300.0);
    c->uFS[1] = 1736.0;
    c->uFS[2] = 0.0;
    c->uFS[3] = 300.0;
    c->p_back = 101325.0;
    memcpy(c->thermo_hi, hi, sizeof hi);
    memcpy(c->thermo_lo, lo, sizeof lo);
    c->device = -1;
    c->rank = 0;
    c->nranks = 1;
}

/* ------------------------------------------------------- point-wise physics */

/* Thermo::CalcEnthalpy, Thermo.cpp:7-30 */
static double enthalpy(const par2d_config* c, double T, int* err) {
    const double Rs = c->Rs;
    if (c->igam > 0) return (T * Rs / (1.0 - 1.0 / (c->igam)) + c->thermo_hi[5]);
    double T2 = T * T, T3 = T * T2, T4 = T * T3, T5 = T * T4;
    if (!(T > 200.0) || !(T < 10000.0)) *err = PAR2D_ERR_THERMO_RANGE;
    const double* q = (T > 1000.0) ? c->thermo_hi : c->thermo_lo;
    return (Rs) * (q[0] * T + q[1] * T2 * 0.5 + q[2] * T3 / 3.0 + q[3] * T4 * 0.25 + q[4] * T5 * 0.2 + q[5]);
}

/* Thermo::CalcCp, Thermo.cpp:32-48 (note: >= 1000 here, > 1000 above) */
static double heat_capacity(const par2d_config* c, double T, int* err) {
    const double Rs = c->Rs;
    if (c->igam > 0) return Rs / (1.0 - 1.0 / (c->igam));
    double T2 = T * T, T3 = T * T2, T4 = T * T3;
    if (!(T > 200.0) || !(T < 10000.0)) *err = PAR2D_ERR_THERMO_RANGE;
    const double* q = (T >= 1000.0) ? c->thermo_hi : c->thermo_lo;
    return (Rs) * (q[0] + q[1] * T + q[2] * T2 + q[3] * T3 + q[4] * T4);
}

/* State::UpdateState, StateVariables.h:38-68.  Clamps T in place. */
static void update_state(const par2d_config* c, double* u, orc_derived* d, int* err) {
    const double Rs = c->Rs, gam = fabs(c->igam);
    u[3] = fmax(u[3], 201.0);
    double T = u[3];
    double rho = 0.0;
    rho += u[0];
    d->v2 = u[1] * u[1] + u[2] * u[2];
    d->p = rho * Rs * T;
    d->p = fmax(d->p, 1.0);
    d->a = sqrt(gam * Rs * T);
    d->h = enthalpy(c, T, err);
    d->h0 = d->h + 0.5 * d->v2;
    d->Cp = heat_capacity(c, T, err);
    d->Cv = d->Cp - Rs;
    d->e0 = d->h - Rs * T + 0.5 * d->v2;
    if (!(d->a > 0.0) || !(d->p > 0.0) || isnan(d->p * d->a * T)) *err = PAR2D_ERR_STATE;
}

static double sgn(double v) { return signbit(v) ? -1 : 1; } /* Indexing.h:22 */

/* LDFSS, EulerFlux.cpp:121-192 */
static void ldfss(const par2d_config* c, double normx, double normy, double len, double yface,
                  const double* uL, const orc_derived* L, const double* uR, const orc_derived* R,
                  double* flux, double* parr) {
    double ahalf = 0.5 * (L->a + R->a);
    double xml = (uL[1] * normx + uL[2] * normy) / ahalf;
    double xmr = (uR[1] * normx + uR[2] * normy) / ahalf;

    double all = 0.5 * (1.0 + sgn(xml));
    double alr = 0.5 * (1.0 - sgn(xmr));

    double btl = -fmax(0.0, 1.0 - (double)((int)(fabs(xml))));
    double btr = -fmax(0.0, 1.0 - (double)((int)(fabs(xmr))));

    double xmml = 0.25 * (xml + 1.0) * (xml + 1.0);
    double xmmr = -0.25 * (xmr - 1.0) * (xmr - 1.0);

    double xmhalf = sqrt(0.5 * (xml * xml + xmr * xmr));
    double xmc = 0.25 * btl * btr * (xmhalf - 1.0) * (xmhalf - 1.0);

    double delp = L->p - R->p;
    double psum = L->p + R->p;

    double xmcp = xmc * fmax(0.0, (1.0 - (delp / psum + 2.0 * fabs(delp) / L->p)));
    double xmcm = xmc * fmax(0.0, (1.0 + (delp / psum - 2.0 * fabs(delp) / R->p)));
    double cvlp = all * (1.0 + btl) * xml - btl * xmml;
    double cvlm = alr * (1.0 + btr) * xmr - btr * xmmr;
    double cep = cvlp - xmcp;
    double cem = cvlm + xmcm;

    double fml = len * uL[0] * ahalf * cep;
    double fmr = len * uR[0] * ahalf * cem;

    double ppl = 0.25 * (xml + 1.0) * (xml + 1.0) * (2.0 - xml);
    double ppr = 0.25 * (xmr - 1.0) * (xmr - 1.0) * (2.0 + xmr);

    double pnet = (all * (1.0 + btl) - btl * ppl) * L->p + (alr * (1.0 + btr) - btr * ppr) * R->p;

    double py_planar;
    if (c->iaxi) {
        parr[0] = len * pnet * normy;
        py_planar = 0.0;
    } else {
        yface = 1.0;
        parr[0] = 0.0;
        py_planar = len * pnet * normy;
    }
    flux[0] = yface * (fml + fmr);
    flux[1] = yface * (fml * uL[1] + fmr * uR[1] + len * pnet * normx);
    flux[2] = yface * (fml * uL[2] + fmr * uR[2] + py_planar);
    flux[3] = yface * (fml * L->h0 + fmr * (R->h0));
}

/* SubsonInflo / SubsonOutfl, BoundaryConditions.cpp:12-59.  `far` is the state
 * passed as unkel0; a0/p0 are var0.a / var0.p -- the reference passes the
 * INTERIOR State as var0 at both call sites (BoundaryConditions.cpp:163,181). */
static void subsonic_inflow(const par2d_config* c, double vxi, double vyi, double cint, const double* far,
                            double a0, double p0, double nx, double ny, double* g) {
    const double gam = fabs(c->igam);
    double vin = (vxi * nx) + (vyi * ny);
    double vfn = (far[1] * nx) + (far[2] * ny);
    double vftx = far[1] - vfn * nx;
    double vfty = far[2] - vfn * ny;
    double vgn = 0.5 * (vin + vfn) + (1 / (gam - 1)) * (cint - a0);
    double cg = a0 + 0.5 * (gam - 1) * (vgn - vfn);
    g[0] = far[0] * pow(cg * cg / (a0 * a0), (gam - 1));
    double pg = p0 * pow(g[0] / far[0], gam);
    g[1] = (vftx + vgn * nx);
    g[2] = (vfty + vgn * ny);
    g[3] = pg / (g[0] * c->Rs);
}
static void subsonic_outflow(const par2d_config* c, double rhoi, double pint, double vxi, double vyi, double cint,
                             const double* far, double a0, double nx, double ny, double* g) {
    const double gam = fabs(c->igam);
    double vin = (vxi * nx) + (vyi * ny);
    double vfn = (far[1] * nx) + (far[2] * ny);
    double vitx = vxi - vin * nx;
    double vity = vyi - vin * ny;
    double vgn = 0.5 * (vin + vfn) + (1 / (gam - 1)) * (cint - a0);
    double cg = a0 + 0.5 * (gam - 1) * (vgn - vfn);
    g[0] = rhoi * pow(cg * cg / (cint * cint), (gam - 1));
    double pg = pint * pow(g[0] / rhoi, gam);
    g[1] = (vitx + vgn * nx);
    g[2] = (vity + vgn * ny);
    g[3] = pg / (g[0] * c->Rs);
}

/* boundary_state, BoundaryConditions.cpp:61-190.  g[] is left untouched when
 * the normal Mach number is NaN, like the reference. */
static void boundary_state(const par2d_config* c, int btype, double normx, double normy,
                           const double* ui, const orc_derived* di, double* g) {
    const double* uFS = c->uFS;
    double rhoL = ui[0], uL = ui[1], vL = ui[2];
    double vn = uL * normx + vL * normy;
    double mn = vn / di->a;
    if (btype == -1) btype = 3;
    if (btype == 0) btype = 4;
    if (btype == 4) { /* slip wall / symmetry */
        g[0] = ui[0];
        g[3] = ui[3];
        g[1] = uL - 2 * vn * normx;
        g[2] = vL - 2 * vn * normy;
        return;
    }
    if (!(btype == 1 || btype == 2 || btype == 3 || btype == 5)) return;
    if (btype == 2) { /* fixed back pressure */
        double T2 = ui[3];
        g[0] = c->p_back / (c->Rs * T2);
        g[1] = fmax(0.0, ui[1]);
        g[2] = ui[2];
        g[3] = ui[3];
        return;
    }
    if (btype == 3) { /* extrapolation */
        g[0] = ui[0]; g[1] = ui[1]; g[2] = ui[2]; g[3] = ui[3];
        return;
    }
    if (btype == 5) { /* Bernoulli inflow */
        double pFS = uFS[0] * c->Rs * uFS[3];
        double pB = pFS - 0.5 * ui[0] * di->v2;
        if (ui[1] <= 0.0) pB = pFS;
        g[0] = pB / (c->Rs * uFS[3]);
        g[1] = fmax(ui[1], 0.0);
        g[2] = ui[2];
        g[3] = ui[3];
        return;
    }
    /* btype 1: characteristic far field */
    if (mn <= -1) { g[0] = uFS[0]; g[1] = uFS[1]; g[2] = uFS[2]; g[3] = uFS[3]; return; }
    if (mn <= 0 && mn > -1) { subsonic_inflow(c, uL, vL, di->a, uFS, di->a, di->p, normx, normy, g); return; }
    if (mn >= 1) { g[0] = ui[0]; g[1] = ui[1]; g[2] = ui[2]; g[3] = ui[3]; return; }
    if (mn > 0 && mn < 1) { subsonic_outflow(c, rhoL, di->p, uL, vL, di->a, uFS, di->a, normx, normy, g); return; }
}

/* BuildJacobian (Jacobian.cpp:7-47) + LUPDecompose/LUPSolve (LUtools.cpp:22-90)
 * + axisymmetric scaling (main.cpp:325-331).  Row exchange is done through an
 * index vector; the arithmetic sequence is the reference's. */
static int jacobian_solve(const par2d_config* c, double dt, double ycc, const double* u, const orc_derived* d,
                          const double* b, double* xs) {
    double A[NV][NV];
    int row[NV], P[NV];
    double dti = 1.0 / dt;
    memset(A, 0, sizeof A);
    A[0][0] = dti * 1.0;
    A[1][0] = dti * u[1];
    A[2][0] = dti * u[2];
    A[1][1] = dti * u[0];
    A[2][2] = dti * u[0];
    A[3][0] = dti * d->e0;
    A[3][1] = dti * u[0] * u[1];
    A[3][2] = dti * u[0] * u[2];
    A[3][3] = dti * u[0] * d->Cv;
    for (int i = 0; i < NV; i++) { row[i] = i; P[i] = i; }
    for (int i = 0; i < NV; i++) {
        double big = 0.0; int imax = i;
        for (int k = i; k < NV; k++) {
            double v = fabs(A[row[k]][i]);
            if (v > big) { big = v; imax = k; }
        }
        if (big < 1e-16) return PAR2D_ERR_SINGULAR;
        if (imax != i) {
            int t = P[i]; P[i] = P[imax]; P[imax] = t;
            t = row[i]; row[i] = row[imax]; row[imax] = t;
        }
        for (int j = i + 1; j < NV; j++) {
            A[row[j]][i] /= A[row[i]][i];
            for (int k = i + 1; k < NV; k++) A[row[j]][k] -= A[row[j]][i] * A[row[i]][k];
        }
    }
    for (int i = 0; i < NV; i++) {
        xs[i] = b[P[i]];
        for (int k = 0; k < i; k++) xs[i] -= A[row[i]][k] * xs[k];
    }
    for (int i = NV - 1; i >= 0; i--) {
        for (int k = i + 1; k < NV; k++) xs[i] -= A[row[i]][k] * xs[k];
        xs[i] /= A[row[i]][i];
    }
    if (c->iaxi == 1)
        for (int k = 0; k < NV; k++) xs[k] *= (1.0 / ycc);
    return 0;
}

/* ------------------------------------------------------------------ metrics */

#define CIDX(b, i, j) ((j) * (b)->ncx + (i))
#define NIDX(b, i, j) ((j) * (b)->nx + (i))

/* calc_geoel_geofa, MeshModule.cpp:10-89 */
static void metrics(orc_block* b) {
    const int nx = b->nx, ny = b->ny;
    const double *x = b->x, *y = b->y;
    b->geoel = (double*)malloc(sizeof(double) * 3 * b->ncx * b->ncy);
    b->geofa = (double*)malloc(sizeof(double) * 6 * nx * ny);
    b->yfa = (double*)malloc(sizeof(double) * 2 * nx * ny);
    b->xfa = (double*)malloc(sizeof(double) * 2 * nx * ny);
    for (int k = 0; k < 6 * nx * ny; k++) b->geofa[k] = NAN;
    for (int k = 0; k < 2 * nx * ny; k++) b->yfa[k] = b->xfa[k] = NAN;
    for (int j = 0; j < b->ncy; j++)
        for (int i = 0; i < b->ncx; i++) {
            int p1 = NIDX(b, i, j), p2 = NIDX(b, i + 1, j), p3 = NIDX(b, i + 1, j + 1), p4 = NIDX(b, i, j + 1);
            double vol = 0.5 * (x[p1] * y[p2] + x[p2] * y[p3] + x[p3] * y[p4] + x[p4] * y[p1]
                                - x[p2] * y[p1] - x[p3] * y[p2] - x[p4] * y[p3] - x[p1] * y[p4]);
            double* g = b->geoel + 3 * CIDX(b, i, j);
            g[0] = fabs(vol);
            g[1] = 0.25 * (x[p1] + x[p2] + x[p3] + x[p4]);
            g[2] = 0.25 * (y[p1] + y[p2] + y[p3] + y[p4]);
        }
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            int p0 = NIDX(b, i, j);
            double* f = b->geofa + 6 * p0;
            if (i < nx - 1) { /* face towards +i ("horizontal") */
                int q = NIDX(b, i + 1, j);
                double len = sqrt(((x[q] - x[p0]) * (x[q] - x[p0])) + ((y[q] - y[p0]) * (y[q] - y[p0])));
                f[0] = len;
                f[1] = (y[q] - y[p0]) / len;
                f[2] = -(x[q] - x[p0]) / len;
                b->yfa[2 * p0 + 0] = 0.5 * (y[q] + y[p0]);
                b->xfa[2 * p0 + 0] = 0.5 * (x[q] + x[p0]);
            }
            if (j < ny - 1) { /* face towards +j ("vertical") */
                int q = NIDX(b, i, j + 1);
                double len = sqrt((x[q] - x[p0]) * (x[q] - x[p0]) + (y[q] - y[p0]) * (y[q] - y[p0]));
                f[3] = len;
                f[4] = (y[q] - y[p0]) / len;
                f[5] = -(x[q] - x[p0]) / len;
                b->yfa[2 * p0 + 1] = 0.5 * (y[q] + y[p0]);
                b->xfa[2 * p0 + 1] = 0.5 * (x[q] + x[p0]);
            }
        }
    /* static part of find_dt, main.cpp:26-32 */
    double m = 1.0;
    for (int i = 0; i < nx - 1; i++)
        for (int j = 0; j < ny - 1; j++) {
            m = fmin(m, b->geofa[6 * NIDX(b, i, j) + 0]);
            m = fmin(m, b->geofa[6 * NIDX(b, i, j) + 3]);
        }
    b->min_face = m;
}

/* ------------------------------------------------------- block-level steps */

static int edge_len(const orc_block* b, int side) { return (side == PAR2D_BOTTOM || side == PAR2D_TOP) ? b->ncx : b->ncy; }
static int desc_edge_len(const par2d_block_desc* d, int side) {
    return (side == PAR2D_BOTTOM || side == PAR2D_TOP) ? d->nx - 1 : d->ny - 1;
}

/* cell index of the n-th cell along a side (the line that is sent to a neighbour,
 * SpatialDiscretization.cpp:240,302,362,423) */
static int edge_cell(const orc_block* b, int side, int n) {
    switch (side) {
        case PAR2D_BOTTOM: return CIDX(b, n, 0);
        case PAR2D_RIGHT: return CIDX(b, b->ncx - 1, n);
        case PAR2D_TOP: return CIDX(b, n, b->ncy - 1);
        default: return CIDX(b, 0, n);
    }
}

/* find_dt, main.cpp:19-37: cell (0,0) only */
static void block_dt(orc_ctx* c, orc_block* b) {
    double vmax = sqrt(b->der[0].v2) + b->der[0].a;
    b->dt_b = c->cfg.cfl * b->min_face / vmax;
    if (isnan(b->dt_b)) c->error = c->error ? c->error : PAR2D_ERR_NAN;
}

/* generate_ghost_cells, SpatialDiscretization.cpp:17-99 */
static void ghost_from_bc(orc_ctx* c, orc_block* b) {
    const int nx = b->nx, ny = b->ny;
    for (int i = 0; i < nx - 1; i++) { /* bottom */
        const double* f = b->geofa + 6 * NIDX(b, i, 0);
        int e = CIDX(b, i, 0);
        boundary_state(&c->cfg, b->ibound[i], f[1], f[2], b->unk + NV * e, &b->der[e], b->ghost[PAR2D_BOTTOM] + NV * i);
        update_state(&c->cfg, b->ghost[PAR2D_BOTTOM] + NV * i, &b->gder[PAR2D_BOTTOM][i], &c->error);
    }
    for (int j = 0; j < ny - 1; j++) { /* right: normal read from column 0 (SpatialDiscretization.cpp:51-52) */
        const double* f = b->geofa + 6 * NIDX(b, 0, j);
        int e = CIDX(b, nx - 2, j);
        boundary_state(&c->cfg, b->ibound[j + nx - 1], f[4], f[5], b->unk + NV * e, &b->der[e], b->ghost[PAR2D_RIGHT] + NV * j);
        update_state(&c->cfg, b->ghost[PAR2D_RIGHT] + NV * j, &b->gder[PAR2D_RIGHT][j], &c->error);
    }
    for (int i = 0; i < nx - 1; i++) { /* top */
        const double* f = b->geofa + 6 * NIDX(b, i, ny - 1);
        int e = CIDX(b, i, ny - 2);
        int ib = 2 * (nx - 1) + ny - 1 - i - 1;
        boundary_state(&c->cfg, b->ibound[ib], -f[1], -f[2], b->unk + NV * e, &b->der[e], b->ghost[PAR2D_TOP] + NV * i);
        update_state(&c->cfg, b->ghost[PAR2D_TOP] + NV * i, &b->gder[PAR2D_TOP][i], &c->error);
    }
    for (int j = 0; j < ny - 1; j++) { /* left */
        const double* f = b->geofa + 6 * NIDX(b, 0, j);
        int e = CIDX(b, 0, j);
        int jb = 2 * (ny - 1) + 2 * (nx - 1) - j - 1;
        boundary_state(&c->cfg, b->ibound[jb], -f[4], -f[5], b->unk + NV * e, &b->der[e], b->ghost[PAR2D_LEFT] + NV * j);
        update_state(&c->cfg, b->ghost[PAR2D_LEFT] + NV * j, &b->gder[PAR2D_LEFT][j], &c->error);
    }
}

/* The side of block `nb` that faces block `me` (the line the reference's
 * neighbour rank sends back, SpatialDiscretization.cpp:227-470).  Opposite side
 * preferred; the shipped cases only pair opposite sides. */
static int facing_side(const par2d_block_desc* nb, int me, int my_side) {
    int opp = (my_side + 2) & 3;
    if (nb->bid[opp] == me) return opp;
    for (int s = 0; s < 4; s++)
        if (nb->bid[s] == me) return s;
    return -1;
}

/* interior + boundary faces and the division by the volume,
 * SpatialDiscretization.cpp:481-625, 643-977, 980-1011 */
static void residual(orc_ctx* c, orc_block* b) {
    const par2d_config* cf = &c->cfg;
    const int nx = b->nx, ny = b->ny, axi = cf->iaxi;
    const int ncell = b->ncx * b->ncy;
    double* rhs = b->res; /* accumulate in place, divide at the end */
    double fl[NV], parr;
    for (int k = 0; k < NV * ncell; k++) rhs[k] = 0.0;

    for (int i = 1; i < nx - 1; i++) /* faces between (i-1,j) and (i,j) */
        for (int j = 0; j < ny - 1; j++) {
            const double* f = b->geofa + 6 * NIDX(b, i, j);
            int eL = CIDX(b, i - 1, j), eR = CIDX(b, i, j);
            double rface = axi == 1 ? b->yfa[2 * NIDX(b, i, j) + 1] : 0.0;
            double ycL = axi == 1 ? b->geoel[3 * eL + 2] : 0.0, ycR = axi == 1 ? b->geoel[3 * eR + 2] : 0.0;
            ldfss(cf, f[4], f[5], f[3], rface, b->unk + NV * eL, &b->der[eL], b->unk + NV * eR, &b->der[eR], fl, &parr);
            for (int k = 0; k < NV; k++) {
                if (k == 2) {
                    rhs[NV * eL + k] -= (fl[k] + (ycL * parr));
                    rhs[NV * eR + k] += (fl[k] + (ycR * parr));
                } else {
                    rhs[NV * eL + k] -= fl[k];
                    rhs[NV * eR + k] += fl[k];
                }
            }
        }
    for (int i = 0; i < nx - 1; i++) /* faces between (i,j) [left state] and (i,j-1) */
        for (int j = 1; j < ny - 1; j++) {
            const double* f = b->geofa + 6 * NIDX(b, i, j);
            int eL = CIDX(b, i, j), eR = CIDX(b, i, j - 1);
            double rface = axi ? b->yfa[2 * NIDX(b, i, j) + 0] : 1.0;
            double ycL = axi ? b->geoel[3 * eL + 2] : 0.0, ycR = axi ? b->geoel[3 * eR + 2] : 0.0;
            ldfss(cf, f[1], f[2], f[0], rface, b->unk + NV * eL, &b->der[eL], b->unk + NV * eR, &b->der[eR], fl, &parr);
            for (int k = 0; k < NV; k++) {
                if (k == 2) {
                    rhs[NV * eL + k] -= (fl[k] + (ycL * parr));
                    rhs[NV * eR + k] += (fl[k] + (ycR * parr));
                } else {
                    rhs[NV * eL + k] -= fl[k];
                    rhs[NV * eR + k] += fl[k];
                }
            }
        }
    for (int j = 0; j < ny - 1; j++) {
        { /* left boundary: ghost is the left state */
            const double* f = b->geofa + 6 * NIDX(b, 0, j);
            int e = CIDX(b, 0, j);
            double rface = axi == 1 ? b->yfa[2 * NIDX(b, 0, j) + 1] : 0.0;
            double yc = axi == 1 ? b->geoel[3 * e + 2] : 0.0;
            ldfss(cf, f[4], f[5], f[3], rface, b->ghost[PAR2D_LEFT] + NV * j, &b->gder[PAR2D_LEFT][j], b->unk + NV * e, &b->der[e], fl, &parr);
            for (int k = 0; k < NV; k++) rhs[NV * e + k] += (k == 2) ? (fl[k] + (yc * parr)) : fl[k];
        }
        { /* right boundary: ghost is the right state */
            const double* f = b->geofa + 6 * NIDX(b, nx - 1, j);
            int e = CIDX(b, nx - 2, j);
            double rface = axi ? b->yfa[2 * NIDX(b, nx - 1, j) + 1] : 0.0;
            double yc = axi ? b->geoel[3 * e + 2] : 0.0;
            ldfss(cf, f[4], f[5], f[3], rface, b->unk + NV * e, &b->der[e], b->ghost[PAR2D_RIGHT] + NV * j, &b->gder[PAR2D_RIGHT][j], fl, &parr);
            for (int k = 0; k < NV; k++) rhs[NV * e + k] -= (k == 2) ? (fl[k] + (yc * parr)) : fl[k];
        }
    }
    for (int i = 0; i < nx - 1; i++) {
        { /* bottom boundary: interior is the left state */
            const double* f = b->geofa + 6 * NIDX(b, i, 0);
            int e = CIDX(b, i, 0);
            double rface = axi ? b->yfa[2 * NIDX(b, i, 0) + 0] : 0.0;
            double yc = axi ? b->geoel[3 * e + 2] : 0.0;
            ldfss(cf, f[1], f[2], f[0], rface, b->unk + NV * e, &b->der[e], b->ghost[PAR2D_BOTTOM] + NV * i, &b->gder[PAR2D_BOTTOM][i], fl, &parr);
            for (int k = 0; k < NV; k++) rhs[NV * e + k] -= (k == 2) ? (fl[k] + (yc * parr)) : fl[k];
        }
        { /* top boundary: ghost is the left state */
            const double* f = b->geofa + 6 * NIDX(b, i, ny - 1);
            int e = CIDX(b, i, ny - 2);
            double rface = axi ? b->yfa[2 * NIDX(b, i, ny - 1) + 0] : 0.0;
            double yc = axi ? b->geoel[3 * e + 2] : 0.0;
            ldfss(cf, f[1], f[2], f[0], rface, b->ghost[PAR2D_TOP] + NV * i, &b->gder[PAR2D_TOP][i], b->unk + NV * e, &b->der[e], fl, &parr);
            for (int k = 0; k < NV; k++) rhs[NV * e + k] += (k == 2) ? (fl[k] + (yc * parr)) : fl[k];
        }
    }
    for (int e = 0; e < ncell; e++) {
        double vol = b->geoel[3 * e];
        for (int k = 0; k < NV; k++) rhs[NV * e + k] = rhs[NV * e + k] / vol;
        if (isnan(rhs[NV * e])) c->error = c->error ? c->error : PAR2D_ERR_NAN;
    }
}

/* calculate_residual, main.cpp:39-57 (i outer, j inner) */
static void residual_norms(orc_ctx* c, const orc_block* b, const double* res, double* ressum) {
    double s[NV] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < b->ncx; i++)
        for (int j = 0; j < b->ncy; j++) {
            const double* r = res + NV * CIDX(b, i, j);
            for (int k = 0; k < NV; k++) {
                if (isnan(r[k])) c->error = c->error ? c->error : PAR2D_ERR_NAN;
                s[k] += r[k] * r[k];
            }
        }
    for (int k = 0; k < NV; k++) ressum[k] = sqrt(s[k]);
}

/* main.cpp:307-351: dv = D^-1 res, start-of-step state */
static void solve_block(orc_ctx* c, orc_block* b) {
    double dt = c->dt;
    for (int e = 0; e < b->ncx * b->ncy; e++) {
        int rc = jacobian_solve(&c->cfg, dt, b->geoel[3 * e + 2], b->unk + NV * e, &b->der[e], b->res + NV * e, b->dv + NV * e);
        if (rc) c->error = c->error ? c->error : rc;
    }
}

/* main.cpp:394-404 */
static void update_block(orc_ctx* c, orc_block* b) {
    for (int e = 0; e < b->ncx * b->ncy; e++) {
        double* u = b->unk + NV * e;
        for (int k = 0; k < NV; k++) u[k] += b->dv[NV * e + k];
        if (!(u[0] > 0.0)) c->error = c->error ? c->error : PAR2D_ERR_DENSITY;
        update_state(&c->cfg, u, &b->der[e], &c->error);
    }
}

static void refresh_derived(orc_ctx* c, orc_block* b) {
    for (int e = 0; e < b->ncx * b->ncy; e++) update_state(&c->cfg, b->unk + NV * e, &b->der[e], &c->error);
}

/* ------------------------------------------------ second-order branch */
static int64_t peer_slot_offset(const orc_ctx* c, int peer, int recv_block, int recv_side, int sending);
#include "par2d_oracle_dg.inc"

/* ------------------------------------------------------------ exchange */

/* Fill ghost lines of sides that have a neighbour: same-rank neighbours are
 * read directly, off-rank ones from the peer receive buffer (filled by the
 * host transport between step_begin and step_end).  The received line is
 * copied verbatim and UpdateState'd, SpatialDiscretization.cpp:263-269. */
/* doubles per edge cell of a halo line: NVAR, or 2*NVAR in the second-order branch
 * (SpatialDiscretization.cpp:231 ntrans *= (ACCUR + 1)) */
static int line_width(const par2d_config* cfg) { return cfg->accur == 1 ? 2 * NV : NV; }

static int64_t peer_slot_offset(const orc_ctx* c, int peer, int recv_block, int recv_side, int sending) {
    /* lines are ordered by (receiving block id, receiving side); when `sending`
     * the receiver is the peer's block, otherwise ours */
    int64_t off = 0;
    const int me = c->cfg.rank;
    for (int g = 0; g < c->nblocks; g++) {
        const par2d_block_desc* d = &c->table[g];
        int recv_owner = sending ? peer : me, send_owner = sending ? me : peer;
        if (d->owner != recv_owner) continue;
        for (int s = 0; s < 4; s++) {
            int nb = d->bid[s];
            if (nb < 0 || c->table[nb].owner != send_owner) continue;
            if (g == recv_block && s == recv_side) return off;
            off += (int64_t)line_width(&c->cfg) * desc_edge_len(d, s);
        }
    }
    return -1;
}

static void pack_sends(orc_ctx* c) {
    const int me = c->cfg.rank;
    for (int p = 0; p < c->cfg.nranks; p++) {
        if (p == me || c->peers[p].n_send == 0) continue;
        int64_t off = 0;
        for (int g = 0; g < c->nblocks; g++) { /* receiving blocks on the peer */
            const par2d_block_desc* d = &c->table[g];
            if (d->owner != p) continue;
            for (int s = 0; s < 4; s++) {
                int nb = d->bid[s];
                if (nb < 0 || c->table[nb].owner != me) continue;
                orc_block* mine = &c->blk[c->owned_index[nb]];
                int ms = facing_side(&c->table[nb], g, s);
                int n = desc_edge_len(d, s);
                if (c->cfg.accur == 1) dg_pack_line(c, mine, ms, n, c->peers[p].send + off);
                else
                    for (int q = 0; q < n; q++)
                        memcpy(c->peers[p].send + off + NV * q, mine->unk + NV * edge_cell(mine, ms, q), sizeof(double) * NV);
                off += (int64_t)line_width(&c->cfg) * n;
            }
        }
    }
}

static void ghost_from_neighbours(orc_ctx* c, orc_block* b) {
    const int me = c->cfg.rank;
    for (int s = 0; s < 4; s++) {
        int nb = b->bid[s];
        if (nb < 0) continue;
        int n = edge_len(b, s);
        if (c->table[nb].owner == me) {
            orc_block* o = &c->blk[c->owned_index[nb]];
            int os = facing_side(&c->table[nb], b->gid, s);
            for (int q = 0; q < n; q++) memcpy(b->ghost[s] + NV * q, o->unk + NV * edge_cell(o, os, q), sizeof(double) * NV);
        } else {
            int p = c->table[nb].owner;
            int64_t off = peer_slot_offset(c, p, b->gid, s, 0);
            memcpy(b->ghost[s], c->peers[p].recv + off, sizeof(double) * NV * n);
        }
        for (int q = 0; q < n; q++) update_state(&c->cfg, b->ghost[s] + NV * q, &b->gder[s][q], &c->error);
    }
}

/* ------------------------------------------------------------ life cycle */

static int validate_table(int nblocks, const par2d_block_desc* t, int nranks) {
    for (int g = 0; g < nblocks; g++) {
        if (t[g].nx < 2 || t[g].ny < 2) return fail(PAR2D_ERR_ARG, "block with fewer than 2 nodes per direction");
        if (t[g].owner < 0 || t[g].owner >= nranks) return fail(PAR2D_ERR_ARG, "block owner out of range");
        for (int s = 0; s < 4; s++) {
            int nb = t[g].bid[s];
            if (nb < 0) continue;
            if (nb >= nblocks || nb == g) return fail(PAR2D_ERR_ARG, "neighbour id out of range");
            int os = facing_side(&t[nb], g, s);
            if (os < 0) return fail(PAR2D_ERR_ARG, "neighbour does not point back");
            if (desc_edge_len(&t[g], s) != desc_edge_len(&t[nb], os)) return fail(PAR2D_ERR_ARG, "shared edge lengths differ");
        }
    }
    return 0;
}

int orc_create(const par2d_config* cfg, int32_t nblocks, const par2d_block_desc* blocks, orc_ctx** out) {
    if (!cfg || !blocks || !out || nblocks < 1) return fail(PAR2D_ERR_ARG, "null argument");
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail(PAR2D_ERR_ARG, "bad rank");
    if (cfg->accur != 0 && cfg->accur != 1) return fail(PAR2D_ERR_ARG, "accur must be 0 or 1 (SpatialDiscretization.cpp:204)");
    int rc = validate_table(nblocks, blocks, cfg->nranks);
    if (rc) return rc;
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof *c);
    c->cfg = *cfg;
    c->nblocks = nblocks;
    c->table = (par2d_block_desc*)malloc(sizeof(par2d_block_desc) * nblocks);
    memcpy(c->table, blocks, sizeof(par2d_block_desc) * nblocks);
    c->owned_index = (int*)malloc(sizeof(int) * nblocks);
    c->dt_table = (double*)calloc(nblocks, sizeof(double));
    for (int g = 0; g < nblocks; g++) {
        c->table[g].x = c->table[g].y = NULL; c->table[g].ibound = NULL;
        c->owned_index[g] = -1;
        if (blocks[g].owner == cfg->rank) c->nowned++;
    }
    c->blk = (orc_block*)calloc(c->nowned ? c->nowned : 1, sizeof(orc_block));
    int k = 0;
    for (int g = 0; g < nblocks; g++) {
        if (blocks[g].owner != cfg->rank) continue;
        const par2d_block_desc* d = &blocks[g];
        if (!d->x || !d->y || !d->ibound) return fail(PAR2D_ERR_ARG, "owned block without coordinates");
        orc_block* b = &c->blk[k];
        c->owned_index[g] = k++;
        b->gid = g; b->nx = d->nx; b->ny = d->ny; b->ncx = d->nx - 1; b->ncy = d->ny - 1;
        memcpy(b->bid, d->bid, sizeof b->bid);
        size_t np = (size_t)d->nx * d->ny, nc = (size_t)b->ncx * b->ncy;
        int nb = 2 * b->ncx + 2 * b->ncy;
        b->x = (double*)malloc(sizeof(double) * np); memcpy(b->x, d->x, sizeof(double) * np);
        b->y = (double*)malloc(sizeof(double) * np); memcpy(b->y, d->y, sizeof(double) * np);
        b->ibound = (int*)malloc(sizeof(int) * nb); memcpy(b->ibound, d->ibound, sizeof(int) * nb);
        metrics(b);
        b->unk = (double*)malloc(sizeof(double) * NV * nc);
        b->res = (double*)calloc(NV * nc, sizeof(double));
        b->dv = (double*)calloc(NV * nc, sizeof(double));
        b->der = (orc_derived*)malloc(sizeof(orc_derived) * nc);
        for (int s = 0; s < 4; s++) {
            int n = edge_len(b, s);
            b->ghost[s] = (double*)calloc(NV * n, sizeof(double));
            b->gder[s] = (orc_derived*)calloc(n, sizeof(orc_derived));
        }
        if (cfg->accur == 1) {
            b->ux = (double*)calloc(NV * nc, sizeof(double)); b->uy = (double*)calloc(NV * nc, sizeof(double));  /* main.cpp:199-209 */
            b->resx = (double*)calloc(NV * nc, sizeof(double)); b->resy = (double*)calloc(NV * nc, sizeof(double));
            b->dvx = (double*)calloc(NV * nc, sizeof(double)); b->dvy = (double*)calloc(NV * nc, sizeof(double));
            for (int s = 0; s < 4; s++) {
                int n = edge_len(b, s);
                b->ghost2[s] = (double*)calloc(2 * NV * n, sizeof(double));
                b->gder2[s] = (orc_derived*)calloc(2 * n, sizeof(orc_derived));
            }
        }
        for (size_t e = 0; e < nc; e++) memcpy(b->unk + NV * e, cfg->uFS, sizeof(double) * NV); /* main.cpp:193-197 */
        refresh_derived(c, b);
    }
    c->peers = (orc_peer*)calloc(cfg->nranks, sizeof(orc_peer));
    for (int p = 0; p < cfg->nranks; p++) {
        c->peers[p].peer = p;
        if (p == cfg->rank) continue;
        for (int g = 0; g < nblocks; g++)
            for (int s = 0; s < 4; s++) {
                int nb = blocks[g].bid[s];
                if (nb < 0) continue;
                if (blocks[g].owner == cfg->rank && blocks[nb].owner == p) c->peers[p].n_recv += (int64_t)line_width(cfg) * desc_edge_len(&blocks[g], s);
                if (blocks[g].owner == p && blocks[nb].owner == cfg->rank) c->peers[p].n_send += (int64_t)line_width(cfg) * desc_edge_len(&blocks[g], s);
            }
        c->peers[p].send = (double*)calloc(c->peers[p].n_send ? c->peers[p].n_send : 1, sizeof(double));
        c->peers[p].recv = (double*)calloc(c->peers[p].n_recv ? c->peers[p].n_recv : 1, sizeof(double));
    }
    c->error = 0; /* uniform free stream is always a valid state */
    *out = c;
    return 0;
}

int orc_destroy(orc_ctx* c) {
    if (!c) return 0;
    for (int k = 0; k < c->nowned; k++) {
        orc_block* b = &c->blk[k];
        free(b->x); free(b->y); free(b->ibound); free(b->geoel); free(b->geofa); free(b->yfa);
        free(b->unk); free(b->res); free(b->dv); free(b->der); free(b->xfa);
        free(b->ux); free(b->uy); free(b->resx); free(b->resy); free(b->dvx); free(b->dvy);
        for (int s = 0; s < 4; s++) { free(b->ghost[s]); free(b->gder[s]); free(b->ghost2[s]); free(b->gder2[s]); }
    }
    for (int p = 0; p < c->cfg.nranks; p++) { free(c->peers[p].send); free(c->peers[p].recv); }
    free(c->peers); free(c->blk); free(c->table); free(c->owned_index); free(c->dt_table);
    free(c->hist); free(c->blk_hist); free(c);
    return 0;
}

static orc_block* owned(orc_ctx* c, int block) {
    if (!c || block < 0 || block >= c->nblocks || c->owned_index[block] < 0) return NULL;
    return &c->blk[c->owned_index[block]];
}

int orc_set_state(orc_ctx* c, int32_t block, const double* unk) {
    orc_block* b = owned(c, block);
    if (!b || !unk) return fail(PAR2D_ERR_ARG, "set_state: block not owned");
    memcpy(b->unk, unk, sizeof(double) * NV * b->ncx * b->ncy);
    int keep = c->error;
    refresh_derived(c, b); /* State::UpdateState after Initialize, main.cpp:226-238 */
    (void)keep;
    return 0;
}
int orc_get_state(orc_ctx* c, int32_t block, double* unk) {
    orc_block* b = owned(c, block);
    if (!b || !unk) return fail(PAR2D_ERR_ARG, "get_state: block not owned");
    memcpy(unk, b->unk, sizeof(double) * NV * b->ncx * b->ncy);
    return 0;
}

int64_t orc_iteration(const orc_ctx* c) { return c->iter; }
int32_t orc_num_owned(const orc_ctx* c) { return c->nowned; }
int32_t orc_owned_block(const orc_ctx* c, int32_t k) { return (k >= 0 && k < c->nowned) ? c->blk[k].gid : -1; }
int64_t orc_owned_cells(const orc_ctx* c) {
    int64_t n = 0;
    for (int k = 0; k < c->nowned; k++) n += (int64_t)c->blk[k].ncx * c->blk[k].ncy;
    return n;
}

int orc_exchange_layout(orc_ctx* c, int32_t peer, int64_t* n_send, int64_t* n_recv, void** send, void** recv) {
    if (!c || peer < 0 || peer >= c->cfg.nranks) return fail(PAR2D_ERR_ARG, "bad peer");
    if (n_send) *n_send = c->peers[peer].n_send;
    if (n_recv) *n_recv = c->peers[peer].n_recv;
    if (send) *send = c->peers[peer].send;
    if (recv) *recv = c->peers[peer].recv;
    return 0;
}
int orc_dt_table(orc_ctx* c, void** table, int32_t* n) {
    if (table) *table = c->dt_table;
    if (n) *n = c->nblocks;
    return 0;
}

/* ------------------------------------------------------------ stepping */

static void ensure_hist(orc_ctx* c, int n) {
    if (n <= c->hist_cap) return;
    c->hist = (par2d_step_info*)realloc(c->hist, sizeof(par2d_step_info) * n);
    c->blk_hist = (double*)realloc(c->blk_hist, sizeof(double) * 2 * n * (c->nowned ? c->nowned : 1));
    c->hist_cap = n;
}

int orc_step_begin(orc_ctx* c) {
    memset(c->dt_table, 0, sizeof(double) * c->nblocks);
    for (int k = 0; k < c->nowned; k++) {
        block_dt(c, &c->blk[k]);
        c->dt_table[c->blk[k].gid] = c->blk[k].dt_b;
    }
    pack_sends(c);
    return 0;
}

/* everything after the dt sync of one iteration, main.cpp:293-531 */
int orc_step_end(orc_ctx* c) {
    double dt = 0.0;
    int first = 1;
    for (int g = 0; g < c->nblocks; g++) { /* main.cpp:270-280: rank order, fmax */
        if (first) { dt = c->dt_table[g]; first = 0; }
        else dt = fmax(dt, c->dt_table[g]);
    }
    c->dt = dt;
    const int dg = c->cfg.accur == 1;
#pragma omp parallel for schedule(dynamic)
    for (int k = 0; k < c->nowned; k++) { if (dg) dg_ghost_from_bc(c, &c->blk[k]); else ghost_from_bc(c, &c->blk[k]); }
#pragma omp parallel for schedule(dynamic)
    for (int k = 0; k < c->nowned; k++) { if (dg) dg_ghost_from_neighbours(c, &c->blk[k]); else ghost_from_neighbours(c, &c->blk[k]); }
#pragma omp parallel for schedule(dynamic)
    for (int k = 0; k < c->nowned; k++) {
        if (dg) { dg_residual(c, &c->blk[k]); dg_solve_block(c, &c->blk[k]); }
        else { residual(c, &c->blk[k]); solve_block(c, &c->blk[k]); }
    }
    ensure_hist(c, c->hist_len + 1);
    par2d_step_info* h = &c->hist[c->hist_len];
    h->iter = c->iter; h->dt = dt; h->ressumsum = 0.0; h->res0sum = 0.0;
    for (int k = 0; k < c->nowned; k++) { /* main.cpp:365-392 */
        orc_block* b = &c->blk[k];
        double ressum[NV];
        residual_norms(c, b, b->res, ressum);
        if (dg) { /* main.cpp:300-302, 360-364 */
            double rx[NV], ry[NV];
            residual_norms(c, b, b->resx, rx);
            residual_norms(c, b, b->resy, ry);
            for (int v = 0; v < NV; v++) ressum[v] += rx[v] + ry[v];
        }
        double a = 0.0, z = 0.0;
        for (int v = 0; v < NV; v++) {
            if (c->iter == 0) b->res0[v] = ressum[v];
            if (b->res0[v] < 1e-16) b->res0[v] = fmax(ressum[v], 1e-16);
            a += ressum[v];
            z += b->res0[v];
        }
        c->blk_hist[2 * (c->hist_len * c->nowned + k) + 0] = a;
        c->blk_hist[2 * (c->hist_len * c->nowned + k) + 1] = z;
        h->ressumsum += a;
        h->res0sum += z;
    }
    h->relres = h->ressumsum / h->res0sum;
#pragma omp parallel for schedule(dynamic)
    for (int k = 0; k < c->nowned; k++) { if (dg) dg_update_block(c, &c->blk[k]); else update_block(c, &c->blk[k]); }
    h->status = c->error;
    h->reserved = 0;
    c->hist_len++;
    c->iter++;
    return 0;
}

int orc_step_collect(orc_ctx* c, int32_t nsteps, par2d_step_info* info) {
    if (nsteps > c->hist_len) return fail(PAR2D_ERR_ARG, "collect: more steps than recorded");
    if (info) memcpy(info, c->hist + (c->hist_len - nsteps), sizeof(par2d_step_info) * nsteps);
    return c->error;
}

int orc_step(orc_ctx* c, int32_t nsteps, int32_t stop_on_tol, par2d_step_info* info) {
    if (!c || nsteps < 0) return fail(PAR2D_ERR_ARG, "step: bad argument");
    if (c->cfg.nranks != 1) return fail(PAR2D_ERR_COMM, "step: use step_begin/step_end with nranks > 1");
    c->hist_len = 0;
    int done = 0;
    for (; done < nsteps; done++) {
        orc_step_begin(c);
        orc_step_end(c);
        if (stop_on_tol && c->hist[done].relres < c->cfg.tol) { done++; break; }
        if (c->error) { done++; break; }
    }
    if (info) {
        memcpy(info, c->hist, sizeof(par2d_step_info) * done);
        for (int s = done; s < nsteps; s++) { memset(&info[s], 0, sizeof info[s]); info[s].iter = -1; }
    }
    return c->error;
}

int orc_get_block_history(orc_ctx* c, int32_t nsteps, double* out) {
    if (nsteps > c->hist_len) return fail(PAR2D_ERR_ARG, "history: more steps than recorded");
    memcpy(out, c->blk_hist + 2 * (size_t)(c->hist_len - nsteps) * c->nowned, sizeof(double) * 2 * nsteps * c->nowned);
    return 0;
}
void orc_reset_history(orc_ctx* c) { c->hist_len = 0; }

/* one calc_dudt evaluation on the current state; nranks == 1 only */
int orc_get_residual(orc_ctx* c, int32_t block, double* res) {
    orc_block* b = owned(c, block);
    if (!b || !res) return fail(PAR2D_ERR_ARG, "get_residual: block not owned");
    if (c->cfg.nranks != 1) return fail(PAR2D_ERR_COMM, "get_residual: single rank only in the oracle");
    for (int k = 0; k < c->nowned; k++) ghost_from_bc(c, &c->blk[k]);
    for (int k = 0; k < c->nowned; k++) ghost_from_neighbours(c, &c->blk[k]);
    residual(c, b);
    memcpy(res, b->res, sizeof(double) * NV * b->ncx * b->ncy);
    return 0;
}

/* second-order branch: slopes in / out and the three right-hand sides of one calc_dudt evaluation */
int orc_set_slopes(orc_ctx* c, int32_t block, const double* ux, const double* uy) {
    orc_block* b = owned(c, block);
    if (!b || !ux || !uy || c->cfg.accur != 1) return fail(PAR2D_ERR_ARG, "set_slopes: block not owned or accur != 1");
    memcpy(b->ux, ux, sizeof(double) * NV * b->ncx * b->ncy);
    memcpy(b->uy, uy, sizeof(double) * NV * b->ncx * b->ncy);
    return 0;
}
int orc_get_slopes(orc_ctx* c, int32_t block, double* ux, double* uy) {
    orc_block* b = owned(c, block);
    if (!b || !ux || !uy || c->cfg.accur != 1) return fail(PAR2D_ERR_ARG, "get_slopes: block not owned or accur != 1");
    memcpy(ux, b->ux, sizeof(double) * NV * b->ncx * b->ncy);
    memcpy(uy, b->uy, sizeof(double) * NV * b->ncx * b->ncy);
    return 0;
}
int orc_get_residual_dg(orc_ctx* c, int32_t block, double* res, double* resx, double* resy) {
    orc_block* b = owned(c, block);
    if (!b || !res || !resx || !resy || c->cfg.accur != 1) return fail(PAR2D_ERR_ARG, "get_residual_dg: block not owned or accur != 1");
    for (int k = 0; k < c->nowned; k++) dg_ghost_from_bc(c, &c->blk[k]);
    for (int k = 0; k < c->nowned; k++) dg_ghost_from_neighbours(c, &c->blk[k]);
    dg_residual(c, b);
    memcpy(res, b->res, sizeof(double) * NV * b->ncx * b->ncy);
    memcpy(resx, b->resx, sizeof(double) * NV * b->ncx * b->ncy);
    memcpy(resy, b->resy, sizeof(double) * NV * b->ncx * b->ncy);
    return 0;
}

/* metrics of an owned block in the reference layouts (test hook) */
int orc_get_metrics(orc_ctx* c, int32_t block, double* geoel, double* geofa, double* yfa) {
    orc_block* b = owned(c, block);
    if (!b) return fail(PAR2D_ERR_ARG, "get_metrics: block not owned");
    if (geoel) memcpy(geoel, b->geoel, sizeof(double) * 3 * b->ncx * b->ncy);
    if (geofa) memcpy(geofa, b->geofa, sizeof(double) * 6 * b->nx * b->ny);
    if (yfa) memcpy(yfa, b->yfa, sizeof(double) * 2 * b->nx * b->ny);
    return 0;
}

/* ------------------------------------------------- point-wise test hooks */

int orc_test_ldfss(const par2d_config* cfg, int64_t n, const double* face, const double* uL, const double* uR,
                   double* flux, double* parr) {
    int err = 0;
    for (int64_t i = 0; i < n; i++) {
        double l[NV], r[NV];
        orc_derived dl, dr;
        memcpy(l, uL + NV * i, sizeof l); memcpy(r, uR + NV * i, sizeof r);
        update_state(cfg, l, &dl, &err); update_state(cfg, r, &dr, &err);
        ldfss(cfg, face[4 * i], face[4 * i + 1], face[4 * i + 2], face[4 * i + 3], l, &dl, r, &dr, flux + NV * i, parr + i);
    }
    return 0;
}
int orc_test_boundary_state(const par2d_config* cfg, int64_t n, const int32_t* btype, const double* norm,
                            const double* uInt, double* uGhost) {
    int err = 0;
    for (int64_t i = 0; i < n; i++) {
        double u[NV];
        orc_derived d, dg;
        memcpy(u, uInt + NV * i, sizeof u);
        update_state(cfg, u, &d, &err);
        boundary_state(cfg, btype[i], norm[2 * i], norm[2 * i + 1], u, &d, uGhost + NV * i);
        update_state(cfg, uGhost + NV * i, &dg, &err);
    }
    return 0;
}
int orc_test_update_state(const par2d_config* cfg, int64_t n, double* unk, double* derived) {
    int err = 0;
    for (int64_t i = 0; i < n; i++) {
        orc_derived d;
        update_state(cfg, unk + NV * i, &d, &err);
        double* o = derived + 8 * i;
        o[0] = d.p; o[1] = d.a; o[2] = d.h; o[3] = d.h0; o[4] = d.e0; o[5] = d.v2; o[6] = d.Cp; o[7] = d.Cv;
    }
    return 0;
}
int orc_test_solve(const par2d_config* cfg, int64_t n, const double* dt_ycc, const double* unk, const double* res, double* dv) {
    int err = 0;
    for (int64_t i = 0; i < n; i++) {
        double u[NV];
        orc_derived d;
        memcpy(u, unk + NV * i, sizeof u);
        update_state(cfg, u, &d, &err);
        int rc = jacobian_solve(cfg, dt_ycc[2 * i], dt_ycc[2 * i + 1], u, &d, res + NV * i, dv + NV * i);
        if (rc) return rc;
    }
    return 0;
}
