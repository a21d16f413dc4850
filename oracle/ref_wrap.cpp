/* oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * extern "C" entry points over the UNMODIFIED reference functions, linked into
 * oracle/_ref/libpar2d_ref.so together with the reference's own translation
 * units (everything except main.cpp).  Lets the tests feed arbitrary inputs to
 *   LDFSS            /root/reference/FV/EulerFlux.cpp:121
 *   boundary_state   /root/reference/FV/BoundaryConditions.cpp:61
 *   State::UpdateState  /root/reference/FV/StateVariables.h:38
 *   BuildJacobian + LUPDecompose + LUPSolve  Jacobian.cpp:7, LUtools.cpp:22,75
 *   calc_geoel_geofa /root/reference/FV/MeshModule.cpp:10
 *   calc_dudt (single block, no exchange)  SpatialDiscretization.cpp:159
 * and compare against oracle/par2d_oracle.c (the restatement) and the CUDA path.
 * No reference source text lives in this file; it only calls into it.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include "Indexing.h"
#include "Thermo.h"
#include "StateVariables.h"
#include "EulerFlux.h"
#include "BoundaryConditions.h"
#include "SpatialDiscretization.h"
#include "Jacobian.h"
#include "LUtools.h"
#include "MeshModule.h"

static Thermo* g_air = nullptr;

extern "C" {

/* The reference Thermo constructor opens "../therm_air.dat" relative to the
 * working directory (Thermo.h:19); run_dir must contain a sub-directory whose
 * parent holds that file, exactly like the reference's build directory. */
int ref_init(const char* build_dir, int iaxi, double igam) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return 1;
    if (chdir(build_dir) != 0) return 2;
    oracle_ref_config().iaxi = iaxi;
    oracle_ref_config().igam = igam;
    delete g_air;
    g_air = new Thermo();
    if (chdir(cwd) != 0) return 3;
    return 0;
}

double ref_Rs(void) { return g_air->Rs[0]; }
double ref_gam(void) { return g_air->gam; }

/* derived[] = p, a, h, h0, e0, v2, Cp, Cv ; unk[3] may be clamped in place. */
void ref_update_state(double* unk, double* derived) {
    State s; s.Initialize(unk); s.UpdateState(*g_air);
    derived[0] = s.p; derived[1] = s.a; derived[2] = s.h; derived[3] = s.h0;
    derived[4] = s.e0; derived[5] = s.v2; derived[6] = s.Cp[0]; derived[7] = s.Cv;
}

void ref_ldfss(double normx, double normy, double len, double yface,
               double* uL, double* uR, double* flux, double* parr) {
    State sl, sr;
    sl.Initialize(uL); sl.UpdateState(*g_air);
    sr.Initialize(uR); sr.UpdateState(*g_air);
    LDFSS(normx, normy, len, yface, uL, sl, uR, sr, flux, parr);
}

void ref_boundary_state(int btype, double normx, double normy, const double* uFS,
                        double* uInt, double* uGhost) {
    State s; s.Initialize(uInt); s.UpdateState(*g_air);
    boundary_state(btype, *g_air, normx, normy, uFS, uInt, s, uGhost);
}

/* dv = D^-1 res with D from BuildJacobian (pivoted LU), as main.cpp:307-323. */
void ref_jacobian_solve(double dt, double* unk, const double* res, double* dv) {
    State s; s.Initialize(unk); s.UpdateState(*g_air);
    double rows[NVAR][NVAR]; double* D[NVAR];
    for (int k = 0; k < NVAR; k++) D[k] = rows[k];
    int P[NVAR + 1] = {0};
    double b[NVAR]; memcpy(b, res, sizeof b);
    BuildJacobian(dt, unk, s, D);
    LUPDecompose(D, NVAR, 1e-16, P);
    LUPSolve(D, P, b, NVAR, dv);
}

/* Metrics into caller-provided arrays: geoel[3*nelem], geofa[6*npoin], yfa/xfa[2*npoin]. */
void ref_metrics(int nx, int ny, double* x, double* y,
                 double* geoel, double* geofa, double* yfa, double* xfa) {
    double *ge, *gf, *yf, *xf;
    calc_geoel_geofa(nx, ny, x, y, &ge, &gf, &yf, &xf);
    int nelem = (nx - 1) * (ny - 1), npoin = nx * ny;
    memcpy(geoel, ge, sizeof(double) * 3 * nelem);
    memcpy(geofa, gf, sizeof(double) * 6 * npoin);
    memcpy(yfa, yf, sizeof(double) * 2 * npoin);
    memcpy(xfa, xf, sizeof(double) * 2 * npoin);
    free(ge); free(gf); free(yf); free(xf);
}

/* One residual evaluation on a single block with no neighbours (world_size==1
 * path of calc_dudt: SpatialDiscretization.cpp:218). unk may get its T clamped. */
void ref_calc_dudt(int nx, int ny, double* x, double* y, int* ibound, double* uFS,
                   double* unk, double* dudt) {
    double *geoel, *geofa, *yfa, *xfa;
    calc_geoel_geofa(nx, ny, x, y, &geoel, &geofa, &yfa, &xfa);
    int nelem = (nx - 1) * (ny - 1);
    State* ev = (State*)malloc(sizeof(State) * nelem);
    for (int e = 0; e < nelem; e++) { ev[e] = State(); ev[e].Initialize(unk + NVAR * e); ev[e].UpdateState(*g_air); }
    int bbounds[4] = {0, 0, 0, 0}, bids[4] = {-1, -1, -1, -1};
    calc_dudt(bbounds, bids, nx, ny, *g_air, ev, uFS, ibound, geoel, geofa, yfa, xfa,
              unk, nullptr, nullptr, dudt, nullptr, nullptr);
    free(ev); free(geoel); free(geofa); free(yfa); free(xfa);
}

/* The same with the second-order branch switched on (ACCUR == 1): dudt, duxdt, duydt of one calc_dudt
 * evaluation for given means and slopes (SpatialDiscretization.cpp:159-1050, DGP1Tools.cpp). */
void ref_calc_dudt_dg(int nx, int ny, double* x, double* y, int* ibound, double* uFS,
                      double* unk, double* ux, double* uy, double* dudt, double* duxdt, double* duydt) {
    const int keep = oracle_ref_config().accur;
    oracle_ref_config().accur = 1;
    double *geoel, *geofa, *yfa, *xfa;
    calc_geoel_geofa(nx, ny, x, y, &geoel, &geofa, &yfa, &xfa);
    int nelem = (nx - 1) * (ny - 1);
    State* ev = (State*)malloc(sizeof(State) * nelem);
    for (int e = 0; e < nelem; e++) { ev[e] = State(); ev[e].Initialize(unk + NVAR * e); ev[e].UpdateState(*g_air); }
    int bbounds[4] = {0, 0, 0, 0}, bids[4] = {-1, -1, -1, -1};
    calc_dudt(bbounds, bids, nx, ny, *g_air, ev, uFS, ibound, geoel, geofa, yfa, xfa,
              unk, ux, uy, dudt, duxdt, duydt);
    free(ev); free(geoel); free(geofa); free(yfa); free(xfa);
    oracle_ref_config().accur = keep;
}

}  /* extern "C" */
