/* oracle/ref_runtime_config.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * The reference solver fixes its physics switches at compile time
 * (/root/reference/FV/Indexing.h:16 IGAM, :18 IAXI, :20 NITER).  Every use of
 * those three names is an ordinary run-time expression (Thermo.cpp:14,34;
 * Thermo.h:47; EulerFlux.cpp:179; SpatialDiscretization.cpp:491,565,665,752,
 * 829,910; main.cpp:89,326), so oracle/build_ref.sh generates -- at build time,
 * into the git-ignored oracle/_ref/ -- a copy of the reference header in which
 * just those three #defines read this struct instead.  One oracle binary then
 * serves the planar/axisymmetric and CPG/TPG variants:
 *     ORACLE_NITER (default 10)  ORACLE_IAXI (default 1)  ORACLE_IGAM (default 1.4)  ORACLE_ACCUR (default 0)
 * A second, "strict" binary whose header differs from the reference's only in
 * the NITER literal is built alongside and must produce bitwise identical
 * solution files (checked by oracle/build_ref.sh verify).
 */
#pragma once
#include <cstdlib>

struct OracleRefConfig {
    int niter;
    int iaxi;
    int accur;   /* Indexing.h:17 ACCUR: every use is a run-time expression as well (if (ACCUR == 1), ntrans *= (ACCUR + 1)) */
    double igam;
    OracleRefConfig() {
        const char* s;
        niter = (s = getenv("ORACLE_NITER")) ? atoi(s) : 10;
        iaxi  = (s = getenv("ORACLE_IAXI"))  ? atoi(s) : 1;
        igam  = (s = getenv("ORACLE_IGAM"))  ? atof(s) : 1.4;
        accur = (s = getenv("ORACLE_ACCUR")) ? atoi(s) : 0;
    }
};
inline OracleRefConfig& oracle_ref_config() { static OracleRefConfig c; return c; }
