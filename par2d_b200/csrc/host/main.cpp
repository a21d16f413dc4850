// main.cpp -- par2d_fv: the reference driver (FV/main.cpp) re-expressed over the C ABI.
//
// Same inputs and outputs as the reference executable MPI2DFV, with its hard-coded paths
// and compile-time switches turned into arguments whose defaults are the shipped values:
//   inputs : <case>/grid.conn + <case>/subgrids/grid.block<b>.dat  (FV/main.cpp:114-142)
//            <therm>  NASA-7 record                                 (FV/Thermo.h:16-34)
//   outputs: stdout line per iteration                              (FV/main.cpp:499)
//            <out>/res.tec                                          (FV/main.cpp:245-254,495)
//            <out>/solution.b<b+1>.tec every saveiter iterations and at the end
//                                                                   (FV/main.cpp:484-491,541-545)
// All blocks are advanced by one process on one GPU; the multi-GPU launch lives in
// par2d_b200/dist.py (one process per GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "par2d_gpu.h"
#include "par2d_host.h"

static void usage() {
    printf("usage: par2d_fv --case DIR [--out DIR] [--therm FILE] [--niter N] [--iaxi 0|1] [--igam G]\n"
           "                [--cfl C] [--tol T] [--printiter N] [--saveiter N] [--batch N] [--faithful]\n"
           "                [--restart DIR]   start from DIR/solution.b<b+1>.tec instead of the uniform state\n"
           "                [--accur 0|1]     1: the DG(P1) second-order branch (ACCUR, FV/Indexing.h:17)\n");
}

int main(int argc, char** argv) {
    std::string case_dir = "../Case", out_dir = "../Outputs", therm = "../therm_air.dat", restart_dir;
    long niter = 5000000;  // NITER, FV/Indexing.h:20
    int printiter = 1, saveiter = 1000, batch = 100;
    par2d_config cfg;
    par2d_default_config(&cfg);
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
        if (a == "--case") case_dir = next();
        else if (a == "--out") out_dir = next();
        else if (a == "--therm") therm = next();
        else if (a == "--niter") niter = atol(next());
        else if (a == "--iaxi") cfg.iaxi = atoi(next());
        else if (a == "--igam") cfg.igam = atof(next());
        else if (a == "--cfl") cfg.cfl = atof(next());
        else if (a == "--tol") cfg.tol = atof(next());
        else if (a == "--printiter") printiter = atoi(next());
        else if (a == "--saveiter") saveiter = atoi(next());
        else if (a == "--batch") batch = atoi(next());
        else if (a == "--faithful") cfg.mode = PAR2D_MODE_FAITHFUL;
        else if (a == "--restart") restart_dir = next();
        else if (a == "--accur") cfg.accur = atoi(next());
        else { usage(); return 2; }
    }
    if (batch < 1) batch = 1;

    printf("==================== Loading Mesh ====================\n");
    par2d_case* cs = nullptr;
    if (par2d_case_load(case_dir.c_str(), &cs)) { printf("%s\n", par2d_host_last_error()); return 1; }
    if (par2d_thermo_load(therm.c_str(), &cfg)) { printf("%s\n", par2d_host_last_error()); return 1; }
    // free stream from the loaded gas constant, FV/main.cpp:84-88
    cfg.uFS[0] = 1000.0 / (cfg.Rs * 300.0);
    const int nb = par2d_case_num_blocks(cs);
    std::vector<par2d_block_desc> blocks(nb);
    for (int b = 0; b < nb; b++) par2d_case_block(cs, b, &blocks[b]);

    printf("==================== Initializing ====================\n");
    par2d_ctx* ctx = nullptr;
    if (par2d_create(&cfg, nb, blocks.data(), &ctx)) { printf("par2d_create: %s\n", par2d_last_error()); return 1; }

    if (!restart_dir.empty()) {
        // restart (the reference writes solution.b<N>.tec every saveiter iterations but cannot read them back)
        for (int b = 0; b < nb; b++) {
            std::vector<double> unk((size_t)4 * (blocks[b].nx - 1) * (blocks[b].ny - 1));
            const std::string path = restart_dir + "/solution.b" + std::to_string(b + 1) + ".tec";
            if (par2d_read_solution(path.c_str(), blocks[b].nx, blocks[b].ny, unk.data(), nullptr, nullptr)) { printf("%s\n", par2d_host_last_error()); return 1; }
            if (par2d_set_state(ctx, b, unk.data())) { printf("set_state: %s\n", par2d_last_error()); return 1; }
        }
        printf("restarted from %s\n", restart_dir.c_str());
    }

    double sim_time = 0.0;   // time += dt, main.cpp:293
    long iter = 0;
    auto save = [&](const char* title) {
        for (int b = 0; b < nb; b++) {
            std::vector<double> unk((size_t)4 * (blocks[b].nx - 1) * (blocks[b].ny - 1));
            if (par2d_get_state(ctx, b, unk.data())) { printf("get_state: %s\n", par2d_last_error()); exit(1); }
            if (cfg.accur == 1) {   // print_state_DGP1, main.cpp:215-217,486-487,541-542
                std::vector<double> ux(unk.size()), uy(unk.size());
                if (par2d_get_slopes(ctx, b, ux.data(), uy.data())) { printf("get_slopes: %s\n", par2d_last_error()); exit(1); }
                char name[64];
                snprintf(name, sizeof name, "/solution.b%03d.it%05ld.tec", b + 1, iter);
                if (par2d_write_solution_dgp1((out_dir + name).c_str(), title, &cfg, blocks[b].nx, blocks[b].ny, blocks[b].x, blocks[b].y,
                                              unk.data(), ux.data(), uy.data(), (int)iter, sim_time))
                    printf("%s\n", par2d_host_last_error());
                continue;
            }
            const std::string path = out_dir + "/solution.b" + std::to_string(b + 1) + ".tec";  // bnum+1, FileIO.cpp:84
            if (par2d_write_solution(path.c_str(), title, &cfg, blocks[b].nx, blocks[b].ny, blocks[b].x, blocks[b].y, unk.data()))
                printf("%s\n", par2d_host_last_error());
        }
    };
    printf("===== Generating Mesh and Initial State Tecplot Files ====\n");
    save("Initial State");

    FILE* fres = fopen((out_dir + "/res.tec").c_str(), "w");
    if (!fres) printf("~~~~~~~~~ Failed to save residual file\n");
    else fprintf(fres, "Residual history\n");

    printf("==================== Starting Solver ====================\n");
    std::vector<par2d_step_info> info(batch);
    bool converged = false;
    int rc = 0;
    while (iter < niter && !converged) {
        // stop at the next save point so the written state is the reference's (iter % saveiter == 0, iter > 0)
        long n = std::min<long>(batch, niter - iter);
        if (saveiter > 0) {
            const long next_save = (iter / saveiter + 1) * (long)saveiter;  // state after iteration next_save is saved
            n = std::min<long>(n, next_save + 1 - iter);
        }
        // the reference only tests the tolerance in multi-block runs (main.cpp:506)
        rc = par2d_step(ctx, (int)n, nb > 1 ? 1 : 0, info.data());
        for (long s = 0; s < n; s++) {
            if (info[s].iter < 0) { converged = true; break; }
            if (fres) fprintf(fres, "%d,\t%le\n", (int)info[s].iter, info[s].relres);
            if (info[s].iter % printiter == 0)
                printf("Iter:%7d\tdt:%7.4e \t\t RelativeTotalResisual:  %8.5e \n", (int)info[s].iter, info[s].dt, info[s].relres);
            iter = info[s].iter + 1;
            sim_time += info[s].dt;
            if (nb > 1 && info[s].relres < cfg.tol) { converged = true; break; }
        }
        if (rc) { printf("Failed Assert on the device: %s\n", par2d_last_error()); break; }
        if (!converged && saveiter > 0 && iter > 1 && (iter - 1) % saveiter == 0) save("Final State");
    }
    if (fres) fclose(fres);
    printf("==================== Calculation Finished ====================\n");
    save("Final State");
    par2d_destroy(ctx);
    par2d_case_free(cs);
    return rc ? 1 : 0;
}
