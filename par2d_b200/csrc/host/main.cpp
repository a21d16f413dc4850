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
//            ACCUR = 1: <out>/time/solution.b%03d.it%05d.tec         (FV/FileIO.cpp:121-128)
// The reference is one MPI rank per block (main.cpp:68-74,125-142); here ONE process drives --gpus N
// GPUs: the blocks are partitioned over N contexts (par2d_case_partition), the contexts are connected
// through peer-mapped arenas (par2d_peer_connect_local) and every batch is advanced by one host thread per
// context (the contexts' kernels wait for each other: a single thread that blocks while queueing for one
// context would never queue what that context waits for).  The residual is gathered in global block order
// (main.cpp:379-392) and the tolerance break of multi-block runs (main.cpp:506-531) is decided on the
// device from the sums of all contexts, so every GPU stops after the same iteration.
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "par2d_gpu.h"
#include "par2d_host.h"

static void usage() {
    printf("usage: par2d_fv --case DIR [--out DIR] [--therm FILE] [--niter N] [--iaxi 0|1] [--igam G]\n"
           "                [--cfl C] [--tol T] [--printiter N] [--saveiter N] [--batch N] [--faithful]\n"
           "                [--restart DIR]   start from DIR/solution.b<b+1>.tec instead of the uniform state\n"
           "                [--accur 0|1]     1: the DG(P1) second-order branch (ACCUR, FV/Indexing.h:17)\n"
           "                [--gpus N]        shard the blocks over N contexts, context r on device r mod (devices present)\n");
}

int main(int argc, char** argv) {
    // the contexts of one process wait for each other's kernels: none may be loaded lazily while another one spins
    // (par2d_create preloads what it launches; this covers the rest)
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    std::string case_dir = "../Case", out_dir = "../Outputs", therm = "../therm_air.dat", restart_dir;
    long niter = 5000000;  // NITER, FV/Indexing.h:20
    int printiter = 1, saveiter = 1000, batch = 100, ngpus = 1, ndev_arg = -1;
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
        else if (a == "--gpus") ngpus = atoi(next());
        else if (a == "--devices") ndev_arg = atoi(next());
        else { usage(); return 2; }
    }
    if (batch < 1) batch = 1;
    if (ngpus < 1) ngpus = 1;
    if (cfg.accur == 1 && !restart_dir.empty()) {
        // the DG(P1) files hold node averages (print_state_DGP1), from which neither the cell means nor the
        // slopes can be recovered
        printf("--restart needs cell-centred solution.b<N>.tec files, which --accur 1 runs do not write\n");
        return 2;
    }
    if (cfg.accur == 1 && ngpus > 1) {
        printf("--gpus N > 1 drives the first-order path; run the second-order branch on several GPUs through par2d_b200/dist.py\n");
        return 2;
    }

    printf("==================== Loading Mesh ====================\n");
    par2d_case* cs = nullptr;
    if (par2d_case_load(case_dir.c_str(), &cs)) { printf("%s\n", par2d_host_last_error()); return 1; }
    if (par2d_thermo_load(therm.c_str(), &cfg)) { printf("%s\n", par2d_host_last_error()); return 1; }
    // free stream from the loaded gas constant, FV/main.cpp:84-88
    cfg.uFS[0] = 1000.0 / (cfg.Rs * 300.0);
    const int nb = par2d_case_num_blocks(cs);
    if (ngpus > nb) { printf("%d blocks cannot be spread over %d GPUs (a context must own a block)\n", nb, ngpus); return 2; }
    if (ngpus > 1 && par2d_case_partition(cs, ngpus)) { printf("%s\n", par2d_host_last_error()); return 1; }
    std::vector<par2d_block_desc> blocks(nb);
    for (int b = 0; b < nb; b++) par2d_case_block(cs, b, &blocks[b]);

    printf("==================== Initializing ====================\n");
    int ndev = ndev_arg;
    if (ndev < 1 && par2d_device_count(&ndev)) { printf("%s\n", par2d_last_error()); return 1; }
    std::vector<par2d_ctx*> ctx(ngpus, nullptr);
    for (int r = 0; r < ngpus; r++) {
        par2d_config c = cfg;
        c.rank = r;
        c.nranks = ngpus;
        c.device = ngpus > 1 ? r % ndev : cfg.device;
        if (par2d_create(&c, nb, blocks.data(), &ctx[r])) { printf("par2d_create (rank %d): %s\n", r, par2d_last_error()); return 1; }
    }
    if (ngpus > 1) {
        if (par2d_peer_connect_local(ctx.data(), ngpus)) { printf("par2d_peer_connect_local: %s\n", par2d_last_error()); return 1; }
        printf("%d blocks on %d contexts (%d device%s), halo exchange through peer-mapped arenas\n", nb, ngpus, std::min(ndev, ngpus), std::min(ndev, ngpus) > 1 ? "s" : "");
    }
    if (ngpus > ndev && batch > 16) {
        // contexts that share a GPU must not run far ahead of each other (their kernels wait for each other inside the
        // same work queues): short batches
        batch = 16;
    }
    auto owner_ctx = [&](int b) { return ctx[ngpus > 1 ? blocks[b].owner : 0]; };

    if (!restart_dir.empty()) {
        // restart (the reference writes solution.b<N>.tec every saveiter iterations but cannot read them back)
        for (int b = 0; b < nb; b++) {
            std::vector<double> unk((size_t)4 * (blocks[b].nx - 1) * (blocks[b].ny - 1));
            const std::string path = restart_dir + "/solution.b" + std::to_string(b + 1) + ".tec";
            if (par2d_read_solution(path.c_str(), blocks[b].nx, blocks[b].ny, unk.data(), nullptr, nullptr)) { printf("%s\n", par2d_host_last_error()); return 1; }
            if (par2d_set_state(owner_ctx(b), b, unk.data())) { printf("set_state: %s\n", par2d_last_error()); return 1; }
        }
        printf("restarted from %s\n", restart_dir.c_str());
    }

    double sim_time = 0.0;   // time += dt, main.cpp:293
    long iter = 0;           // next iteration index
    if (cfg.accur == 1) mkdir((out_dir + "/time").c_str(), 0777);   // FileIO.cpp:121-125
    // label: the iteration index print_state_DGP1 puts into the file name (0, the saved iteration, the loop's exit value)
    auto save = [&](const char* title, long label) {
        for (int b = 0; b < nb; b++) {
            std::vector<double> unk((size_t)4 * (blocks[b].nx - 1) * (blocks[b].ny - 1));
            if (par2d_get_state(owner_ctx(b), b, unk.data())) { printf("get_state: %s\n", par2d_last_error()); exit(1); }
            if (cfg.accur == 1) {   // print_state_DGP1, main.cpp:215-217,486-487,541-542
                std::vector<double> ux(unk.size()), uy(unk.size());
                if (par2d_get_slopes(owner_ctx(b), b, ux.data(), uy.data())) { printf("get_slopes: %s\n", par2d_last_error()); exit(1); }
                char name[64];
                snprintf(name, sizeof name, "/time/solution.b%03d.it%05ld.tec", b + 1, label);
                if (par2d_write_solution_dgp1((out_dir + name).c_str(), title, &cfg, blocks[b].nx, blocks[b].ny, blocks[b].x, blocks[b].y,
                                              unk.data(), ux.data(), uy.data(), (int)label, sim_time))
                    printf("%s\n", par2d_host_last_error());
                continue;
            }
            const std::string path = out_dir + "/solution.b" + std::to_string(b + 1) + ".tec";  // bnum+1, FileIO.cpp:84
            if (par2d_write_solution(path.c_str(), title, &cfg, blocks[b].nx, blocks[b].ny, blocks[b].x, blocks[b].y, unk.data()))
                printf("%s\n", par2d_host_last_error());
        }
    };
    printf("===== Generating Mesh and Initial State Tecplot Files ====\n");
    save("Initial State", 0);

    FILE* fres = fopen((out_dir + "/res.tec").c_str(), "w");
    if (!fres) printf("~~~~~~~~~ Failed to save residual file\n");
    else fprintf(fres, "Residual history\n");

    printf("==================== Starting Solver ====================\n");
    std::vector<std::vector<par2d_step_info>> info(ngpus, std::vector<par2d_step_info>(batch));
    std::vector<double> hist, num, den;
    bool converged = false;
    long last_iter = -1;     // index of the last iteration executed
    int rc = 0;
    while (iter < niter && !converged) {
        // stop where the reference saves: after iteration m for every m > 0 with m % saveiter == 0 (main.cpp:484)
        long n = std::min<long>(batch, niter - iter);
        if (saveiter > 0) {
            const long m = std::max<long>(saveiter, (iter + saveiter - 1) / saveiter * (long)saveiter);   // next save point >= iter
            n = std::min<long>(n, m + 1 - iter);
        }
        // the reference only tests the tolerance in multi-block runs (main.cpp:506)
        const int stop = nb > 1 ? 1 : 0;
        std::vector<int> rcs(ngpus, 0);
        std::vector<std::string> errs(ngpus);
        auto advance = [&](int r) {
            rcs[r] = par2d_step(ctx[r], (int)n, stop, info[r].data());
            if (rcs[r]) errs[r] = par2d_last_error();      // the error text is per thread
        };
        if (ngpus == 1) {
            advance(0);
        } else {
            std::vector<std::thread> workers;
            for (int r = 0; r < ngpus; r++) workers.emplace_back(advance, r);
            for (std::thread& w : workers) w.join();
        }
        for (int r = 0; r < ngpus; r++)
            if (rcs[r] && !rc) { rc = rcs[r]; printf("Failed Assert on the device (context %d): %s\n", r, errs[r].c_str()); }
        long done = 0;
        while (done < n && info[0][done].iter >= 0) done++;
        if (ngpus > 1 && done > 0) {
            // the gather of main.cpp:379-392: per-block sums of every context, added in global block order
            num.assign(done, 0.0);
            den.assign(done, 0.0);
            std::vector<double> table((size_t)done * nb * 2, 0.0);
            for (int r = 0; r < ngpus; r++) {
                const int no = par2d_num_owned(ctx[r]);
                hist.resize((size_t)done * no * 2);
                if (par2d_get_block_history(ctx[r], (int)done, hist.data())) { printf("get_block_history: %s\n", par2d_last_error()); rc = 1; break; }
                for (long s = 0; s < done; s++)
                    for (int k = 0; k < no; k++) {
                        const int g = par2d_owned_block(ctx[r], k);
                        table[((size_t)s * nb + g) * 2] = hist[((size_t)s * no + k) * 2];
                        table[((size_t)s * nb + g) * 2 + 1] = hist[((size_t)s * no + k) * 2 + 1];
                    }
            }
            for (long s = 0; s < done; s++)
                for (int g = 0; g < nb; g++) { num[s] += table[((size_t)s * nb + g) * 2]; den[s] += table[((size_t)s * nb + g) * 2 + 1]; }
        }
        for (long s = 0; s < done; s++) {
            const par2d_step_info& I = info[0][s];
            const double relres = ngpus > 1 ? num[s] / den[s] : I.relres;
            if (fres) fprintf(fres, "%d,\t%le\n", (int)I.iter, relres);
            if (I.iter % printiter == 0)
                printf("Iter:%7d\tdt:%7.4e \t\t RelativeTotalResisual:  %8.5e \n", (int)I.iter, I.dt, relres);
            last_iter = I.iter;
            iter = I.iter + 1;
            sim_time += I.dt;
            if (nb > 1 && relres < cfg.tol) { converged = true; break; }
        }
        if (done < n) converged = true;
        if (rc) break;
        if (!converged && saveiter > 0 && last_iter > 0 && last_iter % saveiter == 0) save("Final State", last_iter);
    }
    if (fres) fclose(fres);
    printf("==================== Calculation Finished ====================\n");
    save("Final State", converged ? last_iter : iter);   // the loop variable's value on exit, main.cpp:541-545
    if (ngpus > 1) for (int r = 0; r < ngpus; r++) par2d_peer_disconnect(ctx[r]);
    for (int r = 0; r < ngpus; r++) par2d_destroy(ctx[r]);
    par2d_case_free(cs);
    return rc ? 1 : 0;
}
