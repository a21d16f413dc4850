/* par2d_gpu.h -- C ABI of the B200-native PAR2D explicit finite-volume
 * residual + update path.
 *
 * The reference (MrPezser/PAR2D, FV/) has no plugin/FFI layer: its driver
 * (FV/main.cpp) calls free functions on raw double and int arrays, one MPI rank
 * per grid block.  This header is the seam a maintainer binds instead of those
 * calls.  Every entry point names the reference interface it replaces.
 *
 * Conventions kept from the reference
 *   - state per cell: 4 doubles (rho, u, v, T), array-of-structures,
 *     unk[(j*(nx-1)+i)*4 + k]           (FV/Indexing.h:10-11, FV/main.cpp:163)
 *   - nx, ny are NODE counts of a block; cells are (nx-1) x (ny-1)
 *   - sides are ordered bottom, right, top, left like bids[]  (FV/main.cpp:110)
 *   - ibound[] holds one BC code per boundary face, counter-clockwise from the
 *     bottom-left corner (FV/SpatialDiscretization.cpp:27,45,64,84)
 *   - y coordinates are the shifted ones (file value + 1e-4, FV/FileIO.cpp:41);
 *     par2d_b200's case loader applies the shift, this ABI does not.
 *
 * Plain C types only; no CUDA or torch types appear in any signature.  All
 * functions return 0 on success or a PAR2D_ERR_* code; par2d_last_error()
 * returns a message for the calling thread's last failure.  The reference
 * aborts with exit(0) instead (FV/Indexing.h:23); a host that wants that
 * behaviour can do so on a non-zero return.
 *
 * There is NO CPU fallback: every compute entry point fails with
 * PAR2D_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PAR2D_GPU_H
#define PAR2D_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAR2D_NVAR 4

/* side order of bid[] / bbound[]  (FV/main.cpp:110 "bottom, right, top left") */
enum { PAR2D_BOTTOM = 0, PAR2D_RIGHT = 1, PAR2D_TOP = 2, PAR2D_LEFT = 3 };

/* return codes; the numerical ones mirror the reference's ASSERT sites */
enum {
    PAR2D_OK = 0,
    PAR2D_ERR_ARG = 1,          /* bad argument / inconsistent block table            */
    PAR2D_ERR_CUDA = 2,         /* CUDA runtime/driver failure or no usable device     */
    PAR2D_ERR_NAN = 3,          /* NaN residual/dt (SpatialDiscretization.cpp:1008, main.cpp:35,47) */
    PAR2D_ERR_DENSITY = 4,      /* rho <= 0 after update            (main.cpp:401)    */
    PAR2D_ERR_STATE = 5,        /* bad wave speed / pressure  (StateVariables.h:65-67) */
    PAR2D_ERR_SINGULAR = 6,     /* Jacobian pivot below 1e-16          (LUtools.cpp:40) */
    PAR2D_ERR_THERMO_RANGE = 7, /* T outside (200,10000) in curve-fit mode (Thermo.cpp:22-23,40-41) */
    PAR2D_ERR_COMM = 8          /* multi-rank exchange misuse / timeout                */
};

/* arithmetic mode of the device path */
enum {
    PAR2D_MODE_FAST = 0,     /* FMA contraction, shared reciprocals, closed-form 4x4 solve;
                                agrees with the reference to ~1e-13 relative per step      */
    PAR2D_MODE_FAITHFUL = 1  /* reference operation order, IEEE div/sqrt, no FMA, pivoted LU;
                                bit-identical to the reference wherever libm pow() is not
                                involved (subsonic far-field BCs)                          */
};

/* Run parameters.  The reference fixes these at compile time
 * (FV/Indexing.h:12-20) or as literals (FV/main.cpp:77-94, Thermo.h:25,46,
 * BoundaryConditions.cpp:120); par2d_default_config() fills the shipped values. */
typedef struct par2d_config {
    int32_t iaxi;          /* 1 axisymmetric (shipped), 0 planar      Indexing.h:18     */
    int32_t mode;          /* PAR2D_MODE_*                                               */
    double igam;           /* >0: calorically perfect shortcut; <=0: NASA-7 curve fit;
                              gamma used for the sound speed is fabs(igam)  Indexing.h:16, Thermo.h:47 */
    double cfl;            /* 0.01                                     main.cpp:79       */
    double tol;            /* 1e-6 relative residual                   main.cpp:78       */
    double uFS[PAR2D_NVAR];/* free stream (rho,u,v,T)                  main.cpp:179-183  */
    double p_back;         /* 101325 back pressure of BC 2   BoundaryConditions.cpp:120  */
    double Rs;             /* 8314.34/28.9                             Thermo.h:25,31,46 */
    double thermo_hi[7];   /* NASA-7 a1..a7, T > 1000 K (Ah..Gh)       Thermo.h:27-28    */
    double thermo_lo[7];   /* NASA-7 a1..a7, T <= 1000 K (Al..Gl)      Thermo.h:28-29    */
    int32_t device;        /* CUDA device ordinal; -1 = current device                   */
    int32_t rank;          /* this process's rank  (0 for single process)                */
    int32_t nranks;        /* number of processes sharing the block table (1 = single)   */
    int32_t accur;         /* 0: first-order FV (shipped); 1: the reference's DG(P1) second-order branch
                              (cell mean + xi/eta slopes, 2 Gauss points per face, damped + limited
                              slopes)                                  Indexing.h:17     */
} par2d_config;

/* One grid block = what one MPI rank of the reference holds after read_mesh()
 * (FV/FileIO.cpp:13-51) and the grid.conn parse (FV/main.cpp:117-142). */
typedef struct par2d_block_desc {
    int32_t nx, ny;        /* node counts                                                */
    const double* x;       /* nx*ny node coordinates, j outer, i inner; may be NULL for  */
    const double* y;       /*   blocks owned by another rank                             */
    const int32_t* ibound; /* 2(nx-1)+2(ny-1) BC codes, CCW from bottom-left; NULL if not owned */
    int32_t bid[4];        /* neighbour block id per side, or -1        main.cpp:131-133 */
    int32_t bbound[4];     /* conn-file BC per side (dead in the reference path, kept)   */
    int32_t owner;         /* rank that advances this block                              */
    int32_t reserved;
} par2d_block_desc;

/* Per-step log record = what the reference prints and appends to res.tec
 * (FV/main.cpp:494-505). */
typedef struct par2d_step_info {
    int64_t iter;          /* iteration index of this step (0-based, as printed)         */
    double dt;             /* global time step = max over blocks        main.cpp:267-292 */
    double ressumsum;      /* sum over OWNED blocks (block order) of sum_k ||res_k||_2   */
    double res0sum;        /* same for the iteration-0 reference norms  main.cpp:365-392 */
    double relres;         /* ressumsum/res0sum; the global value when nranks == 1       */
    int32_t status;        /* PAR2D_OK or first PAR2D_ERR_* raised on the device         */
    int32_t reserved;
} par2d_step_info;

typedef struct par2d_ctx par2d_ctx;

/* ---- configuration ------------------------------------------------------ */

/* Shipped run parameters: FV/Indexing.h:12-20, FV/main.cpp:77-94, and the AIR
 * record of FV/therm_air.dat parsed as Thermo.h:16-34 does. */
void par2d_default_config(par2d_config* cfg);

const char* par2d_last_error(void);

/* ---- life cycle --------------------------------------------------------- */

/* Replaces the per-rank setup of FV/main.cpp:98-176 (read_mesh result,
 * calc_geoel_geofa, malloc of unk/res/dv): uploads node coordinates and BC
 * codes of the blocks owned by cfg->rank, builds the neighbour/ghost tables
 * (SpatialDiscretization.cpp:217-471 pairing rules) and allocates device
 * state.  Host arrays stay caller-owned and may be freed after the call.
 * The state is initialised to cfg->uFS in every cell (main.cpp:193-197). */
int par2d_create(const par2d_config* cfg, int32_t nblocks,
                 const par2d_block_desc* blocks, par2d_ctx** out);
int par2d_destroy(par2d_ctx* ctx);

/* ---- state access (host buffers, reference AoS layout) ------------------ */

/* unk[(j*(nx-1)+i)*4+k] of one owned block; replaces direct access to the
 * reference's unk array (main.cpp:163,193-197,394-404). */
int par2d_set_state(par2d_ctx* ctx, int32_t block, const double* unk);
int par2d_get_state(par2d_ctx* ctx, int32_t block, double* unk);
/* The same for n owned blocks in one call: the host<->device copy of one block overlaps the AoS<->plane transpose
 * of the previous one (two staging buffers, a copy stream); give pinned host buffers for the copies to be
 * asynchronous.  blocks[i] = global block id, unk[i] = that block's array. */
int par2d_set_states(par2d_ctx* ctx, int32_t n, const int32_t* blocks, const double* const* unk);
int par2d_get_states(par2d_ctx* ctx, int32_t n, const int32_t* blocks, double* const* unk);

/* Test hook = one calc_dudt() evaluation (SpatialDiscretization.h:10) on the
 * current state, ghost fill and halo exchange included: res[(j*(nx-1)+i)*4+k]
 * = dudt.  Does not advance the state or the iteration counter. */
int par2d_get_residual(par2d_ctx* ctx, int32_t block, double* res);

/* Second-order branch (cfg.accur == 1): the xi / eta slopes of the DG(P1) representation, reference layout
 * ux/uy[(j*(nx-1)+i)*4+k] (FV/main.cpp:169-172; zero after create, main.cpp:199-209), and the three
 * right-hand sides dudt, duxdt, duydt of one calc_dudt evaluation with ACCUR == 1 (test hook,
 * FV/SpatialDiscretization.cpp:159-1050 + FV/DGP1Tools.cpp). */
int par2d_set_slopes(par2d_ctx* ctx, int32_t block, const double* ux, const double* uy);
int par2d_get_slopes(par2d_ctx* ctx, int32_t block, double* ux, double* uy);
int par2d_get_residual_dg(par2d_ctx* ctx, int32_t block, double* res, double* resx, double* resy);

/* ---- the hot path ------------------------------------------------------- */

/* Advances nsteps explicit-Euler iterations = nsteps passes of the loop body
 * FV/main.cpp:263-532 (find_dt + max over blocks, calc_dudt, calculate_residual,
 * per-cell Jacobian solve with the 1/ycc axisymmetric scaling, unk += dv,
 * UpdateState clamps, residual bookkeeping).  info may be NULL or point to
 * nsteps records.  With nranks == 1 the whole batch runs without host
 * synchronisation; the call returns after the batch has completed.  Stops
 * early (remaining records get iter = -1) once relres < cfg->tol when
 * stop_on_tol != 0 and nranks == 1, like main.cpp:506-511 does for multi-block
 * runs; the state is then the one after the converged iteration. */
int par2d_step(par2d_ctx* ctx, int32_t nsteps, int32_t stop_on_tol, par2d_step_info* info);
/* par2d_step in two halves: _async queues the whole batch on the ctx stream and returns, _wait blocks until it has
 * completed and fills info[nsteps of the _async call].  One host thread can so drive the contexts of several GPUs
 * (par2d_peer_connect_local): _async on every ctx, then _wait on every ctx - the batches of peer-mapped contexts
 * depend on each other's progress.  That only works while the whole batch of the first ctx fits the driver's launch
 * queue (a few dozen steps): a thread that blocks in a launch for ctx 0 never queues what ctx 0's kernels wait for.
 * For longer batches drive each ctx from its own host thread with par2d_step (par2d_fv --gpus N and
 * par2d_b200/local.py do). */
int par2d_step_async(par2d_ctx* ctx, int32_t nsteps, int32_t stop_on_tol);
int par2d_step_wait(par2d_ctx* ctx, par2d_step_info* info);

/* Per-owned-block residual sums of the last `nsteps` steps, in owned-block
 * order: out[(s*nowned + b)*2 + {0: sum_k||res_k||, 1: sum_k res0_k}]
 * (main.cpp:379).  Lets a multi-rank host add blocks in global block order so
 * that relres does not depend on the number of GPUs. */
int par2d_get_block_history(par2d_ctx* ctx, int32_t nsteps, double* out);

int64_t par2d_iteration(const par2d_ctx* ctx);
int32_t par2d_num_owned(const par2d_ctx* ctx);
/* global block id of the k-th owned block */
int32_t par2d_owned_block(const par2d_ctx* ctx, int32_t k);
/* cells advanced per step by this rank */
int64_t par2d_owned_cells(const par2d_ctx* ctx);

/* ---- multi-rank halo exchange (replaces the 4 MPI_Sendrecv sites
 *      SpatialDiscretization.cpp:259,320,380,441 and the dt gather/bcast
 *      main.cpp:270-292) ------------------------------------------------- */

/* External-transport mode: the host moves the bytes (NCCL/MPI/...).
 * For every peer rank the library keeps one contiguous send and one contiguous
 * receive buffer in DEVICE memory holding all edge lines shared with that peer,
 * ordered by (receiving block id, receiving side), each line 4*(edge cells)
 * doubles packed exactly like usend[] (SpatialDiscretization.cpp:236-257).
 * counts are in doubles and may be 0. */
int par2d_exchange_layout(par2d_ctx* ctx, int32_t peer, int64_t* n_send, int64_t* n_recv,
                          void** dev_send, void** dev_recv);
/* Device table of per-block time steps dt_b (nblocks doubles, 0 for blocks not
 * owned); the host max-reduces it across ranks (main.cpp:276 uses fmax). */
int par2d_dt_table(par2d_ctx* ctx, void** dev_table, int32_t* n);
/* step = begin (pack edge lines + dt_b, asynchronous on the ctx stream)
 *        -> host transport on the same stream -> end (ghost fill, residual,
 *        update).  par2d_step() is begin+end for nranks == 1 or peer-mapped mode. */
int par2d_step_begin(par2d_ctx* ctx);
int par2d_step_end(par2d_ctx* ctx);
/* Blocks until queued work is done and returns the log records of the last
 * nsteps steps taken with begin/end (local sums only). */
int par2d_step_collect(par2d_ctx* ctx, int32_t nsteps, par2d_step_info* info);

/* Peer-mapped mode (NVLink P2P between the processes of one box): each rank
 * exports a 64-byte handle to its receive arena (cudaIpcMemHandle_t), the host
 * all-gathers the handles with whatever it has (torch.distributed, MPI, a file),
 * and after par2d_peer_connect() par2d_step() becomes a COLLECTIVE over the
 * ranks (same nsteps everywhere, stop_on_tol ignored) that needs no host
 * transport: the put kernel stores edge lines and dt_b straight into the
 * neighbours' arenas over NVLink and raises a per-source sequence flag there;
 * the ghost kernel of the consuming step waits on those flags (bounded:
 * PAR2D_PEER_TIMEOUT_MS, default 20000, then PAR2D_ERR_COMM).  A whole batch
 * of steps is queued without host synchronisation.  info[] carries this rank's
 * local sums (use par2d_get_block_history for the global-order relres).
 * par2d_get_residual() uses the halos left by the last par2d_step(); after
 * par2d_set_state() call par2d_step(ctx, 0, 0, NULL) on every rank first.
 * par2d_peer_disconnect() unmaps the peers' arenas; call it on every rank
 * (and synchronise the ranks) before any rank destroys its ctx. */
#define PAR2D_PEER_HANDLE_BYTES 64
int par2d_peer_export(par2d_ctx* ctx, void* handle_out);
int par2d_peer_connect(par2d_ctx* ctx, const void* all_handles /* nranks*64 bytes */);
int par2d_peer_disconnect(par2d_ctx* ctx);
/* After every rank has returned from par2d_peer_connect (host barrier): verifies that all ranks cut their blocks
 * into the same strips, i.e. add the residual norms in the same order. */
int par2d_peer_check(par2d_ctx* ctx);
/* Diagnostic: this rank's view of the exchange after a batch (e.g. after PAR2D_ERR_COMM): flags_out[nranks*16] =
 * the flag groups of its arena (slot 0: exchange events completed by each rank, slot 1: iterations whose residual
 * sums each rank has published, slot 15: rows per strip), res_out[4*nblocks*2] = the residual table, counters_out[6]
 * = {exchange events produced, launch whose stop decision is made, stop word, stop iteration, iteration counter,
 * history records}.  Any pointer may be NULL. */
int par2d_peer_snapshot(par2d_ctx* ctx, uint64_t* flags_out, double* res_out, int64_t* counters_out);
/* The same connection for n contexts of ONE process (ctxs[r] = rank r of n, one per GPU; replaces the rank setup of
 * FV/main.cpp:68-74,125-142 for a single-process multi-GPU host): enables peer access between the devices and
 * hands every context the others' arenas directly - no handle exchange, no second process. */
int par2d_peer_connect_local(par2d_ctx** ctxs, int32_t n);

/* All device work of this ctx is issued on one CUDA stream.  stream is a
 * cudaStream_t passed as an integer (0 = a private non-blocking stream). */
int par2d_set_stream(par2d_ctx* ctx, uint64_t stream);
int par2d_synchronize(par2d_ctx* ctx);

/* ---- point-wise test hooks: the device functions of the hot path applied to
 *      n independent inputs (host arrays in, host arrays out) -------------- */

/* LDFSS  (EulerFlux.cpp:121-192): face[i] = {normx,normy,len,yface},
 * uL/uR[i] = (rho,u,v,T); out flux[i*4..], parr[i]. */
int par2d_test_ldfss(const par2d_config* cfg, int64_t n, const double* face,
                     const double* uL, const double* uR, double* flux, double* parr);
/* boundary_state + UpdateState of the ghost (BoundaryConditions.cpp:61-190,
 * SpatialDiscretization.cpp:37-39): norm[i] = {normx,normy}. */
int par2d_test_boundary_state(const par2d_config* cfg, int64_t n, const int32_t* btype,
                              const double* norm, const double* uInt, double* uGhost);
/* State::UpdateState (StateVariables.h:38-68): derived[i*8..] = p,a,h,h0,e0,v2,Cp,Cv;
 * unk_inout gets the T clamp. */
int par2d_test_update_state(const par2d_config* cfg, int64_t n, double* unk_inout, double* derived);
/* BuildJacobian + LUPDecompose + LUPSolve + 1/ycc (main.cpp:307-331). */
int par2d_test_solve(const par2d_config* cfg, int64_t n, const double* dt_ycc /* {dt,ycc} per i */,
                     const double* unk, const double* res, double* dv);

/* Measurement hook: when enabled, every launch of the dominant (residual+update) kernel is
 * bracketed by CUDA events on the ctx stream; par2d_kernel_time returns and resets the sum of
 * their durations and the launch count.  bench.py derives roofline.achieved from it. */
int par2d_kernel_timing(par2d_ctx* ctx, int32_t enable);
int par2d_kernel_time(par2d_ctx* ctx, double* total_ms, int64_t* launches);
/* Kernel launches of every kind this ctx has issued since the last par2d_kernel_timing / par2d_launch_count call
 * (bench.py's gpu_launches). */
int par2d_launch_count(par2d_ctx* ctx, int64_t* n);

/* Diagnostic of the persistent launch (ctx created with PAR2D_PROFILE_PERSIST=1 in the environment): for each of the
 * first nsteps steps of the last batch, out[12*s + {0: SM cycles the first strip spent in its step, 1: ... waiting at the
 * grid barrier, 2: cycles of the first bookkeeping CTA, 3..6: inside the strip (7..10: inside its perimeter epilogue), cumulative from the wait for its first
 * staged rows: that wait, + prologue and rows, + norms, + perimeter epilogue}]. */
int par2d_persist_profile(par2d_ctx* ctx, int64_t* out, int32_t nsteps);

/* Number of CUDA devices this process sees (the reference asks MPI_Comm_size, FV/main.cpp:70-71). */
int par2d_device_count(int32_t* n);

/* Measured fp64 FMA throughput of the device in TFLOP/s (2 flops per FMA);
 * used by bench.py to report the secondary (fp64-pipe) bound. */
int par2d_probe_fp64(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PAR2D_GPU_H */
