// capi.cu -- implementation of the C ABI in include/par2d_gpu.h over the sm_100a kernels.
//
// One par2d_ctx = the blocks one rank owns (the reference: one MPI rank per block,
// FV/main.cpp:124-128), resident in HBM as padded SoA planes with the ghost ring in place
// (common.cuh).  A step is three launches on one stream, no host synchronisation:
//   ghost_kernel (+ the previous step's residual bookkeeping) -> march_kernel [-> finalize_kernel at the end of a batch]
// and a batch of steps returns its dt / residual history in one copy at the end.
// There is no CPU fallback anywhere in this file.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/par2d_gpu.h"
#include "common.cuh"
#include "kernels.cuh"
#include "kernels_common.cuh"

using namespace par2d;

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return fail(PAR2D_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct Peer {
    long long n_send = 0, n_recv = 0;
    double *d_send = nullptr, *d_recv = nullptr;
};

int desc_edge(const par2d_block_desc& d, int side) { return (side == PAR2D_BOTTOM || side == PAR2D_TOP) ? d.nx - 1 : d.ny - 1; }

// The side of block nb that faces block `me` (opposite side preferred; the reference pairs
// sides implicitly through message order, SpatialDiscretization.cpp:227-470).
int facing_side(const par2d_block_desc& nb, int me, int my_side) {
    const int opp = (my_side + 2) & 3;
    if (nb.bid[opp] == me) return opp;
    for (int s = 0; s < 4; s++)
        if (nb.bid[s] == me) return s;
    return -1;
}

// Arena layout of rank r, a pure function of the global block table (so every rank can address every
// other rank's arena without asking): see common.cuh PutArgs.
struct ArenaLayout {
    size_t flags_off = 0, dt_off = 0, res_off = 0, lines_off = 0, bytes = 0;
    long long dt_stride = 0;     // doubles
    long long lines_stride = 0;  // doubles = everything rank r receives per event
};
// lw = doubles per edge cell of a line: 4 (first order: the cell mean) or 8 (second-order branch: two Gauss points)
long long recv_count(const par2d_block_desc* t, int nblocks, int r, int src, int lw) {  // doubles rank r receives from rank src
    long long n = 0;
    for (int g = 0; g < nblocks; g++) {
        if (t[g].owner != r) continue;
        for (int s = 0; s < 4; s++) {
            const int nb = t[g].bid[s];
            if (nb >= 0 && t[nb].owner == src) n += (long long)lw * desc_edge(t[g], s);
        }
    }
    return n;
}
long long recv_segment_off(const par2d_block_desc* t, int nblocks, int r, int src, int lw) {  // where src's segment starts in r's lines
    long long off = 0;
    for (int p = 0; p < src; p++)
        if (p != r) off += recv_count(t, nblocks, r, p, lw);
    return off;
}
ArenaLayout arena_layout(const par2d_block_desc* t, int nblocks, int nranks, int r, int lw) {
    ArenaLayout L;
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    L.flags_off = 0;
    L.dt_off = up(sizeof(unsigned long long) * 16 * (size_t)nranks);
    L.dt_stride = (long long)((nblocks + 31) / 32 * 32);
    L.res_off = up(L.dt_off + sizeof(double) * 2 * (size_t)L.dt_stride);     // [4 slots][nblocks][2] per-block residual sums (stop test)
    L.lines_off = up(L.res_off + sizeof(double) * 8 * (size_t)nblocks);
    for (int p = 0; p < nranks; p++)
        if (p != r) L.lines_stride += recv_count(t, nblocks, r, p, lw);
    L.bytes = up(L.lines_off + sizeof(double) * 2 * (size_t)std::max<long long>(L.lines_stride, 1));
    return L;
}

Gas make_gas(const par2d_config& c) {
    Gas g;
    memset(&g, 0, sizeof g);
    g.Rs = c.Rs;
    g.igam = c.igam;
    g.gam = fabs(c.igam);
    g.gamRs = g.gam * g.Rs;
    g.gamRs4 = 0.25 * g.gamRs;
    g.cpg_den = (1.0 - 1.0 / (c.igam));
    g.cp_cpg = g.Rs / (1.0 - 1.0 / (c.igam));
    g.inv_cv_cpg = 1.0 / (g.cp_cpg - g.Rs);
    for (int k = 0; k < 7; k++) { g.hi[k] = c.thermo_hi[k]; g.lo[k] = c.thermo_lo[k]; }
    for (int k = 0; k < 4; k++) g.uFS[k] = c.uFS[k];
    g.p_back = c.p_back;
    g.cfl = c.cfl;
    return g;
}

int flags_to_code(unsigned f) {
    if (f & 0x80000000u) return PAR2D_ERR_CUDA;  // TMA copy never completed
    if (f & FLAG_COMM) return PAR2D_ERR_COMM;    // a peer's exchange flag never arrived
    if (f & FLAG_NAN) return PAR2D_ERR_NAN;
    if (f & FLAG_DENSITY) return PAR2D_ERR_DENSITY;
    if (f & FLAG_STATE) return PAR2D_ERR_STATE;
    if (f & FLAG_SINGULAR) return PAR2D_ERR_SINGULAR;
    if (f & FLAG_THERMO) return PAR2D_ERR_THERMO_RANGE;
    return PAR2D_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

struct par2d_ctx {
    par2d_config cfg;
    Gas gas;
    bool tpg = false, axi = true;
    int march_rows = 64;     // cell rows per warp strip of the march kernel (PAR2D_MARCH_ROWS overrides, measurements only)
    const KernelTable* kt = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    int nblocks = 0;
    std::vector<par2d_block_desc> table;  // global, pointers cleared
    std::vector<int> owned_gid, owned_index;
    int nowned = 0;
    std::vector<BlockDev> hblk;
    std::vector<double*> block_mem;       // one allocation per owned block: 11 planes [state A x4][x y metric][state B x4]
    std::vector<int*> block_ibound;
    std::vector<double2*> block_fnorm;
    BlockDev* dblk = nullptr;
    void* dtmaps = nullptr;

    Gas* gas_dev = nullptr;          // the Gas constants in global memory (non-inlined boundary-condition code)
    double* dt_table = nullptr;      // single rank: [2][nblocks], the parity of a step is iter_queued & 1
    double* dt_used = nullptr;       // [2]: the dt of the step in flight / of the step whose bookkeeping is pending
    bool fin_pending = false;        // the last queued step's bookkeeping has not been launched yet
    FinalizeArgs fin_args;
    double* partials = nullptr;
    unsigned* flags = nullptr;
    int* stop = nullptr;
    long long* d_iter = nullptr;
    int* hist_pos = nullptr;
    unsigned* ticket = nullptr;
    double* res0 = nullptr;
    double* blk_norm = nullptr;
    unsigned* flag_hist = nullptr;
    std::vector<unsigned> h_flag_hist;
    double* blk_hist = nullptr;
    double* step_hist = nullptr;
    int hist_cap = 0;
    double* staging = nullptr;
    double* staging2 = nullptr;      // second staging buffer + copy stream of the batched state transfers (created on first use)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    long long staging_doubles = 0;
    // second-order branch (cfg.accur == 1)
    bool dg = false;
    std::vector<double*> dg_mem;      // per owned block: ux | uy | uxt | uyt | ghost points
    double* dg_partials = nullptr;
    int dg_total_tiles = 0, dg_max_tiles = 0, dg_max_lim_tiles = 0;

    int cur = 0;
    long long iter = 0;
    int total_tiles = 0, max_tiles = 0, max_faces = 0;
    long long owned_cells = 0;

    std::vector<Peer> peers;
    SendLine* dlines = nullptr;      // external transport: destinations in the local send buffers
    SendLine* dlines_peer = nullptr; // peer-mapped: destinations in the neighbours' arenas
    std::vector<SendLine> hlines;
    std::vector<int> hline_peer;     // destination rank of each line
    std::vector<long long> hline_off;// offset of each line inside that rank's segment for this source
    int nlines = 0, max_line = 0;

    // arena of this rank (nranks > 1): flags | dt[2] | lines[2]   (common.cuh, PutArgs)
    char* arena = nullptr;
    ArenaLayout lay;
    bool peer_connected = false;
    bool peer_ipc = false;           // the peers' arenas are cudaIpc mappings (separate processes), not same-process pointers
    std::vector<char*> peer_base;    // mapped arena of every rank (own: arena)
    double** d_dt_dst = nullptr;
    unsigned long long** d_flag_dst = nullptr;
    unsigned* put_ticket = nullptr;
    unsigned long long seq = 0;      // exchange events produced so far (identical on every rank)
    double** d_res_dst = nullptr;    // stop test across ranks: residual table of every arena / this rank's residual flag there
    unsigned long long** d_rflag_dst = nullptr;
    int* go = nullptr;               // [0]: launch whose stop decision is made; stop_iter follows
    long long* stop_iter = nullptr;
    // fused step (march.cuh, PUSH): one launch per step.  Across ranks the strips run in tile-list order, those that
    // border another rank (their puts travel while the interior strips run) and strip 0 of every block (its dt) first
    int2* d_push_tiles = nullptr;
    int n_push_tiles = 0;
    unsigned n_put_tiles = 0;        // strips of a launch that put something into a peer's arena
    bool fused = true;               // PAR2D_NO_OVERLAP=1: put -> ghost -> march as separate launches (A/B runs)
    bool small_mesh = false;         // the GLOBAL mesh is small enough that a step is bound by latency, not by HBM: fused / persistent steps
    long long global_cells = 0;
    // large meshes across ranks: the strips that put something run first on a high-priority side stream, followed there by
    // the put of the NEW state's lines; the interior strips run beside them on the main stream (tuned launch-per-step kernel)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_ghost = nullptr, ev_edge = nullptr;
    // persistent launch: a whole batch of a small mesh in one cooperative kernel
    int2* d_all_tiles = nullptr;
    GhostOp* d_ops = nullptr;        // fused step: ghost-cell records, grouped by strip (common.cuh)
    int2* d_op_range = nullptr;
    unsigned* grid_bar = nullptr;
    // ... without a stop test: every step's strip sums / dt / error bits of a launch, and the per-block norms made of them
    double *ring_partials = nullptr, *ring_dt = nullptr, *ring_norm = nullptr, *ring_r0 = nullptr;
    unsigned* ring_flags = nullptr;
    bool defer = true;               // PAR2D_NO_DEFER=1: bookkeeping CTAs inside the persistent launch (A/B runs)
    long long* prof = nullptr;       // PAR2D_PROFILE_PERSIST=1: per-step cycle counts of the persistent launch (par2d_persist_profile)
    int coop_capacity = 0;
    bool persist = true;             // PAR2D_NO_PERSIST=1: one launch per step (A/B runs)
    long long partials_stride = 0;   // doubles between the two parities of the strip sums
    // asynchronous batch (par2d_step_async / par2d_step_wait)
    bool batch_open = false;
    int batch_steps = 0, batch_cur0 = 0, batch_stop = 0;
    unsigned long long batch_seq0 = 0;
    long long timeout_ns = 20000000000ll;

    // history of the last batch, host copies
    std::vector<double> h_step_hist, h_blk_hist;
    int last_steps = 0;
    long long last_iter0 = 0;
    int pending_steps = 0;  // begin/end mode: steps queued since the last collect
    long long iter_queued = 0;  // steps queued so far (parity of the dt_used slot)

    // measurement hook (par2d_kernel_timing)
    bool timing = false;
    std::vector<cudaEvent_t> ev;   // pairs
    int ev_used = 0;
    double kernel_ms = 0.0;
    long long kernel_launches = 0;
    long long launches = 0;        // every kernel launch of this ctx since the last par2d_kernel_timing call

    StepArgs step_args(int update) const {
        StepArgs a;
        a.blocks = dblk;
        a.nowned = nowned;
        a.nblocks_global = nblocks;
        a.cur = cur;
        a.total_tiles = total_tiles;
        a.max_tiles = max_tiles;
        a.max_faces = max_faces;
        a.tmaps = dtmaps;
        a.dt_table = dt_table + (arena ? 0 : (iter_queued & 1) * (long long)nblocks);
        a.dt_next = arena ? nullptr : dt_table + ((iter_queued + 1) & 1) * (long long)nblocks;
        a.nfin = 0;
        a.launch_no = 0;
        a.coop_steps = 1;
        a.grid_bar = grid_bar;
        a.prof = prof;
        a.ops = d_ops;
        a.op_range = d_op_range;
        a.gas_dev = gas_dev;
        a.dt_used = dt_used + (iter_queued & 1);
        a.partials = partials + (iter_queued & 1) * partials_stride;
        a.flags = flags;
        a.stop = stop;
        a.res_out = nullptr;
        a.res_block = -1;
        a.update = update;
        a.march_rows = march_rows;
        a.tile_list = nullptr;
        a.tile_count = 0;
        a.dt_ready = 0;
        a.wait_flags = nullptr;
        a.wait_seq = 0;
        a.recv_off = 0;
        a.nranks = cfg.nranks;
        a.rank = cfg.rank;
        a.timeout_ns = timeout_ns;
        if (peer_connected) {
            // consume the latest exchange event (parity seq & 1) from this rank's arena
            const int par = (int)(seq & 1);
            a.wait_flags = reinterpret_cast<const unsigned long long*>(arena + lay.flags_off);
            a.wait_seq = seq;
            a.recv_off = par * lay.lines_stride;
            a.dt_table = reinterpret_cast<double*>(arena + lay.dt_off) + par * lay.dt_stride;
        }
        return a;
    }
    FinalizeArgs finalize_args(int stop_on_tol) const {
        FinalizeArgs f;
        f.blocks = dblk;
        f.nowned = nowned;
        f.partials = partials + (iter_queued & 1) * partials_stride;
        f.res0 = res0;
        f.blk_hist = blk_hist;
        f.step_hist = step_hist;
        f.iter = d_iter;
        f.hist_pos = hist_pos;
        f.stop = stop;
        f.ticket = ticket;
        f.dt_used = dt_used + (iter_queued & 1);
        f.blk_norm = blk_norm;
        f.flag_hist = flag_hist;
        f.flags = flags;
        memset(&f.stop_x, 0, sizeof f.stop_x);
        f.dg_partials = dg ? dg_partials : nullptr;
        f.tol = cfg.tol;
        f.stop_on_tol = stop_on_tol;
        f.hist_cap = hist_cap;
        return f;
    }
};

namespace {

int use_device(const par2d_ctx* c) {
    CU_TRY(cudaSetDevice(c->device));
    return 0;
}

int ensure_hist(par2d_ctx* c, int nsteps) {
    if (nsteps <= c->hist_cap) return 0;
    const int cap = std::max(nsteps, 64);
    if (c->blk_hist) cudaFree(c->blk_hist);
    if (c->step_hist) cudaFree(c->step_hist);
    if (c->flag_hist) cudaFree(c->flag_hist);
    c->blk_hist = c->step_hist = nullptr;
    c->flag_hist = nullptr;
    CU_TRY(cudaMalloc(&c->flag_hist, sizeof(unsigned) * (size_t)cap));
    CU_TRY(cudaMalloc(&c->blk_hist, sizeof(double) * 2 * (size_t)cap * c->nowned));
    CU_TRY(cudaMalloc(&c->step_hist, sizeof(double) * 4 * (size_t)cap));
    c->hist_cap = cap;
    return 0;
}

int validate_table(int nblocks, const par2d_block_desc* t, int nranks) {
    for (int g = 0; g < nblocks; g++) {
        if (t[g].nx < 2 || t[g].ny < 2) return fail(PAR2D_ERR_ARG, "block %d has fewer than 2 nodes in a direction", g);
        if (t[g].owner < 0 || t[g].owner >= nranks) return fail(PAR2D_ERR_ARG, "block %d: owner %d out of range", g, t[g].owner);
        for (int s = 0; s < 4; s++) {
            const int nb = t[g].bid[s];
            if (nb < 0) continue;
            if (nb >= nblocks || nb == g) return fail(PAR2D_ERR_ARG, "block %d side %d: neighbour id %d out of range", g, s, nb);
            const int os = facing_side(t[nb], g, s);
            if (os < 0) return fail(PAR2D_ERR_ARG, "block %d side %d: neighbour %d does not point back", g, s, nb);
            if (desc_edge(t[g], s) != desc_edge(t[nb], os)) return fail(PAR2D_ERR_ARG, "block %d side %d: shared edge lengths differ", g, s);
        }
    }
    return 0;
}

int build_tensor_maps(par2d_ctx* c) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(PAR2D_ERR_CUDA, "cuTensorMapEncodeTiled is not exported by this driver; the row staging of the march kernel needs it");
    std::vector<CUtensorMap> maps((size_t)c->nowned);
    for (int k = 0; k < c->nowned; k++) {
        const BlockDev& b = c->hblk[k];
        // one map over the block's 11 planes [state A x4][x y metric][state B x4]; one box = the 32*MARCH_CPL
        // staged columns x 2 rows x the 7 consecutive planes a step reads (march.cuh)
        void* base = (void*)c->block_mem[k];
        const cuuint64_t dims[3] = {(cuuint64_t)b.pitch, (cuuint64_t)(b.ncy + 2), 11};
        const cuuint64_t strides[2] = {(cuuint64_t)b.pitch * 8, (cuuint64_t)b.plane * 8};
        const cuuint32_t box[3] = {(cuuint32_t)(32 * MARCH_CPL), 2u, 7u};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&maps[(size_t)k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(PAR2D_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for block %d", (int)r, b.gid);
    }
    CU_TRY(cudaMalloc(&c->dtmaps, maps.size() * sizeof(CUtensorMap)));
    CU_TRY(cudaMemcpy(c->dtmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    return 0;
}

void launch_step(par2d_ctx* c, const StepArgs& a) { c->launches++; c->kt->march(c->stream, a, c->gas, c->tpg, c->axi); }

PutArgs put_args(const par2d_ctx* c);

// peer-mapped mode: produce the next exchange event from the current state with a launch of its own (start of a
// batch; every step of the second-order branch and of the serial A/B order)
void queue_put(par2d_ctx* c, cudaStream_t stream = nullptr) {
    c->seq++;
    c->launches++;
    c->kt->pack(stream ? stream : c->stream, c->step_args(1), c->gas, c->dlines_peer, c->nlines, c->max_line, put_args(c));
}

// The bookkeeping of a step (finalize) is deferred into the NEXT step's ghost launch, where it overlaps
// the ghost fill; flush_finalize() launches it on its own at the end of a batch.
void flush_finalize(par2d_ctx* c) {
    if (!c->fin_pending) return;
    c->launches++;
    launch_finalize(c->stream, c->fin_args);
    c->fin_pending = false;
}

DgArgs dg_args(const par2d_ctx* c, int update) {
    DgArgs d;
    d.partials = c->dg_partials;
    d.max_tiles = c->dg_max_tiles;
    d.max_lim_tiles = c->dg_max_lim_tiles;
    d.max_points = 2 * c->max_faces;
    d.res_out = nullptr;
    d.res_block = -1;
    d.update = update;
    return d;
}

// second-order branch: ghost points -> cell kernel (means into the other buffer, damped slopes) -> limiter ->
// bookkeeping (dg.cuh)
void queue_step_dg(par2d_ctx* c, int stop_on_tol, bool with_ghost_dt) {
    StepArgs a = c->step_args(1);
    const DgArgs d = dg_args(c, 1);
    c->launches++;
    c->kt->dg_ghost(c->stream, a, c->gas, d, with_ghost_dt ? 1 : 0);
    c->launches++;
    c->kt->dg_cell(c->stream, a, c->gas, d, c->tpg, c->axi);
    c->launches++;
    c->kt->dg_limit(c->stream, a, d);
    c->launches++;
    launch_finalize(c->stream, c->finalize_args(stop_on_tol));
    c->iter_queued++;
    c->cur ^= 1;
}

void time_begin(par2d_ctx* c) {
    if (!c->timing) return;
    if ((int)c->ev.size() < 2 * (c->ev_used + 1)) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        c->ev.push_back(e0);
        c->ev.push_back(e1);
    }
    cudaEventRecord(c->ev[2 * c->ev_used], c->stream);
}
void time_end(par2d_ctx* c) {
    if (c->timing) cudaEventRecord(c->ev[2 * c->ev_used++ + 1], c->stream);
}

// stop test across ranks (StopArgs, common.cuh) for the batch that starts at iteration c->iter
StopArgs stop_args(const par2d_ctx* c) {
    StopArgs sx;
    memset(&sx, 0, sizeof sx);
    sx.enabled = 1;
    sx.nranks = c->cfg.nranks;
    sx.rank = c->cfg.rank;
    sx.res_dst = c->d_res_dst;
    sx.rflag_dst = c->d_rflag_dst;
    sx.res_table = reinterpret_cast<const double*>(c->arena + c->lay.res_off);
    sx.local_flags = reinterpret_cast<const unsigned long long*>(c->arena + c->lay.flags_off);
    sx.nblocks_global = c->nblocks;
    sx.go = c->go;
    sx.stop_iter = c->stop_iter;
    sx.iter0 = c->iter;
    sx.tol = c->cfg.tol;
    sx.timeout_ns = c->timeout_ns;
    return sx;
}

PutArgs put_args(const par2d_ctx* c) {
    PutArgs put;
    memset(&put, 0, sizeof put);
    if (!c->peer_connected) return put;
    put.peer_mode = 1;
    put.parity = (int)(c->seq & 1);
    put.seq = c->seq;
    put.dt_dst = c->d_dt_dst;
    put.dt_stride = c->lay.dt_stride;
    put.flag_dst = c->d_flag_dst;
    put.ticket = c->put_ticket;
    put.local_flags = reinterpret_cast<const unsigned long long*>(c->arena + c->lay.flags_off);
    put.nranks = c->cfg.nranks;
    put.rank = c->cfg.rank;
    put.timeout_ns = c->timeout_ns;
    put.expected = c->n_put_tiles;
    put.dg = c->dg ? 1 : 0;
    return put;
}

// Large mesh across ranks: the exchange event of the state being read was produced before this call.
//   main stream : ghost (waits for the peers' flags; previous step's bookkeeping) ... march(interior strips) ... wait(edge) ->
//   side stream :           wait(ghost) -> march(strips that border another rank) -> put(new state's edge lines + dt, flag)
// so the NVLink stores, the flag latency and the skew between GPUs hide behind the interior strips, and every strip runs
// the tuned launch-per-step kernel.
void queue_step_overlapped(par2d_ctx* c) {
    StepArgs a = c->step_args(1);
    a.dt_ready = 1;
    c->launches++;
    c->kt->ghost(c->stream, a, c->gas, 0, c->fin_pending ? &c->fin_args : nullptr);
    c->fin_pending = false;
    cudaEventRecord(c->ev_ghost, c->stream);
    cudaStreamWaitEvent(c->side_stream, c->ev_ghost, 0);
    time_begin(c);
    StepArgs ae = a, ai = a;
    ae.tile_list = c->d_push_tiles; ae.tile_count = (int)c->n_put_tiles;
    ai.tile_list = c->d_push_tiles + c->n_put_tiles; ai.tile_count = c->n_push_tiles - (int)c->n_put_tiles;
    c->launches++;
    c->kt->march(c->side_stream, ae, c->gas, c->tpg, c->axi);
    c->launches++;
    c->kt->march(c->stream, ai, c->gas, c->tpg, c->axi);
    c->fin_args = c->finalize_args(0);
    c->fin_pending = true;
    c->iter_queued++;
    c->cur ^= 1;
    queue_put(c, c->side_stream);      // reads the NEW state's edge cells (all in those strips) and cell (0,0) of every block
    cudaEventRecord(c->ev_edge, c->side_stream);
    cudaStreamWaitEvent(c->stream, c->ev_edge, 0);
    time_end(c);
}

// The fused step (first-order path): ONE launch = the bookkeeping of the previous step (leading CTAs) + every strip;
// the strips on a block's perimeter push the new state's ghost ring, the edge lines of off-rank faces (straight into
// the neighbours' arenas over NVLink, flag raised by the last of them) and the block's dt.  Across ranks the strips
// that border another rank run first, so their puts, the flag latency and the skew between GPUs hide behind the
// interior strips; only those strips wait for a peer's lines, the others only for its dt.
void queue_step_fused(par2d_ctx* c, int launch_no, int stop_on_tol) {
    StepArgs a = c->step_args(1);       // across ranks: consumes exchange event c->seq
    a.dt_ready = 0;
    a.launch_no = launch_no;
    a.nfin = c->fin_pending ? c->nowned : 0;
    FinalizeArgs fin;
    memset(&fin, 0, sizeof fin);
    if (c->fin_pending) fin = c->fin_args;
    const bool across = c->peer_connected && stop_on_tol;
    if (across && !c->fin_pending) fin.stop_x = stop_args(c);   // the stop test of a launch that carries no bookkeeping
    if (c->peer_connected) {
        a.tile_list = c->d_push_tiles;
        a.tile_count = c->n_push_tiles;
        c->seq++;                       // ... and produces the next one
    } else {
        a.tile_list = nullptr;
        a.tile_count = 0;
    }
    const PutArgs put = put_args(c);
    time_begin(c);
    c->launches++;
    c->kt->march_push(c->stream, a, c->gas, put, fin, c->tpg, c->axi);
    time_end(c);
    c->fin_args = c->finalize_args(stop_on_tol);
    if (across) c->fin_args.stop_x = stop_args(c);
    c->fin_pending = true;
    c->iter_queued++;
    c->cur ^= 1;
}

// Single rank, small mesh: the whole batch in one cooperative launch (march_persist_kernel).  Returns false when the
// launch was refused (the caller then falls back to one launch per step - still the CUDA path).
// Without a stop test the bookkeeping leaves the step loop (DeferArgs): launches of at most DEFER_CHUNK steps, each followed by
// the three kernels that turn the kept strip sums into history records.  `done` = steps of this batch queued before.
bool queue_batch_persistent(par2d_ctx* c, int nsteps, int stop_on_tol, int done = 0) {
    const bool defer = c->defer && !stop_on_tol && nsteps >= 16;
    if (defer && nsteps > DEFER_CHUNK) {
        for (int s0 = 0; s0 < nsteps; s0 += DEFER_CHUNK) {
            // (a refused launch after the first cannot happen for a reason the first did not have; the remaining steps then
            //  run as fused launches, which continue the batch's history behind the records written so far)
            if (!queue_batch_persistent(c, std::min(DEFER_CHUNK, nsteps - s0), 0, done + s0)) {
                if (s0 == 0) return false;
                for (int s = s0; s < nsteps; s++) queue_step_fused(c, s + 1, 0);
                return true;
            }
        }
        return true;
    }
    DeferArgs df;
    memset(&df, 0, sizeof df);
    if (defer) {
        if (!c->ring_partials) {
            const size_t n = (size_t)DEFER_CHUNK;
            if (cudaMalloc(&c->ring_partials, sizeof(double) * n * (size_t)c->partials_stride) != cudaSuccess ||
                cudaMalloc(&c->ring_dt, sizeof(double) * n) != cudaSuccess || cudaMalloc(&c->ring_flags, sizeof(unsigned) * n) != cudaSuccess ||
                cudaMalloc(&c->ring_norm, sizeof(double) * n * 4 * c->nowned) != cudaSuccess || cudaMalloc(&c->ring_r0, sizeof(double) * n * 4 * c->nowned) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(c->ring_partials); cudaFree(c->ring_dt); cudaFree(c->ring_flags); cudaFree(c->ring_norm); cudaFree(c->ring_r0);
                c->ring_partials = c->ring_dt = c->ring_norm = c->ring_r0 = nullptr; c->ring_flags = nullptr;
                c->defer = false;                  // no room for the rings: bookkeeping CTAs inside the launch, as with a stop test
                return queue_batch_persistent(c, nsteps, stop_on_tol, done);
            }
        }
        df.on = 1; df.stride = c->partials_stride; df.partials = c->ring_partials; df.dt = c->ring_dt; df.flags = c->ring_flags;
    }
    StepArgs a = c->step_args(1);
    a.dt_ready = 0;
    a.coop_steps = nsteps;
    a.tile_list = c->d_all_tiles;
    a.tile_count = c->total_tiles;
    const FinalizeArgs fin = c->finalize_args(stop_on_tol);
    double* dt_used_alt = c->dt_used + ((c->iter_queued + 1) & 1);
    double* partials_alt = c->partials + ((c->iter_queued + 1) & 1) * c->partials_stride;
    if ((long long)(nsteps + 1) * (c->total_tiles + c->nowned) >= (1ll << 31)) return false;   // the barrier counts arrivals in 31 bits
    if (cudaMemsetAsync(c->grid_bar, 0, sizeof(unsigned) * BAR_WORDS, c->stream) != cudaSuccess) return false;
    time_begin(c);
    c->launches++;
    const cudaError_t e = c->kt->march_coop(c->stream, a, c->gas, fin, dt_used_alt, partials_alt, df, c->tpg, c->axi);
    time_end(c);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    if (defer) {
        c->launches += 3;
        launch_finalize_deferred(c->stream, fin, df, nsteps, c->iter + done, done, c->ring_norm, c->ring_r0);
    }
    c->iter_queued += nsteps;
    c->cur ^= (nsteps & 1);
    return true;
}

void queue_step(par2d_ctx* c, int stop_on_tol, bool with_ghost_dt) {
    if (c->dg) { queue_step_dg(c, stop_on_tol, with_ghost_dt); return; }
    StepArgs a = c->step_args(1);
    a.dt_ready = with_ghost_dt ? 0 : 1;
    c->launches++;
    c->kt->ghost(c->stream, a, c->gas, with_ghost_dt ? 1 : 0, c->fin_pending ? &c->fin_args : nullptr);
    c->fin_pending = false;
    time_begin(c);
    launch_step(c, a);
    time_end(c);
    c->fin_args = c->finalize_args(stop_on_tol);
    c->fin_pending = true;
    c->iter_queued++;
    c->cur ^= 1;
}

// after a synchronise: pull the history of `queued` steps, fix up cur/iter if the batch stopped early
int harvest(par2d_ctx* c, int queued, int cur0, par2d_step_info* info, int ninfo, bool stop_across = false, unsigned long long seq0 = 0) {
    int done = 0;
    unsigned fl = 0;
    long long it = 0;
    CU_TRY(cudaMemcpy(&done, c->hist_pos, sizeof done, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(&fl, c->flags, sizeof fl, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(&it, c->d_iter, sizeof it, cudaMemcpyDeviceToHost));
    if (done > queued) done = queued;
    if (stop_across) {
        // the stop decision was made on the device from the sums of all ranks (stop_test): every rank stops after
        // the same iteration.  One more step may have run and been recorded: drop it
        int stopped = 0;
        long long sit = -1;
        CU_TRY(cudaMemcpy(&stopped, c->stop, sizeof stopped, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(&sit, c->stop_iter, sizeof sit, cudaMemcpyDeviceToHost));
        if (stopped && sit >= c->iter) {
            done = std::min<long long>(done, sit + 1 - c->iter);
            it = c->iter + done;
            CU_TRY(cudaMemcpy(c->d_iter, &it, sizeof it, cudaMemcpyHostToDevice));
            // exchange events really produced: the batch-start put + one per launch that ran
            c->seq = seq0 + 1 + (unsigned long long)std::min(queued, done + 1);
        }
    }
    for (int q = 0; q < c->ev_used; q++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev[2 * q], c->ev[2 * q + 1]) == cudaSuccess) { c->kernel_ms += ms; c->kernel_launches++; }
    }
    c->ev_used = 0;
    c->cur = cur0 ^ (done & 1);
    c->last_iter0 = c->iter;
    c->iter = it;
    c->last_steps = done;
    c->h_step_hist.resize((size_t)4 * std::max(done, 1));
    c->h_blk_hist.resize((size_t)2 * std::max(done, 1) * c->nowned);
    c->h_flag_hist.assign((size_t)std::max(done, 1), 0u);
    if (done > 0) {
        CU_TRY(cudaMemcpy(c->h_step_hist.data(), c->step_hist, sizeof(double) * 4 * done, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(c->h_blk_hist.data(), c->blk_hist, sizeof(double) * 2 * (size_t)done * c->nowned, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(c->h_flag_hist.data(), c->flag_hist, sizeof(unsigned) * done, cudaMemcpyDeviceToHost));
    }
    const int code = flags_to_code(fl);
    if (info) {
        for (int s = 0; s < ninfo; s++) {
            par2d_step_info& r = info[s];
            memset(&r, 0, sizeof r);
            if (s < done) {
                const double* h = &c->h_step_hist[4 * (size_t)s];
                r.iter = c->last_iter0 + s;
                r.dt = h[0]; r.ressumsum = h[1]; r.res0sum = h[2]; r.relres = h[3];
                // error bits as of this step's bookkeeping: the first record with a non-zero status is the
                // iteration that raised it (the last record also carries whatever was raised after its bookkeeping)
                r.status = flags_to_code(s + 1 == done ? fl : c->h_flag_hist[(size_t)s]);
            } else {
                r.iter = -1;
            }
        }
    }
    if (code) fail(code, "device raised error flags 0x%x during the batch", fl);
    return code;
}

}  // namespace

extern "C" {

const char* par2d_last_error(void) { return g_err.c_str(); }

void par2d_default_config(par2d_config* c) {
    // FV/Indexing.h:12-20, FV/main.cpp:77-94, Thermo.h:25,46 and the AIR record of
    // FV/therm_air.dat in the order Thermo.h:27-29 reads it.
    static const double hi[7] = {3.08792717E+00, 1.24597184E-03, -4.23718945E-07, 6.74774789E-11, -3.97076972E-15, -9.95262755E+02, 5.95960930E+00};
    static const double lo[7] = {3.56839620E+00, -6.78729429E-04, 1.55371476E-06, -3.29937060E-12, -4.66395387E-13, -1.06234659E+03, 3.71582965E+00};
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

int par2d_create(const par2d_config* cfg, int32_t nblocks, const par2d_block_desc* blocks, par2d_ctx** out) {
    if (!cfg || !blocks || !out || nblocks < 1) return fail(PAR2D_ERR_ARG, "create: null argument");
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail(PAR2D_ERR_ARG, "create: bad rank %d of %d", cfg->rank, cfg->nranks);
    if (cfg->mode != PAR2D_MODE_FAST && cfg->mode != PAR2D_MODE_FAITHFUL) return fail(PAR2D_ERR_ARG, "create: unknown mode %d", cfg->mode);
    if (cfg->accur != 0 && cfg->accur != 1) return fail(PAR2D_ERR_ARG, "create: accur must be 0 or 1 (SpatialDiscretization.cpp:204)");
    int rc = validate_table(nblocks, blocks, cfg->nranks);
    if (rc) return rc;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(PAR2D_ERR_CUDA, "no CUDA device available; par2d_b200 has no CPU fallback");
    int dev = cfg->device;
    if (dev < 0) CU_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(PAR2D_ERR_CUDA, "device %d requested but only %d present", dev, ndev);
    CU_TRY(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(PAR2D_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", dev, prop.major, prop.minor);

    par2d_ctx* c = new par2d_ctx;
    c->cfg = *cfg;
    c->gas = make_gas(*cfg);
    c->tpg = !(cfg->igam > 0);
    c->axi = cfg->iaxi == 1;
    c->kt = cfg->mode == PAR2D_MODE_FAITHFUL ? kernels_faithful() : kernels_fast();
    c->dg = cfg->accur == 1;
    c->device = dev;
    c->kt->preload(c->tpg, c->axi);   // every kernel this ctx can launch is resident before anything of it runs (lazy loading, see there)
    preload_common();
    {
        // Rows per warp strip.  A strip is marched by ONE warp, row after row (~1 us per row), so small meshes
        // need short strips to expose enough warps (the shipped cases hold ~50 k cells: 64-row strips would be
        // 40 warps on 148 SMs); large meshes keep the measured optimum of 64 rows, where the per-strip prologue
        // and the two halo rows are amortised.  The rule looks at the GLOBAL mesh only, so the strip
        // decomposition - and with it the summation order of the residual norms - does not depend on how the
        // blocks are spread over ranks (relres stays bit-identical for any number of GPUs).
        long long global_cells = 0;
        for (int g = 0; g < nblocks; g++) global_cells += (long long)(blocks[g].nx - 1) * (blocks[g].ny - 1);
        const long long want_strips = 148ll * 16;   // one full residency of 16 warps per SM of a B200; a constant, so that every rank of a
                                                    // multi-GPU job derives the same strips (and with them the same summation order)
        long long rows = global_cells / (MCELLS * want_strips);
        rows -= rows & 1;
        c->march_rows = (int)std::min<long long>(64, std::max<long long>(2, rows));
        c->global_cells = global_cells;
        // below ~2 M cells a step is a few tens of microseconds: bound by launch and dependency latency, not by HBM
        c->small_mesh = global_cells < (2ll << 20);
        const char* sm = getenv("PAR2D_SMALL_MESH");   // A/B runs: force the choice (0 / 1)
        if (sm) c->small_mesh = atoi(sm) != 0;
    }
    const char* mrows = getenv("PAR2D_MARCH_ROWS");
    if (mrows && atoi(mrows) > 0) c->march_rows = atoi(mrows);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail(PAR2D_ERR_CUDA, "cannot create stream"); }
    c->own_stream = true;

    c->nblocks = nblocks;
    c->table.assign(blocks, blocks + nblocks);
    c->owned_index.assign(nblocks, -1);
    for (int g = 0; g < nblocks; g++) {
        c->table[g].x = c->table[g].y = nullptr;
        c->table[g].ibound = nullptr;
        if (blocks[g].owner == cfg->rank) {
            c->owned_index[g] = (int)c->owned_gid.size();
            c->owned_gid.push_back(g);
        }
    }
    c->nowned = (int)c->owned_gid.size();
    *out = c;  // from here on failures leave a destroyable ctx behind
    if (c->nowned == 0) return fail(PAR2D_ERR_ARG, "rank %d owns no block", cfg->rank);

    // peer buffers: lines ordered by (receiving block id, receiving side)
    const int lw = c->dg ? 8 : 4;   // doubles per edge cell (SpatialDiscretization.cpp:231: ntrans *= (ACCUR + 1))
    c->peers.resize(cfg->nranks);
    for (int p = 0; p < cfg->nranks; p++) {
        if (p == cfg->rank) continue;
        for (int g = 0; g < nblocks; g++)
            for (int s = 0; s < 4; s++) {
                const int nb = blocks[g].bid[s];
                if (nb < 0) continue;
                if (blocks[g].owner == cfg->rank && blocks[nb].owner == p) c->peers[p].n_recv += (long long)lw * desc_edge(blocks[g], s);
                if (blocks[g].owner == p && blocks[nb].owner == cfg->rank) c->peers[p].n_send += (long long)lw * desc_edge(blocks[g], s);
            }
        if (c->peers[p].n_send) { CU_TRY(cudaMalloc(&c->peers[p].d_send, sizeof(double) * c->peers[p].n_send)); CU_TRY(cudaMemset(c->peers[p].d_send, 0, sizeof(double) * c->peers[p].n_send)); }
    }
    if (cfg->nranks > 1) {
        // the receive side lives in one exportable allocation (the arena); with the external transport
        // only parity 0 of it is used
        c->lay = arena_layout(blocks, nblocks, cfg->nranks, cfg->rank, lw);
        CU_TRY(cudaMalloc(&c->arena, c->lay.bytes));
        CU_TRY(cudaMemset(c->arena, 0, c->lay.bytes));
        double* lines0 = reinterpret_cast<double*>(c->arena + c->lay.lines_off);
        for (int p = 0; p < cfg->nranks; p++)
            if (p != cfg->rank && c->peers[p].n_recv) c->peers[p].d_recv = lines0 + recv_segment_off(blocks, nblocks, cfg->rank, p, lw);
        CU_TRY(cudaMalloc(&c->put_ticket, sizeof(unsigned)));
        CU_TRY(cudaMemset(c->put_ticket, 0, sizeof(unsigned)));
        const char* tmo = getenv("PAR2D_PEER_TIMEOUT_MS");
        if (tmo && atoll(tmo) > 0) c->timeout_ns = atoll(tmo) * 1000000ll;
    }

    // per-block device memory
    c->hblk.resize(c->nowned);
    c->block_mem.assign(c->nowned, nullptr);
    c->block_ibound.assign(c->nowned, nullptr);
    c->block_fnorm.assign(c->nowned, nullptr);
    long long max_cells = 0, max_nodes = 0;
    for (int k = 0; k < c->nowned; k++) {
        const int g = c->owned_gid[k];
        const par2d_block_desc& d = blocks[g];
        if (!d.x || !d.y || !d.ibound) return fail(PAR2D_ERR_ARG, "owned block %d has no coordinates / BC codes", g);
        BlockDev& b = c->hblk[k];
        memset(&b, 0, sizeof b);
        b.gid = g;
        b.ncx = d.nx - 1;
        b.ncy = d.ny - 1;
        b.pitch = ((b.ncx + COL0 + 1 + 15) / 16) * 16;
        b.plane = (long long)b.pitch * (b.ncy + 2);
        if (b.plane >= (1ll << 31)) return fail(PAR2D_ERR_ARG, "block %d: %lld doubles per plane; the kernels index a plane with 32 bits", g, b.plane);
        CU_TRY(cudaMalloc(&c->block_mem[k], sizeof(double) * 11 * b.plane));
        CU_TRY(cudaMemsetAsync(c->block_mem[k], 0, sizeof(double) * 11 * b.plane, c->stream));
        b.state[0] = c->block_mem[k];                  // planes 0..3
        b.nodes = c->block_mem[k] + 4 * b.plane;       // planes 4..6: x, y, metric
        b.state[1] = c->block_mem[k] + 7 * b.plane;    // planes 7..10
        const int nfaces = 2 * (b.ncx + b.ncy);
        CU_TRY(cudaMalloc(&c->block_ibound[k], sizeof(int) * nfaces));
        CU_TRY(cudaMemcpyAsync(c->block_ibound[k], d.ibound, sizeof(int) * nfaces, cudaMemcpyHostToDevice, c->stream));
        b.ibound = c->block_ibound[k];
        CU_TRY(cudaMalloc(&c->block_fnorm[k], sizeof(double2) * nfaces));
        b.fnorm = c->block_fnorm[k];
        b.ux = b.uy = b.uxt = b.uyt = nullptr;
        for (int s = 0; s < 4; s++) b.ghost2[s] = nullptr;
        b.dg_tile0 = b.dg_ntiles = 0;
        if (c->dg) {
            const size_t nc4 = 4 * (size_t)b.ncx * b.ncy, ng = 8 * (size_t)(2 * b.ncx + 2 * b.ncy);
            double* m = nullptr;
            CU_TRY(cudaMalloc(&m, sizeof(double) * (4 * nc4 + ng)));
            CU_TRY(cudaMemsetAsync(m, 0, sizeof(double) * (4 * nc4 + ng), c->stream));   // slopes start at zero, main.cpp:199-209
            c->dg_mem.push_back(m);
            b.ux = m; b.uy = m + nc4; b.uxt = m + 2 * nc4; b.uyt = m + 3 * nc4;
            double* gp = m + 4 * nc4;
            b.ghost2[SIDE_BOTTOM] = gp; gp += 8 * (size_t)b.ncx;
            b.ghost2[SIDE_RIGHT] = gp; gp += 8 * (size_t)b.ncy;
            b.ghost2[SIDE_TOP] = gp; gp += 8 * (size_t)b.ncx;
            b.ghost2[SIDE_LEFT] = gp;
            b.dg_ntiles = ((b.ncx + DG_SW - 1) / DG_SW) * ((b.ncy + DG_ROWS - 1) / DG_ROWS);   // strips of the cell kernel
            c->dg_max_lim_tiles = std::max(c->dg_max_lim_tiles, (int)(((long long)b.ncx * b.ncy + DG_THREADS - 1) / DG_THREADS));
            b.dg_tile0 = c->dg_total_tiles;
            c->dg_total_tiles += b.dg_ntiles;
            c->dg_max_tiles = std::max(c->dg_max_tiles, b.dg_ntiles);
        }
        b.tiles_x = (b.ncx + MCELLS - 1) / MCELLS;
        b.tiles_y = (b.ncy + c->march_rows - 1) / c->march_rows;
        b.tile0 = c->total_tiles;
        c->total_tiles += b.tiles_x * b.tiles_y;
        c->max_tiles = std::max(c->max_tiles, b.tiles_x * b.tiles_y);
        c->max_faces = std::max(c->max_faces, nfaces);
        c->owned_cells += (long long)b.ncx * b.ncy;
        max_cells = std::max(max_cells, (long long)b.ncx * b.ncy);
        max_nodes = std::max(max_nodes, (long long)d.nx * d.ny);
        for (int s = 0; s < 4; s++) {
            const int nb = d.bid[s];
            b.nb_kind[s] = NB_NONE;
            b.nb_block[s] = -1;
            b.nb_side[s] = -1;
            b.recv[s] = nullptr;
            if (nb < 0) continue;
            b.nb_side[s] = facing_side(blocks[nb], g, s);
            if (blocks[nb].owner == cfg->rank) {
                b.nb_kind[s] = NB_LOCAL;
                b.nb_block[s] = c->owned_index[nb];
            } else {
                b.nb_kind[s] = NB_REMOTE;
                // slot offset inside the peer's receive buffer
                const int p = blocks[nb].owner;
                long long off = 0;
                bool found = false;
                for (int gg = 0; gg < nblocks && !found; gg++) {
                    if (blocks[gg].owner != cfg->rank) continue;
                    for (int ss = 0; ss < 4; ss++) {
                        const int n2 = blocks[gg].bid[ss];
                        if (n2 < 0 || blocks[n2].owner != p) continue;
                        if (gg == g && ss == s) { found = true; break; }
                        off += (long long)lw * desc_edge(blocks[gg], ss);
                    }
                }
                b.recv[s] = c->peers[p].d_recv + off;
            }
        }
    }
    c->staging_doubles = std::max(4 * max_cells, 2 * max_nodes);
    CU_TRY(cudaMalloc(&c->staging, sizeof(double) * c->staging_doubles));

    // nodes, min_face, initial state
    std::vector<double> parts(1024);
    double* dparts = nullptr;
    CU_TRY(cudaMalloc(&dparts, sizeof(double) * 1024));
    for (int k = 0; k < c->nowned; k++) {
        const par2d_block_desc& d = blocks[c->owned_gid[k]];
        BlockDev& b = c->hblk[k];
        const long long nn = (long long)d.nx * d.ny;
        CU_TRY(cudaMemcpyAsync(c->staging, d.x, sizeof(double) * nn, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->staging + nn, d.y, sizeof(double) * nn, cudaMemcpyHostToDevice, c->stream));
        launch_nodes_to_planes(c->stream, b, c->staging, c->staging + nn);
        launch_cell_metric(c->stream, b, c->axi ? 1 : 0);
        launch_face_normals(c->stream, b, c->block_fnorm[k]);
        const long long cells = (long long)b.ncx * b.ncy;
        const int nparts = (int)std::min<long long>(1024, (cells + 255) / 256);
        launch_min_face(c->stream, b, dparts, nparts);
        CU_TRY(cudaMemcpyAsync(parts.data(), dparts, sizeof(double) * nparts, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        double m = 1.0;
        for (int q = 0; q < nparts; q++) m = fmin(m, parts[q]);
        b.min_face = m;
        launch_fill_uniform(c->stream, b, 0, cfg->uFS);
    }
    cudaFree(dparts);

    CU_TRY(cudaMalloc(&c->dblk, sizeof(BlockDev) * c->nowned));
    CU_TRY(cudaMemcpyAsync(c->dblk, c->hblk.data(), sizeof(BlockDev) * c->nowned, cudaMemcpyHostToDevice, c->stream));

    // send lines for off-rank neighbours
    std::vector<SendLine> lines;
    for (int p = 0; p < cfg->nranks; p++) {
        if (p == cfg->rank) continue;
        long long off = 0;
        for (int g = 0; g < nblocks; g++) {
            if (blocks[g].owner != p) continue;
            for (int s = 0; s < 4; s++) {
                const int nb = blocks[g].bid[s];
                if (nb < 0 || blocks[nb].owner != cfg->rank) continue;
                SendLine ln;
                ln.block = c->owned_index[nb];
                ln.side = facing_side(blocks[nb], g, s);
                ln.n = desc_edge(blocks[g], s);
                ln.pad = 0;
                ln.dst = c->peers[p].d_send + off;
                ln.parity_stride = 0;
                c->hline_peer.push_back(p);
                c->hline_off.push_back(off);
                off += (long long)lw * ln.n;
                lines.push_back(ln);
                c->max_line = std::max(c->max_line, ln.n);
            }
        }
    }
    c->nlines = (int)lines.size();
    c->hlines = lines;
    if (c->nlines) {
        CU_TRY(cudaMalloc(&c->dlines, sizeof(SendLine) * c->nlines));
        CU_TRY(cudaMemcpyAsync(c->dlines, lines.data(), sizeof(SendLine) * c->nlines, cudaMemcpyHostToDevice, c->stream));
    }

    if (c->arena) {
        c->dt_table = reinterpret_cast<double*>(c->arena + c->lay.dt_off);   // parity 0
    } else {
        CU_TRY(cudaMalloc(&c->dt_table, sizeof(double) * 2 * nblocks));
        CU_TRY(cudaMemsetAsync(c->dt_table, 0, sizeof(double) * 2 * nblocks, c->stream));
    }
    {
        // every strip of every owned block, for the persistent launch (any order)
        std::vector<int2> all;
        for (int k = 0; k < c->nowned; k++)
            for (int t = 0; t < c->hblk[k].tiles_x * c->hblk[k].tiles_y; t++) all.push_back(make_int2(k, t));
        CU_TRY(cudaMalloc(&c->d_all_tiles, sizeof(int2) * all.size()));
        CU_TRY(cudaMemcpyAsync(c->d_all_tiles, all.data(), sizeof(int2) * all.size(), cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        const char* no = getenv("PAR2D_NO_OVERLAP");
        c->fused = !(no && atoi(no) != 0);
        const char* np = getenv("PAR2D_NO_PERSIST");
        c->persist = !(np && atoi(np) != 0);
        const char* nd = getenv("PAR2D_NO_DEFER");
        c->defer = !(nd && atoi(nd) != 0);
        c->coop_capacity = (cfg->nranks == 1 && !c->dg) ? c->kt->coop_capacity(c->tpg, c->axi) : 0;
        const char* pp = getenv("PAR2D_PROFILE_PERSIST");
        if (pp && atoi(pp) != 0) { CU_TRY(cudaMalloc(&c->prof, sizeof(long long) * 12 * 4096)); CU_TRY(cudaMemset(c->prof, 0, sizeof(long long) * 12 * 4096)); }
    }
    {
        // ghost-cell records of the fused step, grouped by the strip that holds the edge cell
        std::vector<std::vector<GhostOp>> per_tile((size_t)c->total_tiles);
        std::vector<int32_t> codes;
        for (int k = 0; k < c->nowned; k++) {
            const BlockDev& b = c->hblk[k];
            const par2d_block_desc& d = blocks[c->owned_gid[k]];
            const int ncx = b.ncx, ncy = b.ncy;
            std::vector<double2> fn((size_t)2 * (ncx + ncy));
            CU_TRY(cudaMemcpy(fn.data(), c->block_fnorm[k], sizeof(double2) * fn.size(), cudaMemcpyDeviceToHost));
            for (int side = 0; side < 4; side++) {
                if (b.nb_kind[side] == NB_REMOTE) continue;          // put by the strips themselves (push_side)
                const int len = (side == SIDE_BOTTOM || side == SIDE_TOP) ? ncx : ncy;
                for (int n = 0; n < len; n++) {
                    int ec, er, gc, gr, f;
                    switch (side) {
                        case SIDE_BOTTOM: ec = COL0 + n; er = 1; gc = COL0 + n; gr = 0; f = n; break;
                        case SIDE_RIGHT: ec = COL0 + ncx - 1; er = n + 1; gc = COL0 + ncx; gr = n + 1; f = ncx + n; break;
                        case SIDE_TOP: ec = COL0 + n; er = ncy; gc = COL0 + n; gr = ncy + 1; f = 2 * ncx + ncy - 1 - n; break;
                        default: ec = COL0; er = n + 1; gc = COL0 - 1; gr = n + 1; f = 2 * ncx + 2 * ncy - 1 - n; break;
                    }
                    GhostOp op;
                    memset(&op, 0, sizeof op);
                    const long long io = (long long)er * b.pitch + ec;
                    op.src[0] = b.state[0] + io; op.src[1] = b.state[1] + io;
                    op.src_plane = (int)b.plane;
                    op.blk = k; op.f = f; op.code = d.ibound[f];
                    if (b.nb_kind[side] == NB_LOCAL) {
                        const BlockDev& o = c->hblk[b.nb_block[side]];
                        int oc, orow;
                        switch (b.nb_side[side]) {
                            case SIDE_BOTTOM: oc = COL0 + n; orow = 0; break;
                            case SIDE_RIGHT: oc = COL0 + o.ncx; orow = n + 1; break;
                            case SIDE_TOP: oc = COL0 + n; orow = o.ncy + 1; break;
                            default: oc = COL0 - 1; orow = n + 1; break;
                        }
                        const long long go = (long long)orow * o.pitch + oc;
                        op.dst[0] = o.state[0] + go; op.dst[1] = o.state[1] + go;
                        op.dst_plane = (int)o.plane;
                        op.kind = 0;
                    } else {
                        const long long go = (long long)gr * b.pitch + gc;
                        op.dst[0] = b.state[0] + go; op.dst[1] = b.state[1] + go;
                        op.dst_plane = (int)b.plane;
                        int code = op.code;
                        if (code == -1) code = 3;                  // BoundaryConditions.cpp:76-77
                        if (code == 0) code = 4;
                        op.code = code;
                        op.kind = code == 3 ? 0 : code == 4 ? 1 : code == 1 ? 3 : 2;
                        op.nx = fn[(size_t)f].x; op.ny = fn[(size_t)f].y;
                    }
                    int tx, ty;
                    if (side == SIDE_BOTTOM || side == SIDE_TOP) { tx = n / MCELLS; ty = side == SIDE_BOTTOM ? 0 : b.tiles_y - 1; }
                    else { tx = side == SIDE_LEFT ? 0 : b.tiles_x - 1; ty = n / c->march_rows; }
                    per_tile[(size_t)(b.tile0 + ty * b.tiles_x + tx)].push_back(op);
                }
            }
        }
        std::vector<GhostOp> ops;
        std::vector<int2> range((size_t)c->total_tiles);
        for (int t = 0; t < c->total_tiles; t++) {
            range[(size_t)t] = make_int2((int)ops.size(), (int)per_tile[(size_t)t].size());
            ops.insert(ops.end(), per_tile[(size_t)t].begin(), per_tile[(size_t)t].end());
        }
        CU_TRY(cudaMalloc(&c->d_ops, sizeof(GhostOp) * std::max<size_t>(ops.size(), 1)));
        CU_TRY(cudaMalloc(&c->d_op_range, sizeof(int2) * range.size()));
        CU_TRY(cudaMemcpy(c->d_ops, ops.data(), sizeof(GhostOp) * ops.size(), cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(c->d_op_range, range.data(), sizeof(int2) * range.size(), cudaMemcpyHostToDevice));
    }
    c->partials_stride = 4ll * c->total_tiles;
    CU_TRY(cudaMalloc(&c->gas_dev, sizeof(Gas)));
    CU_TRY(cudaMemcpyAsync(c->gas_dev, &c->gas, sizeof(Gas), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMalloc(&c->grid_bar, sizeof(unsigned) * BAR_WORDS));
    CU_TRY(cudaMalloc(&c->go, sizeof(int)));
    CU_TRY(cudaMalloc(&c->stop_iter, sizeof(long long)));
    CU_TRY(cudaMemsetAsync(c->go, 0, sizeof(int), c->stream));
    CU_TRY(cudaMemsetAsync(c->stop_iter, 0xff, sizeof(long long), c->stream));
    CU_TRY(cudaMalloc(&c->partials, sizeof(double) * 2 * (size_t)c->partials_stride));
    if (c->dg) CU_TRY(cudaMalloc(&c->dg_partials, sizeof(double) * 12 * (size_t)c->dg_total_tiles));
    CU_TRY(cudaMalloc(&c->flags, sizeof(unsigned)));
    CU_TRY(cudaMalloc(&c->stop, sizeof(int)));
    CU_TRY(cudaMalloc(&c->d_iter, sizeof(long long)));
    CU_TRY(cudaMalloc(&c->hist_pos, sizeof(int)));
    CU_TRY(cudaMalloc(&c->ticket, sizeof(unsigned)));
    CU_TRY(cudaMalloc(&c->res0, sizeof(double) * 4 * c->nowned));
    CU_TRY(cudaMalloc(&c->blk_norm, sizeof(double) * 4 * c->nowned));
    CU_TRY(cudaMalloc(&c->dt_used, sizeof(double) * 2));
    CU_TRY(cudaMemsetAsync(c->dt_used, 0, sizeof(double) * 2, c->stream));
    CU_TRY(cudaMemsetAsync(c->flags, 0, sizeof(unsigned), c->stream));
    CU_TRY(cudaMemsetAsync(c->stop, 0, sizeof(int), c->stream));
    CU_TRY(cudaMemsetAsync(c->d_iter, 0, sizeof(long long), c->stream));
    CU_TRY(cudaMemsetAsync(c->hist_pos, 0, sizeof(int), c->stream));
    CU_TRY(cudaMemsetAsync(c->ticket, 0, sizeof(unsigned), c->stream));
    CU_TRY(cudaMemsetAsync(c->res0, 0, sizeof(double) * 4 * c->nowned, c->stream));
    rc = ensure_hist(c, 64);
    if (rc) return rc;
    rc = build_tensor_maps(c);
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int par2d_destroy(par2d_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (double* p : c->block_mem) cudaFree(p);
    for (int* p : c->block_ibound) cudaFree(p);
    for (double2* p : c->block_fnorm) cudaFree(p);
    par2d_peer_disconnect(c);
    for (Peer& p : c->peers) cudaFree(p.d_send);
    if (c->arena) cudaFree(c->arena); else cudaFree(c->dt_table);
    cudaFree(c->put_ticket);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (int q = 0; q < 2; q++) { if (c->ev_ready[q]) cudaEventDestroy(c->ev_ready[q]); if (c->ev_free[q]) cudaEventDestroy(c->ev_free[q]); }
    cudaFree(c->staging2);
    cudaFree(c->prof); cudaFree(c->d_ops); cudaFree(c->d_op_range);
    cudaFree(c->ring_partials); cudaFree(c->ring_dt); cudaFree(c->ring_flags); cudaFree(c->ring_norm); cudaFree(c->ring_r0);
    cudaFree(c->d_all_tiles); cudaFree(c->gas_dev); cudaFree(c->grid_bar); cudaFree(c->go); cudaFree(c->stop_iter); cudaFree(c->flag_hist);
    for (double* m : c->dg_mem) cudaFree(m);
    cudaFree(c->dg_partials);
    cudaFree(c->dblk); cudaFree(c->dtmaps); cudaFree(c->partials); cudaFree(c->flags);
    cudaFree(c->stop); cudaFree(c->d_iter); cudaFree(c->hist_pos); cudaFree(c->ticket); cudaFree(c->res0);
    cudaFree(c->blk_norm); cudaFree(c->dt_used); cudaFree(c->blk_hist); cudaFree(c->step_hist); cudaFree(c->staging); cudaFree(c->dlines);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

static int owned_k(par2d_ctx* c, int block, const char* what) {
    if (!c) { fail(PAR2D_ERR_ARG, "%s: null ctx", what); return -1; }
    if (block < 0 || block >= c->nblocks || c->owned_index[block] < 0) { fail(PAR2D_ERR_ARG, "%s: block %d is not owned by rank %d", what, block, c->cfg.rank); return -1; }
    return c->owned_index[block];
}

int par2d_set_state(par2d_ctx* c, int32_t block, const double* unk) {
    const int k = owned_k(c, block, "set_state");
    if (k < 0 || !unk) return PAR2D_ERR_ARG;
    if (use_device(c)) return PAR2D_ERR_CUDA;
    const BlockDev& b = c->hblk[k];
    const size_t bytes = sizeof(double) * 4 * (size_t)b.ncx * b.ncy;
    CU_TRY(cudaMemcpyAsync(c->staging, unk, bytes, cudaMemcpyHostToDevice, c->stream));
    launch_aos_to_planes(c->stream, b, c->cur, c->staging);
    CU_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int par2d_get_state(par2d_ctx* c, int32_t block, double* unk) {
    const int k = owned_k(c, block, "get_state");
    if (k < 0 || !unk) return PAR2D_ERR_ARG;
    if (use_device(c)) return PAR2D_ERR_CUDA;
    const BlockDev& b = c->hblk[k];
    const size_t bytes = sizeof(double) * 4 * (size_t)b.ncx * b.ncy;
    launch_planes_to_aos(c->stream, b, c->cur, c->staging);
    CU_TRY(cudaMemcpyAsync(unk, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// Batched state transfer: the host<->device copy of block i+1 (copy stream) overlaps the AoS<->plane transpose of
// block i (ctx stream) through two staging buffers.
static int ensure_copy_pipeline(par2d_ctx* c) {
    if (c->copy_stream) return 0;
    CU_TRY(cudaMalloc(&c->staging2, sizeof(double) * c->staging_doubles));
    CU_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int q = 0; q < 2; q++) {
        CU_TRY(cudaEventCreateWithFlags(&c->ev_ready[q], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_free[q], cudaEventDisableTiming));
    }
    return 0;
}

int par2d_set_states(par2d_ctx* c, int32_t n, const int32_t* blocks, const double* const* unk) {
    if (!c || n < 0 || (n > 0 && (!blocks || !unk))) return fail(PAR2D_ERR_ARG, "set_states: bad argument");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    int rc = ensure_copy_pipeline(c);
    if (rc) return rc;
    double* stg[2] = {c->staging, c->staging2};
    CU_TRY(cudaStreamSynchronize(c->stream));       // the staging buffers are free
    for (int i = 0; i < n; i++) {
        const int k = owned_k(c, blocks[i], "set_states");
        if (k < 0 || !unk[i]) return PAR2D_ERR_ARG;
        const BlockDev& b = c->hblk[k];
        const int q = i & 1;
        if (i >= 2) CU_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_free[q], 0));
        CU_TRY(cudaMemcpyAsync(stg[q], unk[i], sizeof(double) * 4 * (size_t)b.ncx * b.ncy, cudaMemcpyHostToDevice, c->copy_stream));
        CU_TRY(cudaEventRecord(c->ev_ready[q], c->copy_stream));
        CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_ready[q], 0));
        launch_aos_to_planes(c->stream, b, c->cur, stg[q]);
        CU_TRY(cudaEventRecord(c->ev_free[q], c->stream));
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

int par2d_get_states(par2d_ctx* c, int32_t n, const int32_t* blocks, double* const* unk) {
    if (!c || n < 0 || (n > 0 && (!blocks || !unk))) return fail(PAR2D_ERR_ARG, "get_states: bad argument");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    int rc = ensure_copy_pipeline(c);
    if (rc) return rc;
    double* stg[2] = {c->staging, c->staging2};
    CU_TRY(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) {
        const int k = owned_k(c, blocks[i], "get_states");
        if (k < 0 || !unk[i]) return PAR2D_ERR_ARG;
        const BlockDev& b = c->hblk[k];
        const int q = i & 1;
        if (i >= 2) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_free[q], 0));
        launch_planes_to_aos(c->stream, b, c->cur, stg[q]);
        CU_TRY(cudaEventRecord(c->ev_ready[q], c->stream));
        CU_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_ready[q], 0));
        CU_TRY(cudaMemcpyAsync(unk[i], stg[q], sizeof(double) * 4 * (size_t)b.ncx * b.ncy, cudaMemcpyDeviceToHost, c->copy_stream));
        CU_TRY(cudaEventRecord(c->ev_free[q], c->copy_stream));
    }
    CU_TRY(cudaStreamSynchronize(c->copy_stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

int par2d_get_residual(par2d_ctx* c, int32_t block, double* res) {
    const int k = owned_k(c, block, "get_residual");
    if (k < 0 || !res) return PAR2D_ERR_ARG;
    if (use_device(c)) return PAR2D_ERR_CUDA;
    const BlockDev& b = c->hblk[k];
    StepArgs a = c->step_args(0);
    a.res_out = c->staging;
    a.res_block = k;
    // with remote neighbours the receive buffers must already hold this state's edge
    // lines (par2d_step_begin + transport); the dt table is rebuilt here either way
    // peer-mapped mode: the arenas hold the lines and dt of the state left by the last par2d_step
    // (after par2d_set_state, call par2d_step(ctx, 0, 0, NULL) on every rank to refresh them)
    flush_finalize(c);
    CU_TRY(cudaMemsetAsync(c->stop, 0, sizeof(int), c->stream));   // a batch that stopped on the tolerance leaves the stop word set
    c->launches++;
    c->kt->ghost(c->stream, a, c->gas, c->peer_connected ? 0 : 1, nullptr);
    launch_step(c, a);
    const size_t bytes = sizeof(double) * 4 * (size_t)b.ncx * b.ncy;
    CU_TRY(cudaMemcpyAsync(res, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

int par2d_set_slopes(par2d_ctx* c, int32_t block, const double* ux, const double* uy) {
    const int k = owned_k(c, block, "set_slopes");
    if (k < 0 || !ux || !uy) return PAR2D_ERR_ARG;
    if (!c->dg) return fail(PAR2D_ERR_ARG, "set_slopes: the ctx was created with accur = 0");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    const BlockDev& b = c->hblk[k];
    const size_t bytes = sizeof(double) * 4 * (size_t)b.ncx * b.ncy;
    // the reference's AoS layout -> the planes the kernels read
    const double* in[2] = {ux, uy};
    double* out[2] = {b.ux, b.uy};
    for (int q = 0; q < 2; q++) {
        CU_TRY(cudaMemcpyAsync(c->staging, in[q], bytes, cudaMemcpyHostToDevice, c->stream));
        launch_slopes_transpose(c->stream, c->staging, out[q], (long long)b.ncx * b.ncy, 1);
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int par2d_get_slopes(par2d_ctx* c, int32_t block, double* ux, double* uy) {
    const int k = owned_k(c, block, "get_slopes");
    if (k < 0 || !ux || !uy) return PAR2D_ERR_ARG;
    if (!c->dg) return fail(PAR2D_ERR_ARG, "get_slopes: the ctx was created with accur = 0");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    const BlockDev& b = c->hblk[k];
    const size_t bytes = sizeof(double) * 4 * (size_t)b.ncx * b.ncy;
    double* out[2] = {ux, uy};
    const double* in[2] = {b.ux, b.uy};
    for (int q = 0; q < 2; q++) {
        launch_slopes_transpose(c->stream, in[q], c->staging, (long long)b.ncx * b.ncy, 0);
        CU_TRY(cudaMemcpyAsync(out[q], c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int par2d_get_residual_dg(par2d_ctx* c, int32_t block, double* res, double* resx, double* resy) {
    const int k = owned_k(c, block, "get_residual_dg");
    if (k < 0 || !res || !resx || !resy) return PAR2D_ERR_ARG;
    if (!c->dg) return fail(PAR2D_ERR_ARG, "get_residual_dg: the ctx was created with accur = 0");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    const BlockDev& b = c->hblk[k];
    const size_t n4 = 4 * (size_t)b.ncx * b.ncy;
    double* out = nullptr;
    CU_TRY(cudaMalloc(&out, sizeof(double) * 3 * n4));
    StepArgs a = c->step_args(0);
    DgArgs d = dg_args(c, 0);
    d.res_out = out;
    d.res_block = k;
    CU_TRY(cudaMemsetAsync(c->stop, 0, sizeof(int), c->stream));
    c->launches++;
    c->kt->dg_ghost(c->stream, a, c->gas, d, c->peer_connected ? 0 : 1);
    c->launches++;
    c->kt->dg_cell(c->stream, a, c->gas, d, c->tpg, c->axi);
    CU_TRY(cudaMemcpyAsync(res, out, sizeof(double) * n4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(resx, out + n4, sizeof(double) * n4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(resy, out + 2 * n4, sizeof(double) * n4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(out);
    CU_TRY(cudaGetLastError());
    return 0;
}

int par2d_step_async(par2d_ctx* c, int32_t nsteps, int32_t stop_on_tol) {
    if (!c || nsteps < 0) return fail(PAR2D_ERR_ARG, "step: bad argument");
    if (c->cfg.nranks != 1 && !c->peer_connected)
        return fail(PAR2D_ERR_COMM, "step: nranks > 1 needs step_begin/step_end (external transport) or par2d_peer_connect");
    if (c->pending_steps) return fail(PAR2D_ERR_COMM, "step: %d begin/end steps are waiting for step_collect", c->pending_steps);
    if (c->batch_open) return fail(PAR2D_ERR_COMM, "step: the previous par2d_step_async has not been waited for");
    if (stop_on_tol && c->peer_connected && (c->dg || !c->fused))
        return fail(PAR2D_ERR_ARG, "step: the stop test across ranks needs the fused first-order step (accur = 0, PAR2D_NO_OVERLAP unset)");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    int rc = ensure_hist(c, nsteps);
    if (rc) return rc;
    CU_TRY(cudaMemsetAsync(c->hist_pos, 0, sizeof(int), c->stream));
    CU_TRY(cudaMemsetAsync(c->stop, 0, sizeof(int), c->stream));
    c->batch_cur0 = c->cur;
    c->batch_steps = nsteps;
    c->batch_seq0 = c->seq;
    c->batch_stop = stop_on_tol && c->peer_connected;
    // Which step?  (the choice depends on the GLOBAL mesh and the arguments only, so every rank makes the same one)
    //   small mesh, or a stop test across ranks : the fused step - one launch per step, or the whole batch in one
    //                                             persistent launch (single rank) - built for latency
    //   large mesh                              : the tuned launch-per-step kernel behind a ghost launch; across ranks
    //                                             with the edge strips + put overlapped on a side stream
    const bool use_fused = c->fused && !c->dg && (c->small_mesh || (stop_on_tol && c->peer_connected));
    if (c->peer_connected) {
        // collective batch, no host transport
        if (use_fused) {
            // the current state's lines + dt (a launch of its own: the state may come from par2d_set_state), the ghost
            // ring by the pull kernel, then one launch per step; the last step leaves the final state's lines in the arenas
            if (c->batch_stop) CU_TRY(cudaMemsetAsync(c->go, 0, sizeof(int), c->stream));
            queue_put(c);
            StepArgs a0 = c->step_args(1);
            c->launches++;
            c->kt->ghost(c->stream, a0, c->gas, 0, nullptr);
            for (int s = 0; s < nsteps; s++) queue_step_fused(c, s + 1, stop_on_tol);
        } else if (c->fused && !c->dg) {
            queue_put(c);              // the first state's lines; every step then produces the next state's
            for (int s = 0; s < nsteps; s++) queue_step_overlapped(c);
        } else {
            // every step is  put (edge lines + dt_b into the neighbours' arenas, flag) -> ghost (waits for the
            // flags) -> march [-> limiter]; the trailing put leaves the final state's lines for par2d_get_residual
            for (int s = 0; s < nsteps; s++) {
                queue_put(c);
                queue_step(c, 0, false);
            }
            queue_put(c);
        }
    } else if (!use_fused) {
        for (int s = 0; s < nsteps; s++) queue_step(c, stop_on_tol, true);
    } else if (nsteps > 0) {
        // ghost ring + per-block dt of the current state by the pull kernel (it may come from par2d_set_state); every
        // step after that pushes them for the next one
        StepArgs a0 = c->step_args(1);
        c->launches++;
        c->kt->ghost(c->stream, a0, c->gas, 1, nullptr);
        const bool fits = c->persist && c->coop_capacity >= c->total_tiles + c->nowned;
        if (!(fits && queue_batch_persistent(c, nsteps, stop_on_tol)))
            for (int s = 0; s < nsteps; s++) queue_step_fused(c, s + 1, stop_on_tol);
    }
    flush_finalize(c);
    if (c->batch_stop && nsteps > 0) {
        const long long first = std::max<long long>(c->iter, c->iter + nsteps - 2);
        c->launches++;
        c->kt->stop_tail(c->stream, stop_args(c), c->stop, c->flags, first, c->iter + nsteps - 1);
    }
    c->batch_open = true;
    return 0;
}

int par2d_step_wait(par2d_ctx* c, par2d_step_info* info) {
    if (!c) return fail(PAR2D_ERR_ARG, "step_wait: null ctx");
    if (!c->batch_open) return fail(PAR2D_ERR_COMM, "step_wait: no par2d_step_async is in flight");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    c->batch_open = false;
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaGetLastError());
    return harvest(c, c->batch_steps, c->batch_cur0, info, c->batch_steps, c->batch_stop != 0, c->batch_seq0);
}

int par2d_step(par2d_ctx* c, int32_t nsteps, int32_t stop_on_tol, par2d_step_info* info) {
    const int rc = par2d_step_async(c, nsteps, stop_on_tol);
    if (rc) return rc;
    return par2d_step_wait(c, info);
}

int par2d_get_block_history(par2d_ctx* c, int32_t nsteps, double* out) {
    if (!c || !out || nsteps < 0 || nsteps > c->last_steps) return fail(PAR2D_ERR_ARG, "get_block_history: only %d steps recorded", c ? c->last_steps : 0);
    const size_t row = 2 * (size_t)c->nowned;
    memcpy(out, c->h_blk_hist.data() + row * (size_t)(c->last_steps - nsteps), sizeof(double) * row * nsteps);
    return 0;
}

int64_t par2d_iteration(const par2d_ctx* c) { return c ? c->iter : -1; }
int32_t par2d_num_owned(const par2d_ctx* c) { return c ? c->nowned : 0; }
int32_t par2d_owned_block(const par2d_ctx* c, int32_t k) { return (c && k >= 0 && k < c->nowned) ? c->owned_gid[k] : -1; }
int64_t par2d_owned_cells(const par2d_ctx* c) { return c ? c->owned_cells : 0; }

int par2d_exchange_layout(par2d_ctx* c, int32_t peer, int64_t* n_send, int64_t* n_recv, void** dev_send, void** dev_recv) {
    if (!c || peer < 0 || peer >= c->cfg.nranks) return fail(PAR2D_ERR_ARG, "exchange_layout: bad peer");
    const Peer& p = c->peers[peer];
    if (n_send) *n_send = p.n_send;
    if (n_recv) *n_recv = p.n_recv;
    if (dev_send) *dev_send = p.d_send;
    if (dev_recv) *dev_recv = p.d_recv;
    return 0;
}

int par2d_dt_table(par2d_ctx* c, void** dev_table, int32_t* n) {
    if (!c) return fail(PAR2D_ERR_ARG, "dt_table: null ctx");
    if (dev_table) *dev_table = c->dt_table;
    if (n) *n = c->nblocks;
    return 0;
}

int par2d_step_begin(par2d_ctx* c) {
    if (!c) return fail(PAR2D_ERR_ARG, "step_begin: null ctx");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    if (c->pending_steps == 0) {
        int rc = ensure_hist(c, 4096);
        if (rc) return rc;
        CU_TRY(cudaMemsetAsync(c->hist_pos, 0, sizeof(int), c->stream));
        CU_TRY(cudaMemsetAsync(c->stop, 0, sizeof(int), c->stream));
    }
    if (c->pending_steps >= c->hist_cap) return fail(PAR2D_ERR_COMM, "step_begin: %d steps queued without step_collect", c->pending_steps);
    if (c->peer_connected) return fail(PAR2D_ERR_COMM, "step_begin: this ctx exchanges over peer-mapped memory; use par2d_step");
    CU_TRY(cudaMemsetAsync(c->dt_table, 0, sizeof(double) * c->nblocks, c->stream));
    PutArgs put;
    memset(&put, 0, sizeof put);
    put.dg = c->dg ? 1 : 0;
    c->launches++;
    c->kt->pack(c->stream, c->step_args(1), c->gas, c->dlines, c->nlines, c->max_line, put);
    return 0;
}

int par2d_step_end(par2d_ctx* c) {
    if (!c) return fail(PAR2D_ERR_ARG, "step_end: null ctx");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    queue_step(c, 0, false);
    c->pending_steps++;
    return 0;
}

int par2d_step_collect(par2d_ctx* c, int32_t nsteps, par2d_step_info* info) {
    if (!c) return fail(PAR2D_ERR_ARG, "step_collect: null ctx");
    if (nsteps > c->pending_steps) return fail(PAR2D_ERR_ARG, "step_collect: %d requested, %d pending", nsteps, c->pending_steps);
    if (use_device(c)) return PAR2D_ERR_CUDA;
    flush_finalize(c);
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaGetLastError());
    const int queued = c->pending_steps;
    const int cur0 = c->cur ^ (queued & 1);
    std::vector<par2d_step_info> all(queued);
    const int code = harvest(c, queued, cur0, all.data(), queued);
    if (info)
        for (int s = 0; s < nsteps; s++) info[s] = all[queued - nsteps + s];
    c->pending_steps = 0;
    return code;
}

int par2d_peer_export(par2d_ctx* c, void* handle_out) {
    if (!c || !handle_out) return fail(PAR2D_ERR_ARG, "peer_export: null argument");
    if (!c->arena) return fail(PAR2D_ERR_COMM, "peer_export: this ctx was created with nranks == 1");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    static_assert(sizeof(cudaIpcMemHandle_t) == PAR2D_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, c->arena));
    memcpy(handle_out, &h, sizeof h);
    return 0;
}

// Everything that follows from knowing where every rank's arena is mapped (c->peer_base): put destinations of the
// launch-of-its-own put (SendLine) and of the fused step (BlockDev::put_dst), dt / residual tables and flag slots in
// every arena, the tile order of the fused step.
static int finish_connect(par2d_ctx* c) {
    const int nr = c->cfg.nranks, me = c->cfg.rank;
    const int lw = c->dg ? 8 : 4;
    // every rank must cut its blocks into the same strips (the summation order of the residual norms): slot 15 of a
    // rank's own flag group holds its march_rows
    {
        unsigned long long mine = (unsigned long long)c->march_rows;
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(c->arena + c->lay.flags_off) + 16 * me + 15;
        CU_TRY(cudaMemcpy(slot, &mine, sizeof mine, cudaMemcpyHostToDevice));
    }
    std::vector<double*> dt_dst(nr), res_dst(nr);
    std::vector<unsigned long long*> flag_dst(nr), rflag_dst(nr);
    std::vector<ArenaLayout> lay(nr);
    for (int p = 0; p < nr; p++) {
        lay[p] = arena_layout(c->table.data(), c->nblocks, nr, p, lw);
        dt_dst[p] = reinterpret_cast<double*>(c->peer_base[p] + lay[p].dt_off);
        res_dst[p] = reinterpret_cast<double*>(c->peer_base[p] + lay[p].res_off);
        flag_dst[p] = reinterpret_cast<unsigned long long*>(c->peer_base[p] + lay[p].flags_off) + 16 * me;
        rflag_dst[p] = flag_dst[p] + 1;
    }
    std::vector<SendLine> lines = c->hlines;
    for (int k = 0; k < c->nowned; k++)
        for (int sd = 0; sd < 4; sd++) { c->hblk[k].put_dst[sd] = nullptr; c->hblk[k].put_stride[sd] = 0; }
    for (size_t l = 0; l < lines.size(); l++) {
        const int p = c->hline_peer[l];
        lines[l].dst = reinterpret_cast<double*>(c->peer_base[p] + lay[p].lines_off) + recv_segment_off(c->table.data(), c->nblocks, p, me, lw) + c->hline_off[l];
        lines[l].parity_stride = lay[p].lines_stride;
        c->hblk[lines[l].block].put_dst[lines[l].side] = lines[l].dst;
        c->hblk[lines[l].block].put_stride[lines[l].side] = lines[l].parity_stride;
    }
    CU_TRY(cudaMemcpy(c->dblk, c->hblk.data(), sizeof(BlockDev) * c->nowned, cudaMemcpyHostToDevice));
    CU_TRY(cudaMalloc(&c->d_dt_dst, sizeof(double*) * nr));
    CU_TRY(cudaMalloc(&c->d_res_dst, sizeof(double*) * nr));
    CU_TRY(cudaMalloc(&c->d_flag_dst, sizeof(unsigned long long*) * nr));
    CU_TRY(cudaMalloc(&c->d_rflag_dst, sizeof(unsigned long long*) * nr));
    CU_TRY(cudaMemcpy(c->d_dt_dst, dt_dst.data(), sizeof(double*) * nr, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->d_res_dst, res_dst.data(), sizeof(double*) * nr, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->d_flag_dst, flag_dst.data(), sizeof(unsigned long long*) * nr, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->d_rflag_dst, rflag_dst.data(), sizeof(unsigned long long*) * nr, cudaMemcpyHostToDevice));
    if (!lines.empty()) {
        CU_TRY(cudaMalloc(&c->dlines_peer, sizeof(SendLine) * lines.size()));
        CU_TRY(cudaMemcpy(c->dlines_peer, lines.data(), sizeof(SendLine) * lines.size(), cudaMemcpyHostToDevice));
    }
    // tile order of the fused step: the strips that put something (they border another rank, or hold cell (0,0) of a
    // block, whose dt every rank needs) first
    {
        std::vector<int2> first, rest;
        for (int k = 0; k < c->nowned; k++) {
            const BlockDev& b = c->hblk[k];
            for (int t = 0; t < b.tiles_x * b.tiles_y; t++) {
                const int tx = t % b.tiles_x, ty = t / b.tiles_x;
                const bool e = t == 0 || (tx == 0 && b.nb_kind[SIDE_LEFT] == NB_REMOTE) || (tx == b.tiles_x - 1 && b.nb_kind[SIDE_RIGHT] == NB_REMOTE) ||
                               (ty == 0 && b.nb_kind[SIDE_BOTTOM] == NB_REMOTE) || (ty == b.tiles_y - 1 && b.nb_kind[SIDE_TOP] == NB_REMOTE);
                (e ? first : rest).push_back(make_int2(k, t));
            }
        }
        c->n_put_tiles = (unsigned)first.size();
        first.insert(first.end(), rest.begin(), rest.end());
        c->n_push_tiles = (int)first.size();
        CU_TRY(cudaMalloc(&c->d_push_tiles, sizeof(int2) * first.size()));
        CU_TRY(cudaMemcpy(c->d_push_tiles, first.data(), sizeof(int2) * first.size(), cudaMemcpyHostToDevice));
    }
    if (!c->side_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CU_TRY(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi));   // edge strips + put ahead of interior strips
        CU_TRY(cudaEventCreateWithFlags(&c->ev_ghost, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_edge, cudaEventDisableTiming));
    }
    c->peer_connected = true;
    return 0;
}

// second half of a connect, after EVERY rank has finished the first (the host synchronises in between): the strip
// decomposition must be the same everywhere
static int check_peers(par2d_ctx* c) {
    const int nr = c->cfg.nranks, me = c->cfg.rank;
    for (int p = 0; p < nr; p++) {
        if (p == me) continue;
        unsigned long long theirs = 0;
        const unsigned long long* slot = reinterpret_cast<const unsigned long long*>(c->peer_base[p]) + 16 * p + 15;   // flags_off == 0
        CU_TRY(cudaMemcpy(&theirs, slot, sizeof theirs, cudaMemcpyDeviceToHost));
        if (theirs != 0 && theirs != (unsigned long long)c->march_rows)
            return fail(PAR2D_ERR_COMM, "rank %d marches strips of %llu rows, rank %d of %d: the residual norms would be summed in different orders (PAR2D_MARCH_ROWS set on one rank only?)",
                        p, theirs, me, c->march_rows);
    }
    return 0;
}

int par2d_peer_connect(par2d_ctx* c, const void* all_handles) {
    if (!c || !all_handles) return fail(PAR2D_ERR_ARG, "peer_connect: null argument");
    if (!c->arena) return fail(PAR2D_ERR_COMM, "peer_connect: this ctx was created with nranks == 1");
    if (c->peer_connected) return fail(PAR2D_ERR_COMM, "peer_connect: already connected");
    if (c->pending_steps) return fail(PAR2D_ERR_COMM, "peer_connect: begin/end steps are pending");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    const int nr = c->cfg.nranks, me = c->cfg.rank;
    c->peer_base.assign(nr, nullptr);
    c->peer_base[me] = c->arena;
    c->peer_ipc = true;
    for (int p = 0; p < nr; p++) {
        if (p == me) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(all_handles) + (size_t)p * PAR2D_PEER_HANDLE_BYTES, sizeof h);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            par2d_peer_disconnect(c);
            return fail(PAR2D_ERR_COMM, "peer_connect: cannot map the arena of rank %d (%s); no peer access between the GPUs?", p, cudaGetErrorString(e));
        }
        c->peer_base[p] = static_cast<char*>(ptr);
    }
    // The sequence counter is NOT reset: the flag slots of the arenas keep the events of earlier connections, and
    // every rank has counted the same number of them.
    int rc = finish_connect(c);
    if (rc) return rc;
    return check_peers(c);
}

int par2d_peer_check(par2d_ctx* c) {
    if (!c || !c->peer_connected) return fail(PAR2D_ERR_COMM, "peer_check: not connected");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    return check_peers(c);
}

int par2d_peer_snapshot(par2d_ctx* c, uint64_t* flags_out, double* res_out, int64_t* counters_out) {
    if (!c || !c->arena) return fail(PAR2D_ERR_COMM, "peer_snapshot: this ctx has no arena");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    if (flags_out) CU_TRY(cudaMemcpy(flags_out, c->arena + c->lay.flags_off, sizeof(uint64_t) * 16 * c->cfg.nranks, cudaMemcpyDeviceToHost));
    if (res_out) CU_TRY(cudaMemcpy(res_out, c->arena + c->lay.res_off, sizeof(double) * 8 * c->nblocks, cudaMemcpyDeviceToHost));
    if (counters_out) {
        int go = 0, stop = 0, pos = 0;
        long long sit = 0, it = 0;
        CU_TRY(cudaMemcpy(&go, c->go, sizeof go, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(&stop, c->stop, sizeof stop, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(&pos, c->hist_pos, sizeof pos, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(&sit, c->stop_iter, sizeof sit, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(&it, c->d_iter, sizeof it, cudaMemcpyDeviceToHost));
        counters_out[0] = (int64_t)c->seq; counters_out[1] = go; counters_out[2] = stop; counters_out[3] = sit; counters_out[4] = it; counters_out[5] = pos;
    }
    return 0;
}

int par2d_peer_connect_local(par2d_ctx** ctxs, int32_t n) {
    if (!ctxs || n < 2) return fail(PAR2D_ERR_ARG, "peer_connect_local: need at least two contexts");
    for (int r = 0; r < n; r++) {
        par2d_ctx* c = ctxs[r];
        if (!c || !c->arena || c->cfg.nranks != n || c->cfg.rank != r) return fail(PAR2D_ERR_ARG, "peer_connect_local: ctxs[%d] must be rank %d of %d", r, r, n);
        if (c->peer_connected || c->pending_steps) return fail(PAR2D_ERR_COMM, "peer_connect_local: ctxs[%d] is already connected or has steps pending", r);
    }
    for (int r = 0; r < n; r++) {
        par2d_ctx* c = ctxs[r];
        CU_TRY(cudaSetDevice(c->device));
        c->peer_base.assign(n, nullptr);
        c->peer_ipc = false;
        for (int p = 0; p < n; p++) {
            c->peer_base[p] = ctxs[p]->arena;
            if (p == r || ctxs[p]->device == c->device) continue;
            int can = 0;
            CU_TRY(cudaDeviceCanAccessPeer(&can, c->device, ctxs[p]->device));
            if (!can) return fail(PAR2D_ERR_COMM, "peer_connect_local: device %d cannot access device %d", c->device, ctxs[p]->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[p]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(PAR2D_ERR_COMM, "cudaDeviceEnablePeerAccess(%d -> %d): %s", c->device, ctxs[p]->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        const int rc = finish_connect(c);
        if (rc) return rc;
    }
    for (int r = 0; r < n; r++) {
        CU_TRY(cudaSetDevice(ctxs[r]->device));
        const int rc = check_peers(ctxs[r]);
        if (rc) return rc;
    }
    return 0;
}

int par2d_peer_disconnect(par2d_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->peer_ipc)
        for (size_t p = 0; p < c->peer_base.size(); p++)
            if ((int)p != c->cfg.rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
    c->peer_base.clear();
    c->peer_connected = false;
    if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); c->side_stream = nullptr; }
    if (c->ev_ghost) { cudaEventDestroy(c->ev_ghost); c->ev_ghost = nullptr; }
    if (c->ev_edge) { cudaEventDestroy(c->ev_edge); c->ev_edge = nullptr; }
    cudaFree(c->d_push_tiles); cudaFree(c->dlines_peer); cudaFree(c->d_dt_dst); cudaFree(c->d_flag_dst); cudaFree(c->d_res_dst); cudaFree(c->d_rflag_dst);
    c->d_push_tiles = nullptr; c->dlines_peer = nullptr; c->d_dt_dst = nullptr; c->d_flag_dst = nullptr; c->d_res_dst = nullptr; c->d_rflag_dst = nullptr;
    return 0;
}

int par2d_set_stream(par2d_ctx* c, uint64_t stream) {
    if (!c) return fail(PAR2D_ERR_ARG, "set_stream: null ctx");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    if (stream == 0) {
        CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    } else {
        c->stream = reinterpret_cast<cudaStream_t>(stream);
        c->own_stream = false;
    }
    return 0;
}

int par2d_kernel_timing(par2d_ctx* c, int32_t enable) {
    if (!c) return fail(PAR2D_ERR_ARG, "kernel_timing: null ctx");
    c->timing = enable != 0;
    c->kernel_ms = 0.0;
    c->kernel_launches = 0;
    c->launches = 0;
    return 0;
}

int par2d_launch_count(par2d_ctx* c, int64_t* n) {
    if (!c || !n) return fail(PAR2D_ERR_ARG, "launch_count: null argument");
    *n = c->launches;
    c->launches = 0;
    return 0;
}

int par2d_kernel_time(par2d_ctx* c, double* total_ms, int64_t* launches) {
    if (!c) return fail(PAR2D_ERR_ARG, "kernel_time: null ctx");
    if (total_ms) *total_ms = c->kernel_ms;
    if (launches) *launches = c->kernel_launches;
    c->kernel_ms = 0.0;
    c->kernel_launches = 0;
    return 0;
}

int par2d_synchronize(par2d_ctx* c) {
    if (!c) return fail(PAR2D_ERR_ARG, "synchronize: null ctx");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    CU_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- point-wise test hooks ------------------------------------------------

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    int alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8) == cudaSuccess ? 0 : 1; }
};
int test_prolog(const par2d_config* cfg, const KernelTable** kt, Gas* g) {
    if (!cfg) return fail(PAR2D_ERR_ARG, "test hook: null config");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(PAR2D_ERR_CUDA, "no CUDA device available; par2d_b200 has no CPU fallback");
    if (cfg->device >= 0) CU_TRY(cudaSetDevice(cfg->device));
    *kt = cfg->mode == PAR2D_MODE_FAITHFUL ? kernels_faithful() : kernels_fast();
    *g = make_gas(*cfg);
    return 0;
}
}  // namespace

int par2d_test_ldfss(const par2d_config* cfg, int64_t n, const double* face, const double* uL, const double* uR, double* flux, double* parr) {
    const KernelTable* kt; Gas g;
    int rc = test_prolog(cfg, &kt, &g);
    if (rc) return rc;
    DevBuf dface, dl, dr, dflux, dparr;
    if (dface.alloc(32 * n) || dl.alloc(32 * n) || dr.alloc(32 * n) || dflux.alloc(32 * n) || dparr.alloc(8 * n)) return fail(PAR2D_ERR_CUDA, "test_ldfss: out of device memory");
    CU_TRY(cudaMemcpy(dface.p, face, 32 * n, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(dl.p, uL, 32 * n, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(dr.p, uR, 32 * n, cudaMemcpyHostToDevice));
    kt->test_ldfss(0, g, !(cfg->igam > 0), cfg->iaxi == 1, n, (double*)dface.p, (double*)dl.p, (double*)dr.p, (double*)dflux.p, (double*)dparr.p);
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(flux, dflux.p, 32 * n, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(parr, dparr.p, 8 * n, cudaMemcpyDeviceToHost));
    return 0;
}

int par2d_test_boundary_state(const par2d_config* cfg, int64_t n, const int32_t* btype, const double* norm, const double* uInt, double* uGhost) {
    const KernelTable* kt; Gas g;
    int rc = test_prolog(cfg, &kt, &g);
    if (rc) return rc;
    DevBuf dbt, dn, di, dg;
    if (dbt.alloc(4 * n) || dn.alloc(16 * n) || di.alloc(32 * n) || dg.alloc(32 * n)) return fail(PAR2D_ERR_CUDA, "test_boundary_state: out of device memory");
    CU_TRY(cudaMemcpy(dbt.p, btype, 4 * n, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(dn.p, norm, 16 * n, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(di.p, uInt, 32 * n, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(dg.p, uGhost, 32 * n, cudaMemcpyHostToDevice));
    kt->test_boundary(0, g, n, (int*)dbt.p, (double*)dn.p, (double*)di.p, (double*)dg.p);
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(uGhost, dg.p, 32 * n, cudaMemcpyDeviceToHost));
    return 0;
}

int par2d_test_update_state(const par2d_config* cfg, int64_t n, double* unk, double* derived) {
    const KernelTable* kt; Gas g;
    int rc = test_prolog(cfg, &kt, &g);
    if (rc) return rc;
    DevBuf du, dd;
    if (du.alloc(32 * n) || dd.alloc(64 * n)) return fail(PAR2D_ERR_CUDA, "test_update_state: out of device memory");
    CU_TRY(cudaMemcpy(du.p, unk, 32 * n, cudaMemcpyHostToDevice));
    kt->test_update_state(0, g, !(cfg->igam > 0), n, (double*)du.p, (double*)dd.p);
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(unk, du.p, 32 * n, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(derived, dd.p, 64 * n, cudaMemcpyDeviceToHost));
    return 0;
}

int par2d_test_solve(const par2d_config* cfg, int64_t n, const double* dt_ycc, const double* unk, const double* res, double* dv) {
    const KernelTable* kt; Gas g;
    int rc = test_prolog(cfg, &kt, &g);
    if (rc) return rc;
    DevBuf dy, du, dr, dd;
    if (dy.alloc(16 * n) || du.alloc(32 * n) || dr.alloc(32 * n) || dd.alloc(32 * n)) return fail(PAR2D_ERR_CUDA, "test_solve: out of device memory");
    CU_TRY(cudaMemcpy(dy.p, dt_ycc, 16 * n, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(du.p, unk, 32 * n, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(dr.p, res, 32 * n, cudaMemcpyHostToDevice));
    kt->test_solve(0, g, !(cfg->igam > 0), cfg->iaxi == 1, n, (double*)dy.p, (double*)du.p, (double*)dr.p, (double*)dd.p);
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(dv, dd.p, 32 * n, cudaMemcpyDeviceToHost));
    return 0;
}

int par2d_persist_profile(par2d_ctx* c, int64_t* out, int32_t nsteps) {
    if (!c || !out || nsteps < 0 || nsteps > 4096) return fail(PAR2D_ERR_ARG, "persist_profile: bad argument");
    if (!c->prof) return fail(PAR2D_ERR_ARG, "persist_profile: create the ctx with PAR2D_PROFILE_PERSIST=1 in the environment");
    if (use_device(c)) return PAR2D_ERR_CUDA;
    CU_TRY(cudaMemcpy(out, c->prof, sizeof(long long) * 12 * nsteps, cudaMemcpyDeviceToHost));
    return 0;
}

int par2d_device_count(int32_t* n) {
    if (!n) return fail(PAR2D_ERR_ARG, "device_count: null output");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(PAR2D_ERR_CUDA, "no CUDA device available; par2d_b200 has no CPU fallback");
    *n = ndev;
    return 0;
}

int par2d_probe_fp64(int32_t device, double* tflops) {
    if (!tflops) return fail(PAR2D_ERR_ARG, "probe_fp64: null output");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(PAR2D_ERR_CUDA, "no CUDA device available");
    if (device >= 0) CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev;
    CU_TRY(cudaGetDevice(&dev));
    CU_TRY(cudaGetDeviceProperties(&prop, dev));
    const int grid = prop.multiProcessorCount * 8, iters = 2048;
    DevBuf out;
    if (out.alloc(sizeof(double) * grid * 256)) return fail(PAR2D_ERR_CUDA, "probe_fp64: out of device memory");
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    launch_fp64_probe(0, (double*)out.p, grid, iters);
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        CU_TRY(cudaEventRecord(e0, 0));
        launch_fp64_probe(0, (double*)out.p, grid, iters);
        CU_TRY(cudaEventRecord(e1, 0));
        CU_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}

}  // extern "C"
