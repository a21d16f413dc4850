// common.cuh -- host/device shared descriptors of the par2d_b200 CUDA path.
//
// HBM layout of one grid block (all planes share one shape):
//   rows   : ncy+2   (row 0 = bottom ghost line, rows 1..ncy = cells j=0..ncy-1,
//                     row ncy+1 = top ghost line; nodes j=0..ncy live in rows 1..ncy+1)
//   pitch  : multiple of 16 doubles (128 B); column COL0-1 = left ghost, COL0+i = cell i,
//            COL0+ncx = right ghost; node i lives in column COL0+i
//   planes : state A {rho,u,v,T}, nodes {x, y, metric}, state B {rho,u,v,T} (ping-pong; the 7 planes a step
//            reads are always consecutive: A+nodes or nodes+B), where
//            metric[cell (i,j)] = 1/(vol*ycc) (axisymmetric) or 1/vol (planar), stored at node (i,j)'s slot
// i.e. the reference's AoS unk[(j*(nx-1)+i)*4+k] (FV/Indexing.h:11) becomes four padded SoA
// planes with the ghost ring (SpatialDiscretization.cpp:182-185 uGBot/uGRight/uGTop/uGLeft)
// stored in place, so every cell has four in-array neighbours and one TMA box of 32 columns x
// 4 planes stages a strip row with its two halo columns.
#pragma once
#include <stdint.h>
#include "physics.cuh"

namespace par2d {

constexpr int COL0 = 17;  // column of cell/node i = 0; odd, so that a strip's first staged column COL0+30k-1 is even (TMA 16-byte rule)

// marching kernel (march.cuh): a warp strip updates MCELLS cell columns - MARCH_CPL columns per lane, the
// first and the last of the 32*MARCH_CPL staged columns being halo columns - over StepArgs::march_rows
// cell rows
#ifndef MARCH_CPL
#define MARCH_CPL 1   // measured on B200: 2 columns per lane need ~200 registers (10 warps/SM) and lose to 1 column at 16 warps/SM
#endif
constexpr int MCELLS = 32 * MARCH_CPL - 2;

enum : int { SIDE_BOTTOM = 0, SIDE_RIGHT = 1, SIDE_TOP = 2, SIDE_LEFT = 3 };
enum : int { NB_NONE = 0, NB_LOCAL = 1, NB_REMOTE = 2 };

struct BlockDev {
    int gid;            // global block id
    int ncx, ncy;       // cells
    int pitch;          // doubles per row
    long long plane;    // doubles per plane = pitch*(ncy+2)
    double* state[2];   // 4 planes each
    double* nodes;      // 3 planes: x, y, cell metric
    const int* ibound;  // BC code per boundary face, CCW (SpatialDiscretization.cpp:27,45,64,84)
    const double2* fnorm; // outward unit normal per boundary face, same CCW index (SpatialDiscretization.cpp:33-34,51-52,71-72,
                        // 91-92 over MeshModule.cpp:52-81; the right side takes the normals of column 0, as the reference does)
    int nb_kind[4];     // NB_* per side
    int nb_block[4];    // owned-block index of a same-rank neighbour
    int nb_side[4];     // the neighbour's side that faces us
    double* recv[4];    // AoS line [n*4+k] received from an off-rank neighbour (parity 0 of this rank's arena)
    double* put_dst[4]; // peer-mapped mode: where this block's edge line of a side goes in the off-rank neighbour's arena
    long long put_stride[4];  //   (parity 0) and the doubles between the two parities there; null / 0 otherwise
    int tiles_x, tiles_y; // work decomposition: 30-column strips x row chunks (marching kernel)
    int tile0;          // first tile/strip of this block in the partial-sum array
    double min_face;    // static part of find_dt (main.cpp:26-32)
    // second-order branch (cfg.accur == 1, dg.cuh); null / 0 otherwise
    double *ux, *uy;    // xi / eta slopes, SoA planes [k][j*ncx+i] (the reference's AoS (j*ncx+i)*4+k of main.cpp:169-172 is
                        // transposed by par2d_set_slopes / par2d_get_slopes)
    double *uxt, *uyt;  // damped, not yet limited slopes of the step in flight
    double* ghost2[4];  // two Gauss-point ghost states per boundary face, [(2n+pt)*4+k]
    int dg_tile0, dg_ntiles;  // this block's strips of the cell kernel (one record each in the [.][12] partial-sum array)
};

// One ghost cell of the NEW state, as the fused step's perimeter strips fill it (march.cuh, tile_epilogue): where the
// edge cell sits, where its ghost goes (the block's own ring for a boundary-condition face, the same-rank neighbour's
// ring for a shared face) and what happens in between.  Built once per ctx, grouped by strip: the whole perimeter
// work of a strip is one short loop over its records - it runs once per step out of a cold instruction cache, so
// every instruction it does not have counts (measured on the persistent launch: ~12 cycles per cold instruction).
struct GhostOp {
    double* src[2];          // the edge cell, plane 0, in state buffer 0 / 1
    double* dst[2];          // the ghost cell, plane 0, in state buffer 0 / 1
    int src_plane, dst_plane;  // doubles between planes of the two blocks
    int kind;                // 0 copy (shared face; extrapolation, codes 3 and -1), 1 slip wall (codes 4 and 0), 3 far field (code 1),
                             // 2 any other code
    int blk;                 // kind 2: owned block, CCW face index and BC code
    int f, code;
    int pad[2];
    double nx, ny;           // kind 1: outward unit normal (BlockDev::fnorm)
};
static_assert(sizeof(GhostOp) == 80 && sizeof(GhostOp) % 16 == 0, "GhostOp is read with 16-byte loads");

// one edge line to pack for an off-rank neighbour (SpatialDiscretization.cpp:236-257)
struct SendLine {
    int block;   // owned-block index
    int side;
    int n;       // edge cells
    int pad;
    double* dst; // AoS [n*4+k] in the peer send buffer (or, peer-mapped, in the peer's arena, parity 0)
    long long parity_stride;  // doubles between the two parities of the destination (peer-mapped mode), else 0
};

// Peer-mapped exchange (NVLink P2P): every rank owns one "arena" (a single cudaMalloc, exported by
// cudaIpcGetMemHandle) that its neighbours write into:
//   flags[nranks][16]   u64, slot src*16 = number of exchange events rank src has completed into this arena
//   dt[2][nblocks_pad]  per-block time steps of the event's state, written by each block's owner
//   lines[2][total]     received edge lines, per source rank in rank order, per (receiving block, side)
// Event e (1-based) lives in parity e&1.  A producer may write event e only after every peer has produced
// event e-1 (which, in stream order, follows that peer's consumption of event e-2 from the same parity).
struct PutArgs {
    int peer_mode;                       // 0: external transport (plain pack into local send buffers)
    int parity;                          // parity written by this event
    unsigned long long seq;              // event number being produced
    double* const* dt_dst;               // [nranks] dt table (parity 0) of every rank's arena, own included
    long long dt_stride;                 // doubles between dt parities
    unsigned long long* const* flag_dst; // [nranks] this rank's flag slot in every peer's arena (own: unused)
    unsigned* ticket;                    // CTA completion counter of the producing kernel
    const unsigned long long* local_flags; // this rank's own flag array (for the pre-wait)
    int nranks, rank;
    long long timeout_ns;
    unsigned expected;                   // fused step: CTAs of the launch that put something (the last one raises the flags)
    int dg;                              // second-order branch: a line carries the reconstruction at the two Gauss points
                                         // of every edge face, 8 doubles per edge cell (SpatialDiscretization.cpp:243-256)
};

// Stop test across ranks (main.cpp:506-531 breaks on the GLOBAL relres): every rank's bookkeeping publishes its
// per-block sums of iteration `it` into slot (it+1)&3 of every arena's residual table and raises its residual flag
// (slot src*16+1 of the flag array) to it+1; the first CTA of launch m adds the blocks of iteration m-2 in global
// block order - identical bits on every rank - and sets the stop word before any strip of that launch runs.
struct StopArgs {
    int enabled;                          // 0: no test between the steps of this batch
    int nranks, rank;
    double* const* res_dst;               // [nranks] residual table (slot 0) of every arena, own included
    unsigned long long* const* rflag_dst; // [nranks] this rank's residual-flag slot in every arena
    const double* res_table;              // this rank's own table: [4 slots][nblocks_global][2]
    const unsigned long long* local_flags;// this rank's flag array
    int nblocks_global;
    int* go;                              // launch number whose stop decision has been made (strips wait for it)
    long long* stop_iter;                 // iteration whose relres fell below the tolerance
    long long iter0;                      // iteration advanced by launch 1 of this batch
    double tol;
    long long timeout_ns;
};

struct StepArgs {
    const BlockDev* blocks;
    int nowned;
    int nblocks_global;
    int cur;                 // state buffer read this step
    int total_tiles;         // sum over owned blocks (size of partials)
    int max_tiles;           // strips of the largest owned block (grid.x of the march kernel)
    int max_faces;           // longest block perimeter (grid.x of the ghost kernel)
    const void* tmaps;       // CUtensorMap[nowned] over each block's 11 planes; boxes of 32 columns x 2 rows x 7 planes
    double* dt_table;        // nblocks_global
    double* dt_used;         // the global dt this step advances with (read by finalize; written by the march kernel, or by
                             // the ghost kernel when dt_ready)
    int dt_ready;            // the dt table is complete BEFORE the ghost launch (multi-rank modes): the ghost kernel reduces
                             // it into *dt_used and the march kernel reads that one value - the table may already be
                             // receiving the next event's entries while interior strips still run
    double* partials;        // [total_tiles][4] sum of res_k^2
    unsigned* flags;         // device error bits
    const int* stop;         // non-zero -> batch already converged, kernels return at once
    double* res_out;         // optional AoS residual dump (get_residual test hook)
    int res_block;           // owned index whose residual is dumped, or -1
    int update;              // 0: residual only (test hook), 1: advance the state
    int march_rows;          // cell rows per strip of the marching kernel
    const int2* tile_list;   // optional subset launch of the march kernel: (owned block, strip) per CTA, grid.x = tile_count;
    int tile_count;          // null: grid = (max_tiles, nowned), every strip of every block
    // peer-mapped mode: the ghost kernel consumes exchange event wait_seq from this rank's arena
    const unsigned long long* wait_flags;  // null: nothing to wait for
    unsigned long long wait_seq;
    long long recv_off;      // doubles added to BlockDev::recv[] (parity of the consumed event)
    int nranks, rank;
    long long timeout_ns;
    // fused step (march.cuh, PUSH instantiation): the strips on a block's perimeter write the NEW state's ghost ring
    // (boundary-condition faces, same-rank neighbours) and put the edge lines of off-rank faces + the block's dt
    // straight into the peers' arenas, so a step is ONE launch
    double* dt_next;         // per-block dt table of the NEXT step (single rank: the other parity of dt_table)
    int nfin;                // leading CTAs of the launch that run the previous step's bookkeeping (0: none pending)
    int launch_no;           // 1-based launch number within the batch (stop test across ranks)
    int coop_steps;          // persistent launch: steps advanced inside the kernel (else 1)
    unsigned* grid_bar;      // persistent launch: grid barrier words (BAR_WORDS)
    const Gas* gas_dev;      // the Gas constants in global memory (for the non-inlined boundary-condition code)
    long long* prof;         // diagnostics of the persistent launch (null in production), see march_persist_kernel
    const GhostOp* ops;      // fused step: ghost-cell records of this rank, grouped by strip
    const int2* op_range;    // [total_tiles] (first record, count) of strip tile0 + s
    // (new members go at the END: the offsets of the members above are part of what the tuned launch-per-step kernel was
    //  compiled against - ptxas schedules its row loop differently when they move, see march.cuh)
};

struct FinalizeArgs {
    const BlockDev* blocks;
    int nowned;
    const double* partials;
    double* res0;            // [nowned][4] iteration-0 norms (main.cpp:260,365-374)
    double* blk_hist;        // [hist_cap][nowned][2]
    double* step_hist;       // [hist_cap][4] = dt, ressumsum, res0sum, relres
    long long* iter;         // device iteration counter
    int* hist_pos;           // next history slot
    int* stop;
    unsigned* ticket;        // last-block-done counter
    const double* dt_used;   // StepArgs::dt_used of the step being finalised
    double* blk_norm;        // [nowned][4] per-block residual norms (scratch between the block CTAs and the last one)
    double tol;
    int stop_on_tol;
    int hist_cap;
    const double* dg_partials;  // second-order branch: [dg tiles][12] sums of squares of dudt, duxdt, duydt (else null)
    unsigned* flag_hist;     // [hist_cap] device error bits as of each step's bookkeeping (per-step status)
    const unsigned* flags;
    StopArgs stop_x;         // stop test across ranks (peer-mapped mode); enabled = 0 otherwise
};

// Persistent launch without a stop test: the bookkeeping is taken out of the step loop.  The strips keep every step's
// sums, dt and an error-bit snapshot; three short kernels after the launch turn them into the history records
// (kernels_common.cu, launch_finalize_deferred) - same arithmetic, same order, identical bits.
struct DeferArgs {
    int on;
    long long stride;        // doubles between the strip sums of consecutive steps (4 x total strips)
    double* partials;        // [steps][total strips][4]
    double* dt;              // [steps] the global dt of each step
    unsigned* flags;         // [steps] device error bits as of the end of each step
};
constexpr int DEFER_CHUNK = 1024;   // steps per persistent launch in this mode

// grid barrier of the persistent launch (march.cuh, grid_barrier): the arrival counter and 16 release words, one 128-byte line each
constexpr int BAR_RELEASE_WORDS = 16;
constexpr int BAR_WORDS = 32 * (1 + BAR_RELEASE_WORDS);

__host__ __device__ inline int edge_cells(int ncx, int ncy, int side) {
    return (side == SIDE_BOTTOM || side == SIDE_TOP) ? ncx : ncy;
}

}  // namespace par2d
