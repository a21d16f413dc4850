// halo.cuh -- what surrounds a block: ghost-cell positions, boundary-condition ghost states, the per-block time
// step, and the flag protocol of the peer-mapped (NVLink) exchange.  Shared by the ghost kernel (pull model: a
// launch of its own at the start of a batch and for the test hooks) and by the perimeter strips of the fused step
// kernel (push model, march.cuh), which must produce the same bits: the two arithmetic pieces
// (bc_ghost_fill, block_dt_value) are therefore NOT inlined - both kernels call the same machine code.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "physics.cuh"

namespace par2d {

// (col,row) in the padded arrays of the n-th cell along a side of a block
__device__ __forceinline__ void edge_cell_pos(int ncx, int ncy, int side, int n, int& col, int& row) {
    switch (side) {
        case SIDE_BOTTOM: col = COL0 + n; row = 1; break;
        case SIDE_RIGHT: col = COL0 + ncx - 1; row = n + 1; break;
        case SIDE_TOP: col = COL0 + n; row = ncy; break;
        default: col = COL0; row = n + 1; break;
    }
}
__device__ __forceinline__ void ghost_cell_pos(int ncx, int ncy, int side, int n, int& col, int& row) {
    switch (side) {
        case SIDE_BOTTOM: col = COL0 + n; row = 0; break;
        case SIDE_RIGHT: col = COL0 + ncx; row = n + 1; break;
        case SIDE_TOP: col = COL0 + n; row = ncy + 1; break;
        default: col = COL0 - 1; row = n + 1; break;
    }
}

// ------------------------------------------------------------------ peer-mapped exchange: flags
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// generic-proxy writes to global memory (ghost cells, other CTAs' new state) before the TMA (async proxy) reads them
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Waits until every other rank's flag (slot `slot` of its 16-word group) has reached `need`.  Bounded: a
// peer that never arrives surfaces as FLAG_COMM (PAR2D_ERR_COMM), never as a hung GPU.  One thread.
static __device__ __noinline__ bool wait_peers(const unsigned long long* flags, int nranks, int rank, unsigned long long need, long long timeout_ns, int slot = 0) {
    if (need == 0) return true;
    const unsigned long long t0 = global_ns();
    for (int r = 0; r < nranks; r++) {
        if (r == rank) continue;
        while (ld_acquire_sys(flags + 16 * r + slot) < need) {
            if ((long long)(global_ns() - t0) > timeout_ns) return false;
            __nanosleep(200);
        }
    }
    return true;
}
// The same wait by a whole warp, one peer per lane (the perimeter / every strip of the fused step).
static __device__ __forceinline__ bool wait_peers_warp(const unsigned long long* flags, int nranks, int rank, unsigned long long need, long long timeout_ns) {
    bool ok = true;
    const int lane = threadIdx.x & 31;
    for (int r = lane; r < nranks; r += 32) {
        if (r == rank) continue;
        if (ld_acquire_sys(flags + 16 * r) >= need) continue;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flags + 16 * r) < need) {
            if ((long long)(global_ns() - t0) > timeout_ns) { ok = false; break; }
            __nanosleep(100);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    return ok;
}

// per-block time step from cell (0,0) of state S: find_dt, main.cpp:19-37 (the static min over the face
// lengths is BlockDev::min_face)
#ifndef HALO_NOINLINE
#define HALO_NOINLINE __noinline__
#endif
// The arithmetic is a function of its own (one machine code for every caller: identical bits from the ghost kernel and
// from the fused step), so that a caller on a latency-critical path can request the three values early and use them late.
static __device__ HALO_NOINLINE double block_dt_math(double u, double v, double T, double min_face, double gamRs, double cfl) {
    const double v2 = u * u + v * v;
    const double vmax = sqrt(v2) + sqrt(gamRs * T);
    return cfl * min_face / vmax;
}
static __device__ __forceinline__ double block_dt_value(const double* S, long long plane, int pitch, double min_face, double gamRs, double cfl) {
    const long long o = (long long)pitch + COL0;
    const double u = __ldcg(S + plane + o), v = __ldcg(S + 2 * plane + o), T = __ldcg(S + 3 * plane + o);
    return block_dt_math(u, v, T, min_face, gamRs, cfl);
}

// CCW index of boundary face n of a side = index into ibound[] and fnorm[] (SpatialDiscretization.cpp:27,45,64,84)
__device__ __forceinline__ int face_index(int ncx, int ncy, int side, int n) {
    switch (side) {
        case SIDE_BOTTOM: return n;
        case SIDE_RIGHT: return ncx + n;
        case SIDE_TOP: return 2 * ncx + ncy - 1 - n;
        default: return 2 * ncx + 2 * ncy - 1 - n;
    }
}

// Ghost cell of boundary face n of side `side` of block b, computed from and written into the state buffer S
// (generate_ghost_cells, SpatialDiscretization.cpp:17-99): outward normal (precomputed once, BlockDev::fnorm),
// boundary_state, the ghost's T clamp.  The ghost keeps its present value if the boundary condition leaves it
// unwritten.  Arguments are pointers to GLOBAL memory (a by-reference kernel parameter would be copied to the stack
// by every CTA).  The derived interior quantities are only evaluated for the codes that read them (the far field
// needs the sound speed and the pressure, the Bernoulli inflow the kinetic energy): walls and extrapolation, most
// of every perimeter, cost neither a square root nor a division.
// everything but the two cheap codes: far field (1), back pressure (2), Bernoulli inflow (5), unknown codes
static __device__ __noinline__ void bc_ghost_general(const Gas* gp, const BlockDev* bp, double* S, int code, int f, long long io, long long go) {
    const Gas& g = *gp;
    const BlockDev& b = *bp;
    const double2 nrm = __ldg(b.fnorm + f);
    double ui[4], gh[4];
#pragma unroll
    for (int v = 0; v < 4; v++) { ui[v] = __ldcg(S + v * b.plane + io); gh[v] = __ldcg(S + v * b.plane + go); }
    double v2 = 0.0, p = 0.0, aint = 0.0;
#if defined(PAR2D_LEAN_BC)
    if (code == 1) { p = fmax(ui[0] * g.Rs * ui[3], 1.0); aint = fast_sqrt(g.gamRs * ui[3]); }
#else
    if (code == 1) { p = fmax(ui[0] * g.Rs * ui[3], 1.0); aint = sqrt(g.gamRs * ui[3]); }
#endif
    if (code == 5) v2 = ui[1] * ui[1] + ui[2] * ui[2];
    boundary_state(g, code, nrm.x, nrm.y, ui, aint, p, v2, gh);
#pragma unroll
    for (int v = 0; v < 4; v++) __stcg(S + v * b.plane + go, gh[v]);
}

static __device__ HALO_NOINLINE void bc_ghost_fill(const Gas* gp, const BlockDev* bp, double* S, int side, int n) {
    const BlockDev& b = *bp;
    const int f = face_index(b.ncx, b.ncy, side, n);
    int code = __ldg(b.ibound + f);
    int c, r;
    edge_cell_pos(b.ncx, b.ncy, side, n, c, r);
    const long long io = (long long)r * b.pitch + c;
    ghost_cell_pos(b.ncx, b.ncy, side, n, c, r);
    const long long go = (long long)r * b.pitch + c;
    if (code == -1) code = 3;                      // BoundaryConditions.cpp:76-77
    if (code == 0) code = 4;
    // The two codes that make up most of every perimeter, spelled with never-contracted operations (the same bits in
    // both arithmetic modes, the reference's operation order) and kept short: this code runs once per step and
    // strip, i.e. out of a cold instruction cache - measured, ~12 cycles per instruction instead of ~1
    if (code == 3 || code == 4) {
        double u0 = __ldcg(S + io), u1 = __ldcg(S + b.plane + io), u2 = __ldcg(S + 2 * b.plane + io), u3 = __ldcg(S + 3 * b.plane + io);
        if (code == 4) {                           // slip wall: V - 2 (V.n) n, BoundaryConditions.cpp:79-88
            const double2 nrm = __ldg(b.fnorm + f);
            const double vn = __dadd_rn(__dmul_rn(u1, nrm.x), __dmul_rn(u2, nrm.y));
            const double t = __dmul_rn(2.0, vn);
            const double w1 = __dsub_rn(u1, __dmul_rn(t, nrm.x)), w2 = __dsub_rn(u2, __dmul_rn(t, nrm.y));
            u1 = w1; u2 = w2;
        }
        __stcg(S + go, u0); __stcg(S + b.plane + go, u1); __stcg(S + 2 * b.plane + go, u2); __stcg(S + 3 * b.plane + go, fmax(u3, 201.0));
        return;
    }
    bc_ghost_general(gp, bp, S, code, f, io, go);
}

}  // namespace par2d
