/* oracle/shim/mpi.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Single-header stand-in for the 10 MPI symbols the reference solver uses
 * (/root/reference/FV/main.cpp:68-74,160,270-292,380-392,492-531,557 and
 * /root/reference/FV/SpatialDiscretization.cpp:209-213,259,320,380,441).
 * No MPI implementation exists in this image, so the unmodified reference
 * sources are compiled against this header instead.  "Ranks" are processes
 * forked inside MPI_Init (count from the SHIM_NP environment variable) that
 * talk over a full mesh of AF_UNIX socket pairs created before the fork.
 * All reference traffic is tag 0, blocking and issued in the same order on
 * both ends of every pair, so FIFO-per-pair byte streams are sufficient.
 */
#pragma once
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>

typedef int MPI_Comm;
typedef int MPI_Datatype;               /* value == sizeof(element) */
typedef struct { int unused; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define SHIM_MAX_RANKS 128

struct ShimWorld {
    int rank = 0, size = 1;
    int peer_fd[SHIM_MAX_RANKS];
};
inline ShimWorld& shim_world() { static ShimWorld w; return w; }

static inline void shim_write_all(int fd, const void* buf, size_t n) {
    const char* p = static_cast<const char*>(buf);
    while (n) { ssize_t k = write(fd, p, n); if (k <= 0) _exit(3); p += k; n -= (size_t)k; }
}
static inline void shim_read_all(int fd, void* buf, size_t n) {
    char* p = static_cast<char*>(buf);
    while (n) { ssize_t k = read(fd, p, n); if (k <= 0) _exit(4); p += k; n -= (size_t)k; }
}

static inline int MPI_Init(void*, void*) {
    ShimWorld& w = shim_world();
    const char* np = getenv("SHIM_NP");
    w.size = np ? atoi(np) : 1;
    if (w.size < 1 || w.size > SHIM_MAX_RANKS) { fprintf(stderr, "shim: bad SHIM_NP\n"); exit(2); }
    static int pairs[SHIM_MAX_RANKS][SHIM_MAX_RANKS][2];
    const int bufsz = 4 << 20;
    for (int a = 0; a < w.size; a++)
        for (int b = a + 1; b < w.size; b++) {
            if (socketpair(AF_UNIX, SOCK_STREAM, 0, pairs[a][b]) != 0) { perror("socketpair"); exit(2); }
            for (int e = 0; e < 2; e++) {
                setsockopt(pairs[a][b][e], SOL_SOCKET, SO_SNDBUF, &bufsz, sizeof bufsz);
                setsockopt(pairs[a][b][e], SOL_SOCKET, SO_RCVBUF, &bufsz, sizeof bufsz);
            }
        }
    for (int r = 1; r < w.size; r++) {
        fflush(stdout);
        if (fork() == 0) { w.rank = r; break; }
    }
    for (int a = 0; a < w.size; a++)
        for (int b = a + 1; b < w.size; b++) {
            if (a == w.rank)      { w.peer_fd[b] = pairs[a][b][0]; close(pairs[a][b][1]); }
            else if (b == w.rank) { w.peer_fd[a] = pairs[a][b][1]; close(pairs[a][b][0]); }
            else                  { close(pairs[a][b][0]); close(pairs[a][b][1]); }
        }
    return 0;
}
static inline int MPI_Comm_size(MPI_Comm, int* n) { *n = shim_world().size; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = shim_world().rank; return 0; }
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int, MPI_Comm) {
    shim_write_all(shim_world().peer_fd[dst], b, (size_t)n * (size_t)t); return 0;
}
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int, MPI_Comm, MPI_Status*) {
    shim_read_all(shim_world().peer_fd[src], b, (size_t)n * (size_t)t); return 0;
}
static inline int MPI_Sendrecv(const void* sb, int sn, MPI_Datatype st, int dst, int,
                               void* rb, int rn, MPI_Datatype rt, int src, int,
                               MPI_Comm, MPI_Status*) {
    shim_write_all(shim_world().peer_fd[dst], sb, (size_t)sn * (size_t)st);
    shim_read_all(shim_world().peer_fd[src], rb, (size_t)rn * (size_t)rt);
    return 0;
}
static inline int MPI_Barrier(MPI_Comm) {
    ShimWorld& w = shim_world();
    char token = 1;
    if (w.rank == 0) {
        for (int r = 1; r < w.size; r++) shim_read_all(w.peer_fd[r], &token, 1);
        for (int r = 1; r < w.size; r++) shim_write_all(w.peer_fd[r], &token, 1);
    } else {
        shim_write_all(w.peer_fd[0], &token, 1);
        shim_read_all(w.peer_fd[0], &token, 1);
    }
    return 0;
}
static inline int MPI_Finalize() {
    ShimWorld& w = shim_world();
    fflush(stdout);
    if (w.rank == 0) { int st; while (wait(&st) > 0) {} }
    return 0;
}
