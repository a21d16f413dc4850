/* par2d_host.h -- host-side (no CUDA) helpers that sit ABOVE the par2d_gpu.h
 * seam: PAR2D case-directory reader/writer, thermo file reader, solution
 * writer, synthetic mesh generator and block->rank partitioner.
 *
 * These restate, in our own C++ (par2d_b200/csrc/host/), the I/O conventions of
 * the reference so that its case directories and output tooling keep working:
 *   read_mesh            FV/FileIO.cpp:13-51   (3 header lines, nodes, "BC", codes; y += 1e-4)
 *   grid.conn parse      FV/main.cpp:117-142   (left,right,bottom,top tuples)
 *   Thermo::LoadCurveFits FV/Thermo.h:16-34
 *   print_state          FV/FileIO.cpp:78-116  (Tecplot POINT, %.16e)
 *   GridPart slicing     GridPart/main.c       (neighbours share a node line, ids i-fastest)
 *   MeshGen ramp         MeshGen/main.cpp:26-36,262-267 (10 degree compression ramp)
 */
#ifndef PAR2D_HOST_H
#define PAR2D_HOST_H

#include "par2d_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct par2d_case par2d_case;

/* Reads <dir>/grid.conn and <dir>/subgrids/grid.block<b>.dat.  Accepts both
 * numbering bases found in the reference's case directories: 0-based (FV/Case,
 * Case_rktnoz9, Case_spikefine24) and 1-based (Case_ffstep1, Case_ffstep8);
 * detection rule: grid.block0.dat missing and grid.block<nblocks>.dat present. */
int par2d_case_load(const char* dir, par2d_case** out);

/* Deterministic synthetic multi-block mesh: 10 degree compression-ramp channel
 * x in [0,3], y from the ramp to 1, ncx x ncy cells split into nbx x nby equal
 * blocks (ids jb*nbx+ib), BCs left=1 right=3 bottom=4 top=4.  The +1e-4 y
 * shift is applied, like the loader does. */
int par2d_case_synthetic_ramp(int32_t ncx, int32_t ncy, int32_t nbx, int32_t nby, par2d_case** out);
/* Same mesh, partitioned over nranks (par2d_case_partition rules) with node coordinates
 * generated only for the blocks rank owns (rank < 0: all) -- a 16k x 16k mesh is 4.3 GB of
 * coordinates, each process of a multi-GPU run materialises only its share. */
int par2d_case_synthetic_ramp_ranked(int32_t ncx, int32_t ncy, int32_t nbx, int32_t nby, int32_t nranks, int32_t rank,
                                     par2d_case** out);

/* Splits a single-block case into nbx x nby blocks with GridPart's rules
 * (ceil slicing, shared node lines, conn tuples); used to test block/halo
 * indexing bit-exactly against the unsplit run. */
int par2d_case_split(const par2d_case* single, int32_t nbx, int32_t nby, par2d_case** out);

/* GridPart-compatible decomposer (GridPart/main.c): reads <dir>/gridin.conn (block count; per source block
 * "ixm ixp iym iyp" - neighbour block >= 0 or minus the side's BC code - then "ipart jpart") and <dir>/gridin.dat
 * (per block "ni nj", then ni*nj lines "x y", j outer) and cuts every source block into ipart x jpart sub-blocks
 * that share a node line, numbered (source block, jblock, iblock), conn tuples across source blocks as
 * GridPart/main.c:182-226 resolves them.  j split: ceil(nj/jpart - 1e-8) nodes.  i split: isteps == NULL applies
 * the same rule; the committed GridPart source instead hard-codes the node table {0, 99, 178, 257, 356}
 * (main.c:103-108) - pass it as isteps[nisteps] to reproduce that binary. */
int par2d_gridpart(const char* dir, const int32_t* isteps, int32_t nisteps, par2d_case** out);
/* Writes a case in the byte format of GridPart's own output (header lines npartb/ipartb/ngrid, "%le %le \n"
 * coordinates, grid.conn), 0-based. */
int par2d_case_write_gridpart(const par2d_case* c, const char* dir);

/* Writes grid.conn + subgrids/ in the reference format (y written without the
 * shift, %.16e).  one_based selects the Case_ffstep8 naming. */
int par2d_case_write(const par2d_case* c, const char* dir, int32_t one_based);

int32_t par2d_case_num_blocks(const par2d_case* c);
int64_t par2d_case_num_cells(const par2d_case* c);
/* Fills a descriptor whose pointers stay owned by the case. */
int par2d_case_block(const par2d_case* c, int32_t b, par2d_block_desc* out);
/* Assigns owners 0..nranks-1: contiguous patches when the case is a regular
 * nbx x nby arrangement of equal blocks, otherwise largest-block-first onto
 * the least-loaded rank (ties -> lowest rank).  Deterministic. */
int par2d_case_partition(par2d_case* c, int32_t nranks);
void par2d_case_free(par2d_case* c);

/* Reads the 4-line NASA-7 record like Thermo::LoadCurveFits and fills
 * cfg->thermo_hi/lo and cfg->Rs (Mw = 28.9 hard-coded there, Thermo.h:25). */
int par2d_thermo_load(const char* path, par2d_config* cfg);

/* solution.b<N>.tec writer, byte format of print_state (FV/FileIO.cpp:84-115):
 * cell centroid, rho,u,v,T, p, c, M per cell, j outer. */
int par2d_write_solution(const char* path, const char* title, const par2d_config* cfg,
                         int32_t nx, int32_t ny, const double* x, const double* y, const double* unk);

/* Output of the second-order branch, byte format of print_state_DGP1 (FV/FileIO.cpp:117-217; the reference names
 * the file ../Outputs/time/solution.b%03d.it%05d.tec): rho,u,v,T of the DG(P1) representation averaged at the
 * nx x ny NODES, then p, c, M of that state. */
int par2d_write_solution_dgp1(const char* path, const char* title, const par2d_config* cfg, int32_t nx, int32_t ny,
                              const double* x, const double* y, const double* unk, const double* ux, const double* uy,
                              int32_t iter, double time);

/* Restart reader for the same file format (the reference writes these files, FV/FileIO.cpp:78-116,
 * but cannot read them back): fills unk[(j*(nx-1)+i)*4+k] = rho,u,v,T and, when non-null,
 * the cell centroids xc/yc[(j*(nx-1)+i)].  The zone header must match nx-1 x ny-1 cells. */
int par2d_read_solution(const char* path, int32_t nx, int32_t ny, double* unk, double* xc, double* yc);

const char* par2d_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
