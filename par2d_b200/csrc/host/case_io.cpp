// case_io.cpp -- host-side (no CUDA) implementation of include/par2d_host.h.
//
// Our own C++ for the file formats and conventions of the reference solver
// (MrPezser/PAR2D): the block files read by read_mesh (FV/FileIO.cpp:13-51),
// the connectivity file parsed in FV/main.cpp:117-142, the thermo record read
// by Thermo::LoadCurveFits (FV/Thermo.h:16-34) and the Tecplot writer
// print_state (FV/FileIO.cpp:78-116).  Nothing here touches the GPU.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>
#include <sys/stat.h>

#include "par2d_host.h"

namespace {

thread_local std::string g_host_err;
int host_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_host_err = buf;
    return code;
}

struct Block {
    int nx = 0, ny = 0;
    std::vector<double> x, y;      // nx*ny, j outer; y already shifted
    std::vector<int32_t> ibound;   // 2(nx-1)+2(ny-1), CCW from bottom-left
    int bid[4] = {-1, -1, -1, -1}; // bottom,right,top,left
    int bbound[4] = {-1, -1, -1, -1};
    int owner = 0;
    // GridPart provenance (par2d_gridpart): partition counts and this block's position in its source block, for
    // the header lines of the block file; and the y coordinates as they stand in the file (without the load shift)
    int npart[2] = {1, 1}, ipart[2] = {0, 0};
    std::vector<double> yfile;
};

bool file_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

// y offset the reference adds to every node on load (FV/FileIO.cpp:41)
constexpr double kYShift = 1e-4;

}  // namespace

struct par2d_case {
    std::vector<Block> blocks;
    int grid_nbx = 0, grid_nby = 0;  // >0 when blocks form a regular nbx x nby arrangement
};

namespace {

// One block file: three "<label>: a b" header lines (the labels differ between
// cases: npart/ipart/ngrid, npartb/..., asdf), the third holding the node
// counts; nx*ny "x y" lines with j outer; the word BC; one code per face.
int read_block_file(const std::string& path, Block& b) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return host_fail(PAR2D_ERR_ARG, "cannot open block file %s", path.c_str());
    int a = 0, c = 0;
    for (int line = 0; line < 3; line++) {
        if (fscanf(f, "%*[^:]: %d %d", &a, &c) != 2) {
            fclose(f);
            return host_fail(PAR2D_ERR_ARG, "bad header line %d in %s", line + 1, path.c_str());
        }
    }
    b.nx = a;
    b.ny = c;
    if (b.nx < 2 || b.ny < 2) {
        fclose(f);
        return host_fail(PAR2D_ERR_ARG, "bad node counts %d x %d in %s", b.nx, b.ny, path.c_str());
    }
    const size_t np = (size_t)b.nx * b.ny;
    b.x.resize(np);
    b.y.resize(np);
    for (size_t p = 0; p < np; p++) {
        if (fscanf(f, "%le %le", &b.x[p], &b.y[p]) != 2) {
            fclose(f);
            return host_fail(PAR2D_ERR_ARG, "short node list in %s", path.c_str());
        }
        b.y[p] += kYShift;
    }
    char word[64];
    if (fscanf(f, "%63s", word) != 1) {
        fclose(f);
        return host_fail(PAR2D_ERR_ARG, "missing BC marker in %s", path.c_str());
    }
    // The reference sizes ibound 2nx+2ny but the files hold one code per face
    // (FV/FileIO.cpp:31,46-48); only the latter are ever indexed.
    const int nfaces = 2 * (b.nx - 1) + 2 * (b.ny - 1);
    b.ibound.assign(nfaces, 0);
    for (int k = 0; k < nfaces; k++) {
        if (fscanf(f, "%d", &b.ibound[k]) != 1) {
            fclose(f);
            return host_fail(PAR2D_ERR_ARG, "short BC list in %s (%d of %d)", path.c_str(), k, nfaces);
        }
    }
    fclose(f);
    return 0;
}

}  // namespace

extern "C" {

const char* par2d_host_last_error(void) { return g_host_err.c_str(); }

int par2d_case_load(const char* dir, par2d_case** out) {
    if (!dir || !out) return host_fail(PAR2D_ERR_ARG, "case_load: null argument");
    const std::string root(dir);
    FILE* fc = fopen((root + "/grid.conn").c_str(), "r");
    if (!fc) return host_fail(PAR2D_ERR_ARG, "cannot open %s/grid.conn", dir);
    int nb = 0;
    if (fscanf(fc, "%d", &nb) != 1 || nb < 1) {
        fclose(fc);
        return host_fail(PAR2D_ERR_ARG, "bad block count in %s/grid.conn", dir);
    }
    std::unique_ptr<par2d_case> cs(new par2d_case);
    cs->blocks.resize(nb);
    // tuple order in the file: left, right, bottom, top  (FV/main.cpp:131-133)
    static const int side_of_tuple[4] = {PAR2D_LEFT, PAR2D_RIGHT, PAR2D_BOTTOM, PAR2D_TOP};
    for (int b = 0; b < nb; b++) {
        int bc[4], id[4];
        if (fscanf(fc, "%*d: (%d,%d) (%d,%d) (%d,%d) (%d,%d)", &bc[0], &id[0], &bc[1], &id[1], &bc[2], &id[2], &bc[3], &id[3]) != 8) {
            fclose(fc);
            return host_fail(PAR2D_ERR_ARG, "bad connectivity line %d in %s/grid.conn", b + 1, dir);
        }
        for (int t = 0; t < 4; t++) {
            cs->blocks[b].bbound[side_of_tuple[t]] = bc[t];
            cs->blocks[b].bid[side_of_tuple[t]] = id[t];
        }
    }
    fclose(fc);

    auto block_path = [&](int n) { return root + "/subgrids/grid.block" + std::to_string(n) + ".dat"; };
    const bool one_based = !file_exists(block_path(0)) && file_exists(block_path(nb));
    for (int b = 0; b < nb; b++) {
        int rc = read_block_file(block_path(b + (one_based ? 1 : 0)), cs->blocks[b]);
        if (rc) return rc;
        if (one_based)
            for (int s = 0; s < 4; s++)
                if (cs->blocks[b].bid[s] >= 0) cs->blocks[b].bid[s] -= 1;
        for (int s = 0; s < 4; s++)
            if (cs->blocks[b].bid[s] >= nb)
                return host_fail(PAR2D_ERR_ARG, "block %d: neighbour id %d out of range", b, cs->blocks[b].bid[s]);
    }
    *out = cs.release();
    return 0;
}

// Node (I,J) of the ramp channel, I in [0,ncx], J in [0,ncy].
static void ramp_node(int ncx, int ncy, int I, int J, double* x, double* y) {
    const double kPi = 3.14159265358979323846;
    const double slope = tan(10.0 * kPi / 180.0);
    const double xv = 3.0 * (double)I / (double)ncx;
    double yb;
    if (xv < 0.25) yb = 0.0;
    else if (xv <= 1.0) yb = slope * (xv - 0.25);
    else yb = slope * 0.75;
    *x = xv;
    *y = yb + (1.0 - yb) * (double)J / (double)ncy + kYShift;
}

int par2d_case_synthetic_ramp(int32_t ncx, int32_t ncy, int32_t nbx, int32_t nby, par2d_case** out) {
    return par2d_case_synthetic_ramp_ranked(ncx, ncy, nbx, nby, 1, -1, out);
}

int par2d_case_synthetic_ramp_ranked(int32_t ncx, int32_t ncy, int32_t nbx, int32_t nby, int32_t nranks, int32_t rank,
                                     par2d_case** out) {
    if (!out || ncx < 1 || ncy < 1 || nbx < 1 || nby < 1 || ncx % nbx || ncy % nby)
        return host_fail(PAR2D_ERR_ARG, "synthetic_ramp: cells must divide evenly into blocks");
    if (nranks < 1 || rank >= nranks) return host_fail(PAR2D_ERR_ARG, "synthetic_ramp: bad rank");
    std::unique_ptr<par2d_case> cs(new par2d_case);
    cs->blocks.resize((size_t)nbx * nby);
    cs->grid_nbx = nbx;
    cs->grid_nby = nby;
    const int bx = ncx / nbx, by = ncy / nby;
    for (Block& b : cs->blocks) { b.nx = bx + 1; b.ny = by + 1; }
    if (int rc = par2d_case_partition(cs.get(), nranks)) return rc;
    for (int jb = 0; jb < nby; jb++)
        for (int ib = 0; ib < nbx; ib++) {
            Block& b = cs->blocks[(size_t)jb * nbx + ib];
            if (rank < 0 || b.owner == rank) {
                b.x.resize((size_t)b.nx * b.ny);
                b.y.resize((size_t)b.nx * b.ny);
                for (int j = 0; j < b.ny; j++)
                    for (int i = 0; i < b.nx; i++)
                        ramp_node(ncx, ncy, ib * bx + i, jb * by + j, &b.x[(size_t)j * b.nx + i], &b.y[(size_t)j * b.nx + i]);
            }
            const int code_left = ib == 0 ? 1 : -1, code_right = ib == nbx - 1 ? 3 : -1;
            const int code_bot = jb == 0 ? 4 : -1, code_top = jb == nby - 1 ? 4 : -1;
            b.ibound.clear();
            b.ibound.insert(b.ibound.end(), bx, code_bot);
            b.ibound.insert(b.ibound.end(), by, code_right);
            b.ibound.insert(b.ibound.end(), bx, code_top);
            b.ibound.insert(b.ibound.end(), by, code_left);
            b.bid[PAR2D_LEFT] = ib > 0 ? jb * nbx + ib - 1 : -1;
            b.bid[PAR2D_RIGHT] = ib < nbx - 1 ? jb * nbx + ib + 1 : -1;
            b.bid[PAR2D_BOTTOM] = jb > 0 ? (jb - 1) * nbx + ib : -1;
            b.bid[PAR2D_TOP] = jb < nby - 1 ? (jb + 1) * nbx + ib : -1;
            b.bbound[PAR2D_LEFT] = code_left;
            b.bbound[PAR2D_RIGHT] = code_right;
            b.bbound[PAR2D_BOTTOM] = code_bot;
            b.bbound[PAR2D_TOP] = code_top;
        }
    *out = cs.release();
    return 0;
}

int par2d_case_split(const par2d_case* single, int32_t nbx, int32_t nby, par2d_case** out) {
    if (!single || !out || single->blocks.size() != 1 || nbx < 1 || nby < 1)
        return host_fail(PAR2D_ERR_ARG, "case_split: need a single-block case");
    const Block& src = single->blocks[0];
    // GridPart: step = ceil(n/part - 1e-8) nodes; sub-block spans nodes
    // [step*k, min(n-1, step*(k+1))] so neighbours share a node line.
    const int istep = (int)ceil((double)src.nx / nbx - 1e-8), jstep = (int)ceil((double)src.ny / nby - 1e-8);
    std::unique_ptr<par2d_case> cs(new par2d_case);
    cs->blocks.resize((size_t)nbx * nby);
    const int ncx = src.nx - 1, ncy = src.ny - 1;
    for (int jb = 0; jb < nby; jb++)
        for (int ib = 0; ib < nbx; ib++) {
            const int i0 = istep * ib, i1 = std::min(src.nx - 1, istep * (ib + 1));
            const int j0 = jstep * jb, j1 = std::min(src.ny - 1, jstep * (jb + 1));
            if (i1 <= i0 || j1 <= j0) return host_fail(PAR2D_ERR_ARG, "case_split: empty sub-block");
            Block& b = cs->blocks[(size_t)jb * nbx + ib];
            b.nx = i1 - i0 + 1;
            b.ny = j1 - j0 + 1;
            b.x.resize((size_t)b.nx * b.ny);
            b.y.resize((size_t)b.nx * b.ny);
            for (int j = 0; j < b.ny; j++)
                for (int i = 0; i < b.nx; i++) {
                    b.x[(size_t)j * b.nx + i] = src.x[(size_t)(j0 + j) * src.nx + i0 + i];
                    b.y[(size_t)j * b.nx + i] = src.y[(size_t)(j0 + j) * src.nx + i0 + i];
                }
            const bool at_l = ib == 0, at_r = ib == nbx - 1, at_b = jb == 0, at_t = jb == nby - 1;
            b.ibound.clear();
            for (int i = i0; i < i1; i++) b.ibound.push_back(at_b ? src.ibound[i] : -1);
            for (int j = j0; j < j1; j++) b.ibound.push_back(at_r ? src.ibound[ncx + j] : -1);
            for (int i = i1 - 1; i >= i0; i--) b.ibound.push_back(at_t ? src.ibound[2 * ncx + ncy - 1 - i] : -1);
            for (int j = j1 - 1; j >= j0; j--) b.ibound.push_back(at_l ? src.ibound[2 * ncx + 2 * ncy - 1 - j] : -1);
            b.bid[PAR2D_LEFT] = at_l ? -1 : jb * nbx + ib - 1;
            b.bid[PAR2D_RIGHT] = at_r ? -1 : jb * nbx + ib + 1;
            b.bid[PAR2D_BOTTOM] = at_b ? -1 : (jb - 1) * nbx + ib;
            b.bid[PAR2D_TOP] = at_t ? -1 : (jb + 1) * nbx + ib;
            b.bbound[PAR2D_LEFT] = at_l ? src.bbound[PAR2D_LEFT] : -1;
            b.bbound[PAR2D_RIGHT] = at_r ? src.bbound[PAR2D_RIGHT] : -1;
            b.bbound[PAR2D_BOTTOM] = at_b ? src.bbound[PAR2D_BOTTOM] : -1;
            b.bbound[PAR2D_TOP] = at_t ? src.bbound[PAR2D_TOP] : -1;
        }
    *out = cs.release();
    return 0;
}

// GridPart (GridPart/main.c): reads <dir>/gridin.conn - the block count, then per source block "ixm ixp iym iyp"
// (a neighbour block >= 0, or minus the boundary-condition code of that side) and "ipart jpart" (:23-40) - and
// <dir>/gridin.dat - per block "ni nj" and ni*nj lines "x y", j outer (:43-74) - and cuts every source block into
// ipart x jpart sub-blocks that share a node line with their neighbours (:132-145).  Sub-blocks are numbered in the
// order (source block, jblock, iblock) (:85-94); a side on the source block's boundary takes the neighbouring
// source block's facing sub-block (:182-226) or the side's code.
// The j split is step = ceil(nj/jpart - 1e-8) nodes (:128).  The i split of the committed source is a hard-coded
// table of node indices, istepmanual[] = {0, 99, 178, 257, 356} (:103-108,134,136; the automatic rule is commented
// out): pass that table as isteps[] to reproduce it, or NULL for the automatic ceil rule in both directions.
int par2d_gridpart(const char* dir, const int32_t* isteps, int32_t nisteps, par2d_case** out) {
    if (!dir || !out) return host_fail(PAR2D_ERR_ARG, "gridpart: null argument");
    const std::string root(dir);
    FILE* fc = fopen((root + "/gridin.conn").c_str(), "r");
    if (!fc) return host_fail(PAR2D_ERR_ARG, "cannot open %s/gridin.conn", dir);
    int nblock = 0;
    if (fscanf(fc, "%d", &nblock) != 1 || nblock < 1) { fclose(fc); return host_fail(PAR2D_ERR_ARG, "bad block count in %s/gridin.conn", dir); }
    std::vector<int> ixm(nblock), ixp(nblock), iym(nblock), iyp(nblock), ipart(nblock), jpart(nblock);
    for (int b = 0; b < nblock; b++) {
        if (fscanf(fc, " %d %d %d %d%*[^\n]", &ixm[b], &ixp[b], &iym[b], &iyp[b]) != 4 || fscanf(fc, " %d %d%*[^\n]", &ipart[b], &jpart[b]) != 2 ||
            ipart[b] < 1 || jpart[b] < 1) {
            fclose(fc);
            return host_fail(PAR2D_ERR_ARG, "bad record of block %d in %s/gridin.conn", b, dir);
        }
    }
    fclose(fc);
    FILE* fg = fopen((root + "/gridin.dat").c_str(), "r");
    if (!fg) return host_fail(PAR2D_ERR_ARG, "cannot open %s/gridin.dat", dir);
    std::vector<int> ni(nblock), nj(nblock);
    std::vector<std::vector<double>> X(nblock), Y(nblock);
    for (int b = 0; b < nblock; b++) {
        if (fscanf(fg, "%d %d", &ni[b], &nj[b]) != 2 || ni[b] < 2 || nj[b] < 2) { fclose(fg); return host_fail(PAR2D_ERR_ARG, "bad node counts of block %d in %s/gridin.dat", b, dir); }
        X[b].resize((size_t)ni[b] * nj[b]);
        Y[b].resize((size_t)ni[b] * nj[b]);
        for (size_t p = 0; p < X[b].size(); p++)
            if (fscanf(fg, "%lE %lE ", &X[b][p], &Y[b][p]) != 2) { fclose(fg); return host_fail(PAR2D_ERR_ARG, "short node list of block %d in %s/gridin.dat", b, dir); }
    }
    fclose(fg);
    // the consistency checks of GridPart/main.c:112-127 (it exits with status 100)
    for (int b = 0; b < nblock; b++) {
        const int nbx[2] = {ixm[b], ixp[b]}, nby[2] = {iym[b], iyp[b]};
        for (int n : nbx)
            if (n >= 0 && (n >= nblock || nj[b] != nj[n] || jpart[b] != jpart[n])) return host_fail(PAR2D_ERR_ARG, "gridpart: blocks %d and %d do not match across their x face", b, n);
        for (int n : nby)
            if (n >= 0 && (n >= nblock || ni[b] != ni[n] || ipart[b] != ipart[n])) return host_fail(PAR2D_ERR_ARG, "gridpart: blocks %d and %d do not match across their y face", b, n);
        if (isteps && (nisteps < ipart[b] + 1 || isteps[ipart[b]] > ni[b] - 1))
            return host_fail(PAR2D_ERR_ARG, "gridpart: the manual i split does not fit block %d (%d nodes, %d parts)", b, ni[b], ipart[b]);
    }
    std::vector<int> first(nblock + 1, 0);    // blkind[b][j][i] = first[b] + j*ipart[b] + i
    for (int b = 0; b < nblock; b++) first[b + 1] = first[b] + ipart[b] * jpart[b];
    auto blkind = [&](int b, int j, int i) { return first[b] + j * ipart[b] + i; };
    std::unique_ptr<par2d_case> cs(new par2d_case);
    cs->blocks.resize((size_t)first[nblock]);
    for (int b = 0; b < nblock; b++) {
        const int istep = (int)ceil((double)ni[b] / ipart[b] - 1e-8), jstep = (int)ceil((double)nj[b] / jpart[b] - 1e-8);
        for (int jb = 0; jb < jpart[b]; jb++)
            for (int ib = 0; ib < ipart[b]; ib++) {
                const int i0 = isteps ? isteps[ib] : istep * ib, i1 = isteps ? isteps[ib + 1] : std::min(ni[b] - 1, istep * (ib + 1));
                const int j0 = jstep * jb, j1 = std::min(nj[b] - 1, jstep * (jb + 1));
                if (i1 <= i0 || j1 <= j0) return host_fail(PAR2D_ERR_ARG, "gridpart: empty sub-block (%d,%d) of block %d", ib, jb, b);
                const int me = blkind(b, jb, ib);
                Block& k = cs->blocks[(size_t)me];
                k.nx = i1 - i0 + 1;
                k.ny = j1 - j0 + 1;
                k.npart[0] = ipart[b]; k.npart[1] = jpart[b];
                k.ipart[0] = ib; k.ipart[1] = jb;
                k.x.resize((size_t)k.nx * k.ny);
                k.y.resize(k.x.size());
                k.yfile.resize(k.x.size());
                for (int j = 0; j < k.ny; j++)
                    for (int i = 0; i < k.nx; i++) {
                        const size_t src = (size_t)(j0 + j) * ni[b] + i0 + i, dst = (size_t)j * k.nx + i;
                        k.x[dst] = X[b][src];
                        k.yfile[dst] = Y[b][src];
                        k.y[dst] = Y[b][src] + kYShift;
                    }
                // (boundary code, neighbour block) per side, GridPart/main.c:170-226
                int code[4] = {-1, -1, -1, -1}, nbk[4];
                nbk[PAR2D_LEFT] = me - 1; nbk[PAR2D_RIGHT] = me + 1; nbk[PAR2D_BOTTOM] = me - ipart[b]; nbk[PAR2D_TOP] = me + ipart[b];
                if (ib == 0) {
                    if (ixm[b] >= 0) nbk[PAR2D_LEFT] = blkind(ixm[b], jb, ipart[ixm[b]] - 1);
                    else { nbk[PAR2D_LEFT] = -1; code[PAR2D_LEFT] = -ixm[b]; }
                }
                if (ib == ipart[b] - 1) {
                    if (ixp[b] >= 0) nbk[PAR2D_RIGHT] = blkind(ixp[b], jb, 0);
                    else { nbk[PAR2D_RIGHT] = -1; code[PAR2D_RIGHT] = -ixp[b]; }
                }
                if (jb == 0) {
                    if (iym[b] >= 0) nbk[PAR2D_BOTTOM] = blkind(iym[b], jpart[iym[b]] - 1, ib);
                    else { nbk[PAR2D_BOTTOM] = -1; code[PAR2D_BOTTOM] = -iym[b]; }
                }
                if (jb == jpart[b] - 1) {
                    if (iyp[b] >= 0) nbk[PAR2D_TOP] = blkind(iyp[b], 0, ib);
                    else { nbk[PAR2D_TOP] = -1; code[PAR2D_TOP] = -iyp[b]; }
                }
                for (int sd = 0; sd < 4; sd++) { k.bid[sd] = nbk[sd]; k.bbound[sd] = code[sd]; }
                // one code per face, counter-clockwise: bottom, right, top, left (:227-261)
                k.ibound.clear();
                k.ibound.insert(k.ibound.end(), k.nx - 1, code[PAR2D_BOTTOM]);
                k.ibound.insert(k.ibound.end(), k.ny - 1, code[PAR2D_RIGHT]);
                k.ibound.insert(k.ibound.end(), k.nx - 1, code[PAR2D_TOP]);
                k.ibound.insert(k.ibound.end(), k.ny - 1, code[PAR2D_LEFT]);
            }
    }
    *out = cs.release();
    return 0;
}

// Writes a case in the byte format of GridPart's output (GridPart/main.c:100,152-166,264-265): grid.conn, and
// subgrids/grid.block<b>.dat with the "npartb/ipartb/ngrid" header lines, coordinates as "%le %le \n" (6 digits -
// what the reference's own case directories hold), the word BC and one code per face.
int par2d_case_write_gridpart(const par2d_case* c, const char* dir) {
    if (!c || !dir) return host_fail(PAR2D_ERR_ARG, "case_write_gridpart: null argument");
    const std::string root(dir);
    mkdir(root.c_str(), 0755);
    mkdir((root + "/subgrids").c_str(), 0755);
    FILE* fc = fopen((root + "/grid.conn").c_str(), "w");
    if (!fc) return host_fail(PAR2D_ERR_ARG, "cannot write %s/grid.conn", dir);
    fprintf(fc, "%d\n", (int)c->blocks.size());
    for (size_t b = 0; b < c->blocks.size(); b++) {
        const Block& k = c->blocks[b];
        fprintf(fc, "%d: (%d,%d) (%d,%d) (%d,%d) (%d,%d)\n", (int)b, k.bbound[PAR2D_LEFT], k.bid[PAR2D_LEFT], k.bbound[PAR2D_RIGHT], k.bid[PAR2D_RIGHT],
                k.bbound[PAR2D_BOTTOM], k.bid[PAR2D_BOTTOM], k.bbound[PAR2D_TOP], k.bid[PAR2D_TOP]);
    }
    fclose(fc);
    for (size_t b = 0; b < c->blocks.size(); b++) {
        const Block& k = c->blocks[b];
        const std::string path = root + "/subgrids/grid.block" + std::to_string((int)b) + ".dat";
        FILE* f = fopen(path.c_str(), "w");
        if (!f) return host_fail(PAR2D_ERR_ARG, "cannot write %s", path.c_str());
        fprintf(f, "npartb: %d %d \nipartb: %d %d \nngrid: %d %d \n", k.npart[0], k.npart[1], k.ipart[0], k.ipart[1], k.nx, k.ny);
        for (size_t p = 0; p < k.x.size(); p++) fprintf(f, "%le %le \n", k.x[p], k.yfile.empty() ? k.y[p] - kYShift : k.yfile[p]);
        fprintf(f, "BC\n");
        for (int code : k.ibound) fprintf(f, "%d\n", code);
        fclose(f);
    }
    return 0;
}

int par2d_case_write(const par2d_case* c, const char* dir, int32_t one_based) {
    if (!c || !dir) return host_fail(PAR2D_ERR_ARG, "case_write: null argument");
    const std::string root(dir);
    mkdir(root.c_str(), 0755);
    mkdir((root + "/subgrids").c_str(), 0755);
    const int base = one_based ? 1 : 0;
    FILE* fc = fopen((root + "/grid.conn").c_str(), "w");
    if (!fc) return host_fail(PAR2D_ERR_ARG, "cannot write %s/grid.conn", dir);
    fprintf(fc, "%d\n", (int)c->blocks.size());
    for (size_t b = 0; b < c->blocks.size(); b++) {
        const Block& k = c->blocks[b];
        auto id = [&](int s) { return k.bid[s] < 0 ? -1 : k.bid[s] + base; };
        fprintf(fc, "%d: (%d,%d) (%d,%d) (%d,%d) (%d,%d)\n", (int)b + base, k.bbound[PAR2D_LEFT], id(PAR2D_LEFT),
                k.bbound[PAR2D_RIGHT], id(PAR2D_RIGHT), k.bbound[PAR2D_BOTTOM], id(PAR2D_BOTTOM), k.bbound[PAR2D_TOP], id(PAR2D_TOP));
    }
    fclose(fc);
    for (size_t b = 0; b < c->blocks.size(); b++) {
        const Block& k = c->blocks[b];
        const std::string path = root + "/subgrids/grid.block" + std::to_string((int)b + base) + ".dat";
        FILE* f = fopen(path.c_str(), "w");
        if (!f) return host_fail(PAR2D_ERR_ARG, "cannot write %s", path.c_str());
        fprintf(f, "npartb: 1 1 \nipartb: 0 0 \nngrid: %d %d \n", k.nx, k.ny);
        // undo the load-time shift with the exact inverse only when it round-trips;
        // %.17g keeps the stored double, the loader then adds 1e-4 again.
        for (size_t p = 0; p < k.x.size(); p++) fprintf(f, "%.17e %.17e \n", k.x[p], k.y[p] - kYShift);
        fprintf(f, "BC\n");
        for (int code : k.ibound) fprintf(f, "%d\n", code);
        fclose(f);
    }
    return 0;
}

int32_t par2d_case_num_blocks(const par2d_case* c) { return c ? (int32_t)c->blocks.size() : 0; }

int64_t par2d_case_num_cells(const par2d_case* c) {
    int64_t n = 0;
    if (c)
        for (const Block& b : c->blocks) n += (int64_t)(b.nx - 1) * (b.ny - 1);
    return n;
}

int par2d_case_block(const par2d_case* c, int32_t b, par2d_block_desc* out) {
    if (!c || !out || b < 0 || b >= (int32_t)c->blocks.size()) return host_fail(PAR2D_ERR_ARG, "case_block: bad index");
    const Block& k = c->blocks[b];
    memset(out, 0, sizeof *out);
    out->nx = k.nx;
    out->ny = k.ny;
    out->x = k.x.empty() ? nullptr : k.x.data();
    out->y = k.y.empty() ? nullptr : k.y.data();
    out->ibound = k.ibound.empty() ? nullptr : k.ibound.data();
    for (int s = 0; s < 4; s++) {
        out->bid[s] = k.bid[s];
        out->bbound[s] = k.bbound[s];
    }
    out->owner = k.owner;
    return 0;
}

int par2d_case_partition(par2d_case* c, int32_t nranks) {
    if (!c || nranks < 1) return host_fail(PAR2D_ERR_ARG, "case_partition: bad argument");
    const int nb = (int)c->blocks.size();
    bool patched = false;
    if (c->grid_nbx > 0 && nb % nranks == 0) {
        // factor nranks = px*py with px | nbx, py | nby, minimising the patch perimeter
        int best_px = 0, best_py = 0;
        long best_cost = -1;
        for (int px = 1; px <= nranks; px++) {
            if (nranks % px) continue;
            const int py = nranks / px;
            if (c->grid_nbx % px || c->grid_nby % py) continue;
            const long cost = (long)(c->grid_nbx / px) + (long)(c->grid_nby / py);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_px = px; best_py = py; }
        }
        if (best_px) {
            const int wx = c->grid_nbx / best_px, wy = c->grid_nby / best_py;
            for (int jb = 0; jb < c->grid_nby; jb++)
                for (int ib = 0; ib < c->grid_nbx; ib++)
                    c->blocks[(size_t)jb * c->grid_nbx + ib].owner = (jb / wy) * best_px + (ib / wx);
            patched = true;
        }
    }
    if (!patched) {
        std::vector<int> order(nb);
        std::iota(order.begin(), order.end(), 0);
        auto cells = [&](int b) { return (long)(c->blocks[b].nx - 1) * (c->blocks[b].ny - 1); };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cells(a) > cells(b); });
        std::vector<long> load(nranks, 0);
        for (int b : order) {
            int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            c->blocks[b].owner = r;
            load[r] += cells(b);
        }
    }
    return 0;
}

void par2d_case_free(par2d_case* c) { delete c; }

int par2d_thermo_load(const char* path, par2d_config* cfg) {
    if (!path || !cfg) return host_fail(PAR2D_ERR_ARG, "thermo_load: null argument");
    FILE* f = fopen(path, "r");
    if (!f) return host_fail(PAR2D_ERR_ARG, "cannot open thermo file %s", path);
    // line 1 is skipped; lines 2-4 hold fourteen 15-column fields (the last one
    // unused) followed by the line number: hi a1..a5 | hi a6,a7, lo a1..a3 | lo a4..a7
    double v[14];
    int got = 0;
    if (fscanf(f, "%*[^\n]\n") != 0) { /* header skipped */ }
    got += fscanf(f, "%15lE%15lE%15lE%15lE%15lE%*d", &v[0], &v[1], &v[2], &v[3], &v[4]);
    got += fscanf(f, "%15lE%15lE%15lE%15lE%15lE%*d", &v[5], &v[6], &v[7], &v[8], &v[9]);
    got += fscanf(f, "%15lE%15lE%15lE%15lE%*15E%*d", &v[10], &v[11], &v[12], &v[13]);
    fclose(f);
    if (got != 14) return host_fail(PAR2D_ERR_ARG, "thermo file %s: expected 14 coefficients, got %d", path, got);
    for (int k = 0; k < 7; k++) {
        cfg->thermo_hi[k] = v[k];
        cfg->thermo_lo[k] = v[7 + k];
    }
    cfg->Rs = 8314.34 / 28.9;
    return 0;
}

int par2d_write_solution(const char* path, const char* title, const par2d_config* cfg, int32_t nx, int32_t ny,
                         const double* x, const double* y, const double* unk) {
    if (!path || !cfg || !x || !y || !unk) return host_fail(PAR2D_ERR_ARG, "write_solution: null argument");
    FILE* f = fopen(path, "w");
    if (!f) return host_fail(PAR2D_ERR_ARG, "cannot write %s", path);
    fprintf(f, "TITLE = \"%s\"\n", title ? title : "");
    fprintf(f, "VARIABLES = \"X\", \"Y\", \"rho\", \"u\", \"v\", \"T\", \"p\", \"c\", \"M\"\n");
    fprintf(f, "ZONE I=%d, J=%d, DATAPACKING=POINT\n", nx - 1, ny - 1);
    const double gam = fabs(cfg->igam);
    for (int j = 0; j < ny - 1; j++)
        for (int i = 0; i < nx - 1; i++) {
            const size_t p1 = (size_t)j * nx + i, p2 = p1 + 1, p4 = p1 + nx, p3 = p4 + 1;
            const double xc = 0.25 * (x[p1] + x[p2] + x[p3] + x[p4]);
            const double yc = 0.25 * (y[p1] + y[p2] + y[p3] + y[p4]);
            const double* u = unk + 4 * ((size_t)j * (nx - 1) + i);
            const double T = fmax(u[3], 201.0);
            const double p = fmax(u[0] * cfg->Rs * T, 1.0);
            const double a = sqrt(gam * cfg->Rs * T);
            const double M = sqrt(u[1] * u[1] + u[2] * u[2]) / a;
            fprintf(f, "%.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e \n", xc, yc, u[0], u[1],
                    u[2], T, p, a, M);
        }
    fclose(f);
    return 0;
}

// print_state_DGP1 (FV/FileIO.cpp:117-217): node values of the DG(P1) representation - every node takes the mean
// of the standard reconstruction unk + xsi*ux + eta*uy of its 1, 2 or 4 adjacent cells at the corner facing it.
int par2d_write_solution_dgp1(const char* path, const char* title, const par2d_config* cfg, int32_t nx, int32_t ny,
                              const double* x, const double* y, const double* unk, const double* ux, const double* uy,
                              int32_t iter, double time) {
    if (!path || !cfg || !x || !y || !unk || !ux || !uy) return host_fail(PAR2D_ERR_ARG, "write_solution_dgp1: null argument");
    FILE* f = fopen(path, "w");
    if (!f) return host_fail(PAR2D_ERR_ARG, "cannot write %s", path);
    fprintf(f, "TITLE = \"%s\"\n", title ? title : "");
    fprintf(f, "VARIABLES = \"X\", \"Y\", \"rho\", \"u\", \"v\", \"T\", \"p\", \"c\", \"M\"\n");
    fprintf(f, "ZONE I=%d, J=%d, DATAPACKING=POINT, STRANDID=%d, SOLUTIONTIME=%lf\n", nx, ny, iter, time);
    const double gam = fabs(cfg->igam);
    const int ncx = nx - 1, ncy = ny - 1;
    auto recon = [&](int ci, int cj, double xsi, double eta, double* out) {
        const size_t e = 4 * ((size_t)cj * ncx + ci);
        for (int k = 0; k < 4; k++) out[k] = unk[e + k] + (xsi * ux[e + k]) + (eta * uy[e + k]);
    };
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            // adjacent cells in the reference's order: pp (i,j), pm (i,j-1), mp (i-1,j), mm (i-1,j-1)
            const int ci[4] = {i, i, i - 1, i - 1}, cj[4] = {j, j - 1, j, j - 1};
            const double xs[4] = {-1.0, -1.0, 1.0, 1.0}, et[4] = {-1.0, 1.0, -1.0, 1.0};
            double t[4][4], u[4];
            int n = 0;
            for (int q = 0; q < 4; q++)
                if (ci[q] >= 0 && ci[q] < ncx && cj[q] >= 0 && cj[q] < ncy) recon(ci[q], cj[q], xs[q], et[q], t[n++]);
            for (int k = 0; k < 4; k++)
                u[k] = n == 1 ? t[0][k] : n == 2 ? 0.5 * (t[0][k] + t[1][k]) : 0.25 * (t[0][k] + t[1][k] + t[2][k] + t[3][k]);
            u[3] = fmax(u[3], 201.0);
            const double p = fmax(u[0] * cfg->Rs * u[3], 1.0);
            const double a = sqrt(gam * cfg->Rs * u[3]);
            const double M = sqrt(u[1] * u[1] + u[2] * u[2]) / a;
            const size_t nd = (size_t)j * nx + i;
            fprintf(f, "%.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e,\t %.16e \n", x[nd], y[nd], u[0], u[1], u[2],
                    u[3], p, a, M);
        }
    fclose(f);
    return 0;
}

// Restart reader (the reference has none): parses a solution.b<N>.tec written by print_state
// (FV/FileIO.cpp:84-115) or by par2d_write_solution back into the state layout unk[(j*(nx-1)+i)*4+k].
// The zone header must match the block's cell counts; %.16e round-trips every double exactly.
int par2d_read_solution(const char* path, int32_t nx, int32_t ny, double* unk, double* xc, double* yc) {
    if (!path || !unk || nx < 2 || ny < 2) return host_fail(PAR2D_ERR_ARG, "read_solution: bad argument");
    FILE* f = fopen(path, "r");
    if (!f) return host_fail(PAR2D_ERR_ARG, "cannot read %s", path);
    char line[1024];
    int zi = -1, zj = -1;
    for (int h = 0; h < 3; h++) {
        if (!fgets(line, sizeof line, f)) { fclose(f); return host_fail(PAR2D_ERR_ARG, "%s: truncated header", path); }
        if (h == 2 && sscanf(line, "ZONE I=%d, J=%d", &zi, &zj) != 2) { fclose(f); return host_fail(PAR2D_ERR_ARG, "%s: no ZONE I=, J= line", path); }
    }
    if (zi != nx - 1 || zj != ny - 1) {
        fclose(f);
        return host_fail(PAR2D_ERR_ARG, "%s: zone is %d x %d cells, the block has %d x %d", path, zi, zj, nx - 1, ny - 1);
    }
    const size_t ncell = (size_t)zi * zj;
    for (size_t c = 0; c < ncell; c++) {
        double v[9];
        if (fscanf(f, " %lf , %lf , %lf , %lf , %lf , %lf , %lf , %lf , %lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8]) != 9) {
            fclose(f);
            return host_fail(PAR2D_ERR_ARG, "%s: cell %zu of %zu unreadable", path, c, ncell);
        }
        if (xc) xc[c] = v[0];
        if (yc) yc[c] = v[1];
        unk[4 * c + 0] = v[2]; unk[4 * c + 1] = v[3]; unk[4 * c + 2] = v[4]; unk[4 * c + 3] = v[5];
    }
    fclose(f);
    return 0;
}

}  // extern "C"
