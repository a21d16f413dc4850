// kernels_march.cu -- the tuned launch-per-step kernel of PAR2D_MODE_FAST (march_kernel, march.cuh) in a translation
// unit - i.e. a device module - of its own.
// Why: the kernel the headline number rests on is compiled - and placed by the loader - the same way whatever the
// rest of the library grows into (its row loop is 12.5 KB of straight-line code shared by 16 resident warps per SM;
// A/B runs against the round-1 library showed the two builds equal to within the run-to-run spread once the
// bookkeeping of the ghost launch was off the critical path, profiles/r02_ab_large.txt).
#include "march.cuh"
namespace par2d {

void launch_march_tuned(cudaStream_t s, const StepArgs& a, const Gas& g, bool tpg, bool axi) {
    const dim3 grid = a.tile_list ? dim3(a.tile_count, 1) : dim3(a.max_tiles, a.nowned);   // one warp strip per CTA
    if (a.tile_list && a.tile_count == 0) return;
    if (tpg) { if (axi) march_kernel<false, true, true, false><<<grid, MARCH_THREADS, 0, s>>>(a, g); else march_kernel<false, true, false, false><<<grid, MARCH_THREADS, 0, s>>>(a, g); }
    else { if (axi) march_kernel<false, false, true, false><<<grid, MARCH_THREADS, 0, s>>>(a, g); else march_kernel<false, false, false, false><<<grid, MARCH_THREADS, 0, s>>>(a, g); }
}

// see Launch::preload (kernels.cuh)
void preload_march_tuned() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, (const void*)march_kernel<false, true, true, false>);
    cudaFuncGetAttributes(&fa, (const void*)march_kernel<false, true, false, false>);
    cudaFuncGetAttributes(&fa, (const void*)march_kernel<false, false, true, false>);
    cudaFuncGetAttributes(&fa, (const void*)march_kernel<false, false, false, false>);
    cudaGetLastError();
}

}  // namespace par2d
