// kernels_common.cuh -- launchers of the mode-independent kernels (kernels_common.cu).
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
namespace par2d {
void launch_finalize(cudaStream_t s, const FinalizeArgs& a, double* blk_norm);
void launch_aos_to_planes(cudaStream_t s, const BlockDev& b, int buf, const double* aos);
void launch_planes_to_aos(cudaStream_t s, const BlockDev& b, int buf, double* aos);
void launch_fill_uniform(cudaStream_t s, const BlockDev& b, int buf, const double u[4]);
void launch_nodes_to_planes(cudaStream_t s, const BlockDev& b, const double* x, const double* y);
void launch_cell_metric(cudaStream_t s, const BlockDev& b, int axi);
void launch_min_face(cudaStream_t s, const BlockDev& b, double* out, int nparts);
void launch_fp64_probe(cudaStream_t s, double* out, int grid, int iters);
}  // namespace par2d
