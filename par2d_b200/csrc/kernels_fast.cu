// kernels_fast.cu -- PAR2D_MODE_FAST instantiation: default FMA contraction,
// Newton reciprocals, closed-form per-cell solve (see physics.cuh).
#include "kernels.cuh"
namespace par2d {
const KernelTable* kernels_fast() { return Launch<false>::table(); }
}  // namespace par2d
