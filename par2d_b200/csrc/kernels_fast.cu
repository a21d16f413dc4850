// kernels_fast.cu -- PAR2D_MODE_FAST instantiation: default FMA contraction,
// Newton reciprocals, closed-form per-cell solve (see physics.cuh).
#define PAR2D_LEAN_BC 1   // subsonic far-field ghosts from one log + two exp instead of two pow (physics.cuh: isentropic_ghost)
#include "kernels.cuh"
namespace par2d {
const KernelTable* kernels_fast() { return Launch<false>::table(); }
}  // namespace par2d
