// kernels_faithful.cu -- PAR2D_MODE_FAITHFUL instantiation.  This translation unit is
// compiled with -fmad=false (see __graft_entry__.build) so that no multiply-add is
// contracted: together with IEEE division/sqrt and the reference's operation order the
// state is bit-identical to the reference's CPU path wherever pow() is not involved.
#include "kernels.cuh"
namespace par2d {
const KernelTable* kernels_faithful() { return Launch<true>::table(); }
}  // namespace par2d
