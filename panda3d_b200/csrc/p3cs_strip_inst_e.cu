// Instantiations of the strip kernel (p3cs_strip.cuh) for modes: kAlignedPN, kAlignedPNVV.
#include "p3cs_strip.cuh"

namespace p3cs {
template void launch_mode<kAlignedPN>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
template void launch_mode<kAlignedPNVV>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
}  // namespace p3cs
