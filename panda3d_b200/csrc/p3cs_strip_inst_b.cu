// Instantiations of the strip kernel (p3cs_strip.cuh) for modes: kFastP, kFastPN, kFastPNVV.
#include "p3cs_strip.cuh"

namespace p3cs {
template void launch_mode<kFastP>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
template void launch_mode<kFastPN>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
template void launch_mode<kFastPNVV>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
}  // namespace p3cs
