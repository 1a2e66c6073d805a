// Instantiations of the strip kernel (p3cs_strip.cuh) for modes: kGenericCompact, kGenericFull.
#include "p3cs_strip.cuh"

namespace p3cs {
template void launch_mode<kGenericCompact>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
template void launch_mode<kGenericFull>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
}  // namespace p3cs
