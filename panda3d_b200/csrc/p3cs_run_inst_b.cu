// Instantiations of the run kernel (p3cs_run.cuh) for modes: kRow48, kRow56.
#include "p3cs_run.cuh"

namespace p3cs {
template void launch_run_mode<kRow48>(const MeshDev &, const RunPlanDev &, const GroupDev &, int, dim3, size_t, int, RunTail, int, cudaStream_t);
template void launch_run_mode<kRow56>(const MeshDev &, const RunPlanDev &, const GroupDev &, int, dim3, size_t, int, RunTail, int, cudaStream_t);
}  // namespace p3cs
