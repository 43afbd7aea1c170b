// fp64 audit kernels (compiled with --fmad=false: reference operation order, nothing contracted).
#include "bgr_rollout_kernel.cuh"
namespace bgr {
BGR_FOR_ALL_LAWS(BGR_DEFINE_LAUNCH, double, true)
}
