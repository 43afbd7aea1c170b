// fp32 kernels with every runtime option: cone hits, Philox observation noise, trajectory dump,
// near-tie audit channel, single-step plumbing.
#include "bgr_rollout_kernel.cuh"
namespace bgr {
BGR_FOR_ALL_LAWS(BGR_DEFINE_LAUNCH, float, true)
}
