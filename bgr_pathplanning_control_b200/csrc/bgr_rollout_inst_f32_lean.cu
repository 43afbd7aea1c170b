// fp32 production kernels, lean build: no cone hits, no noise, no trajectory dump, no audit channel.
#include "bgr_rollout_f32.cuh"
namespace bgr {
BGR_FOR_ALL_LAWS(BGR_DEFINE_LAUNCH_F32, float, false)
}
