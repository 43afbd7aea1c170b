// fp32 kernels with every runtime option: cone hits, Philox observation noise, trajectory dump,
// near-tie audit channel, single-step plumbing.  Same arithmetic as the lean build, bit for bit.
#include "bgr_rollout_f32.cuh"
namespace bgr {
BGR_FOR_ALL_LAWS(BGR_DEFINE_LAUNCH_F32, float, true)
}
