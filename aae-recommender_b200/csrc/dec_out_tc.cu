// K3/K5 tensor-core variant (tcgen05 + TMEM) -- placeholder until the kernel lands.
#include "common.cuh"
namespace aae {
int dec_out_train_tc(const float*, int, int, float*, float*, float*, float*, float*, float*, int, int, const int32_t*,
                     const int32_t*, double, const aae_step_state*, float*, double*, int, cudaStream_t) {
  set_error("dec_out_train: tensor-core kernel not built");
  return AAE_E_UNSUPPORTED;
}
int dec_out_scores_tc(const float*, int, int, const float*, const float*, int, int, float*, int64_t, int,
                      cudaStream_t) {
  set_error("dec_out_scores: tensor-core kernel not built");
  return AAE_E_UNSUPPORTED;
}
}  // namespace aae
