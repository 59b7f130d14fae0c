// ilqg_kernels.cuh - iLQG sweeps on the device (finite-difference model derivatives, Gauss-Newton cost
// derivatives, Riccati backward pass).  Reference: mjpc/planners/model_derivatives.cc:45-165,
// mjpc/planners/cost_derivatives.cc:77-230, mjpc/planners/ilqg/backward_pass.cc:65-250.
#pragma once
#include <cuda_runtime.h>

#include "rollout_kernels.cuh"

namespace mjpc_dev {

struct IlqgBuffers {
  int H = 0;
  float* d_buf = nullptr;
};

inline int ilqg_init(IlqgBuffers& b, const DevModel& M, int H, size_t smem_debug) {
  (void)M; (void)smem_debug;
  b.H = H;
  return 0;
}
inline void ilqg_free(IlqgBuffers& b) {
  if (b.d_buf) cudaFree(b.d_buf);
  b.d_buf = nullptr;
}
inline int ilqg_model_derivatives(IlqgBuffers&, const DevModel&, const float*, cudaStream_t, const float*, const float*,
                                  const float*, const float*, const float*, int, float, float*, float*, float*, float*,
                                  size_t, int*) {
  return -5;
}
inline int ilqg_cost_derivatives(IlqgBuffers&, const DevModel&, const float*, cudaStream_t, const float*, const float*,
                                 const float*, int, float*, float*, float*, float*, float*, int*) {
  return -5;
}
inline int ilqg_backward_pass(IlqgBuffers&, const DevModel&, const float*, cudaStream_t, const float*, const float*,
                              const float*, const float*, const float*, const float*, const float*, const float*, int,
                              float, int, int, float*, float*, float*, float*, float*, int*, int*) {
  return -5;
}

}  // namespace mjpc_dev
