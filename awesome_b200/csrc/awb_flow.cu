// RealNVP coupling flow -- placeholder translation unit, filled in next.
#include "awb_internal.cuh"

namespace awb {
int flow_forward(const awb_prior*, const float*, const awb_grid_spec*, const Workspace&, float*, cudaStream_t) {
  set_error("flow prior kernels not built yet");
  return AWB_ERR_UNSUPPORTED;
}
int flow_backward(const awb_prior*, const float*, const awb_grid_spec*, const Workspace&, cudaStream_t) {
  set_error("flow prior kernels not built yet");
  return AWB_ERR_UNSUPPORTED;
}
int flow_actnorm_init(const awb_prior*, float*, const awb_grid_spec*, const Workspace&, cudaStream_t) {
  set_error("flow prior kernels not built yet");
  return AWB_ERR_UNSUPPORTED;
}
}  // namespace awb
