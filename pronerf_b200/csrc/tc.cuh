// fp16 tcgen05 tier of the three MLPs (mlp_tc.cu).
#pragma once
#include "common.cuh"

namespace pn {

struct NetTC {
  int n_layers = 0;
  int in_dim[kMaxLayers] = {0};
  int out_dim[kMaxLayers] = {0};
  int net_id = -1;
  void* blob = nullptr;          // device: packed fp16 weight images + fp32 biases (layout in mlp_tc.cu)
  size_t blob_bytes = 0;
  int* error_flag = nullptr;     // device: set by the kernel's barrier watchdog before it traps
  float* dirterm = nullptr;      // device scratch (NeRF): per-ray view-direction term of the last layer, grown on demand
  size_t dirterm_rows = 0;
  bool classic = false;          // classic NeRF topology (tc_load_nerf_classic)
  float alpha_bias = 0.f;        // classic NeRF: alpha_linear.bias
  bool supported = false;        // shape within the tensor-core kernel's limits
  bool loaded = false;
};

void tc_free_net(NetTC& n);
int tc_load_net(NetTC& n, int net_id, int n_layers, const int* in_dims, const int* out_dims, const float* const* W,
                const float* const* b, cudaStream_t stream);
bool tc_available();
void tc_set_timeline(long long* dev_buf);   // debug: clock64 stamps of CTA 0's second tile (208 slots)
void tc_set_clock(long long* dev_buf);      // measurement aid: per network {clock64, globaltimer} at the start and end of CTA 0 (12 slots)
int tc_launch_mlp(NetTC& n, const MlpLaunch& L, cudaStream_t stream);
// device pointer to the 4 x 27 fp32 view-direction weights of a loaded DoNeRFTRT (NULL: not loaded / classic topology)
const float* tc_wdir(const NetTC& n);
// classic NeRF (helpers.py:792-847) on the tensor-core tier: 12 tensors in checkpoint order; run_network form only
int tc_load_nerf_classic(NetTC& n, const int* in_dims, const int* out_dims, const float* const* W, const float* const* b,
                         cudaStream_t stream);
int tc_launch_nerf_classic(NetTC& n, const float* pts, const float* viewdirs, int viewdir_stride, int S, int64_t M, float* raw,
                           cudaStream_t stream);

}  // namespace pn
