// bf16 tcgen05 / TMEM tier of the three MLPs -- placeholder until the kernel lands.
#include "tc.cuh"

namespace pn {

void tc_free_net(NetTC& n) {
  if (n.blob) cudaFree(n.blob);
  n = NetTC();
}

int tc_load_net(NetTC& n, int, int n_layers, const int* in_dims, const int* out_dims, const float* const*,
                const float* const*, cudaStream_t) {
  n.n_layers = n_layers;
  for (int l = 0; l < n_layers; ++l) { n.in_dim[l] = in_dims[l]; n.out_dim[l] = out_dims[l]; }
  return PN_OK;
}

bool tc_available() { return false; }

int tc_launch_mlp(const NetTC&, const MlpLaunch&, cudaStream_t) {
  set_error("PN_PREC_BF16: the tcgen05 MLP kernel is not built in this revision");
  return PN_ESTATE;
}

}  // namespace pn
