// Tensor-core (tcgen05/TMEM/TMA) convolution - placeholder until the kernel lands: reports that no
// layer is supported so every conv runs on the CUDA-core path.
#include "vt_hift.cuh"
namespace vt {
bool conv_tc_supported(const ConvLayer&) { return false; }
int conv_tc_tile_rows(const ConvLayer&) { return 128; }
int pack_conv_tc(ConvLayer&, const std::vector<float>&, int, std::vector<void*>&) { return VT_OK; }
int launch_conv_tc(const ConvArgs&, const ConvLayer&, int, const void*, int, int, cudaStream_t) {
  set_error("tensor-core conv not built");
  return VT_ERR_UNSUPPORTED;
}
}  // namespace vt
