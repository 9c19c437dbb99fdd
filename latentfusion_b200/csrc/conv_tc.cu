// tcgen05 implicit-GEMM convolution (placeholder until the UMMA path lands; reports "unsupported").
#include "common.cuh"
namespace lf {
int conv_tc_supported(const lf_conv_desc*) { return 0; }
int conv_tc_launch(const lf_conv_desc*, const float*, const float*, const float*, float*, float*, cudaStream_t) {
    set_error("conv: tcgen05 path not built");
    return LF_EUNSUPPORTED;
}
}  // namespace lf
