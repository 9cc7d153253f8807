// tcgen05 TF32 implicit-GEMM convolution (sm_100a).  Placeholder until the kernel lands: reports
// "unsupported" for every shape so callers route through the fp32 FMA path.
#include "common.cuh"

namespace danet {
int conv_tc_launch(const danet_conv_desc*, const float*, const void*, const float*, const float*, float*, cudaStream_t) {
    set_error("conv_tc_launch: tcgen05 path not built");
    return -1;
}
}  // namespace danet

extern "C" int danet_conv_tc_supported(const danet_conv_desc*) { return 0; }
extern "C" int64_t danet_conv_tc_packed_bytes(const danet_conv_desc*) { return 0; }
extern "C" int danet_conv_tc_pack(const danet_conv_desc*, const float*, void*, danet_stream_t) {
    danet::set_error("danet_conv_tc_pack: tcgen05 path not built");
    return -1;
}
