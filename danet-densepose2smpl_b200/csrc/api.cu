// Error plumbing and device queries for libdanet_b200.so.
#include "common.cuh"

namespace danet {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace danet

extern "C" const char* danet_last_error(void) { return danet::g_err; }
extern "C" int danet_version(void) { return 3; }   // 3: danet_act views, danet_conv_tc_group (split-fp16 tensor-core engine)
extern "C" int danet_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    DANET_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    DANET_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return 0;
}
