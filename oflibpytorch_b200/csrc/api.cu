// Library-level entry points of the C ABI (include/oflib_b200.h).
#include "common.cuh"

namespace ofb {
std::atomic<int64_t> g_launches{0};
}

extern "C" int ofb_abi_version(void) { return OFB_ABI_VERSION; }

extern "C" const char *ofb_error_string(int status)
{
    if (status == 0) return "success";
    if (status == OFB_EINVAL) return "oflib_b200: invalid argument";
    if (status == OFB_ESHAPE) return "oflib_b200: unsupported shape (H and W must be >= 2)";
    if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
    return "oflib_b200: unknown error";
}

extern "C" int64_t ofb_launch_count(void) { return ofb::g_launches.load(); }
extern "C" void ofb_launch_count_reset(void) { ofb::g_launches = 0; }
