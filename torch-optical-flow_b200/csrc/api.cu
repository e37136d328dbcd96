// api.cu -- version / error strings / launch counter of libofb200.so.
#include "common.cuh"

int64_t g_ofb_launches = 0;

OFB_API int ofb_version(void) { return OFB_VERSION; }

OFB_API int64_t ofb_launch_count(void) { return __atomic_load_n(&g_ofb_launches, __ATOMIC_RELAXED); }

OFB_API const char* ofb_strerror(int code) {
    switch (code) {
        case OFB_OK: return "ok";
        case OFB_EINVAL: return "invalid argument (null pointer, negative size or unknown enum value)";
        case OFB_EUNSUPPORTED: return "request not supported by this kernel build";
        case OFB_EALIGN: return "pointer, pitch or stride not aligned as the kernel requires";
        case OFB_EDRIVER: return "cuTensorMapEncodeTiled unavailable or rejected the tensor map";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown ofb200 error";
}
