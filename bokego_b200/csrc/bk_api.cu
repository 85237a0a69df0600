// bk_api.cu -- the small host-only entry points of the C ABI (include/bokego_b200.h).
#include <cuda_runtime.h>
#include <stddef.h>

#include "bk_layout.h"

extern "C" int bk_version(void) { return 100; }

extern "C" const char *bk_strerror(int code)
{
    switch (code) {
        case 0: return "ok";
        case -1: return "bad argument";
        case -2: return "device is not sm_100 (B200); this library has no other code path";
        case -3: return "CUDA launch failed";
        default: return "unknown error";
    }
}

extern "C" int bk_device_check(void)
{
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -2;
    return major == 10 ? 0 : -2;
}

extern "C" size_t bk_feats_conv_bytes(int B)
{
    if (B <= 0) return 0;
    return (size_t)((B + BK_GROUP - 1) / BK_GROUP) * BK_F_GROUP_BYTES;
}

extern "C" size_t bk_weights_blob_bytes(void) { return BK_W_BLOB_BYTES; }
