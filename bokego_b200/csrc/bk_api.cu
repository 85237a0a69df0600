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

// The binary holds sm_100a code only (arch-specific: it does not run on sm_103 or any other 10.x part).
extern "C" int bk_device_check(void)
{
    int dev = 0, major = 0, minor = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -2;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -2;
    return (major == 10 && minor == 0) ? 0 : -2;
}

// Index of the current device for the per-device launch state the .cu files keep (kernel attributes, SM count, tensor
// maps): -1 when there is no current device or it is beyond BK_MAX_DEVICES.
int bk_current_device_slot(void)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= BK_MAX_DEVICES) return -1;
    return dev;
}

extern "C" size_t bk_feats_conv_bytes(int B)
{
    if (B <= 0) return 0;
    return (size_t)((B + BK_GROUP - 1) / BK_GROUP) * BK_F_GROUP_BYTES;
}

extern "C" size_t bk_weights_blob_bytes(void) { return BK_W_BLOB_BYTES; }
