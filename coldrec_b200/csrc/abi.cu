// Library-level entry points: error strings, version, device check.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace cr {

static thread_local char g_last_error[512] = "";

int note_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    return CR_ERR_CUDA;
}

int require_device() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_rc = CR_ERR_NO_DEVICE;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return CR_ERR_NO_DEVICE;
    }
    if (dev == cached_dev) return cached_rc;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return CR_ERR_NO_DEVICE;
    }
    cached_dev = dev;
    cached_rc = (major == 10) ? CR_OK : CR_ERR_NO_DEVICE;
    return cached_rc;
}

}  // namespace cr

extern "C" {

const char* cr_strerror(int code) {
    switch (code) {
        case CR_OK: return "ok";
        case CR_ERR_ARG: return "invalid argument (null pointer, negative size or inconsistent shape)";
        case CR_ERR_ALIGN: return "pointer or leading dimension is not 16-byte aligned";
        case CR_ERR_UNSUPPORTED: return "parameter outside the supported set (d % 4, d <= 512, K <= 64)";
        case CR_ERR_WORKSPACE: return "workspace too small";
        case CR_ERR_CUDA: return "CUDA call failed (see cr_last_cuda_error)";
        case CR_ERR_NO_DEVICE: return "no sm_100 CUDA device available: coldrec_b200 has no CPU fallback";
        default: return "unknown error code";
    }
}

const char* cr_last_cuda_error(void) { return cr::g_last_error; }

int cr_version(void) { return 100; }

int cr_device_check(void) { return cr::require_device(); }

}  // extern "C"
