// ABI bookkeeping for libkdpc.
#include "common.cuh"

KDPC_API int kdpc_abi_version(void) { return KDPC_ABI_VERSION; }

KDPC_API const char *kdpc_error_string(int code) {
    switch (code) {
        case KDPC_OK: return "ok";
        case KDPC_EINVAL: return "kdpc: invalid argument (null pointer, non-positive size or misaligned buffer)";
        case KDPC_EUNSUPPORTED: return "kdpc: size outside the instantiated kernel range";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "kdpc: unknown error";
}

static int g_sm_limit = 0;
/* 0 = all SMs (default).  n > 0: persistent kernels and split plans size themselves for at most n SMs. */
KDPC_API void kdpc_set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }
KDPC_API int kdpc_sm_limit(void) { return g_sm_limit; }
