// capi.cu -- version / error strings of the C ABI (include/satb200.h).
#include <cuda_runtime.h>
#include "../../include/satb200.h"

extern "C" {

int sat_abi_version(void) { return SATB200_ABI_VERSION; }

const char* sat_strerror(int code) {
    switch (code) {
        case SAT_OK: return "ok";
        case SAT_ERR_NULL: return "required pointer is NULL";
        case SAT_ERR_SIZE: return "bad size, stride or alignment";
        case SAT_ERR_MODE: return "bad mode / enum value";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

}  // extern "C"
