// ABI bookkeeping for libalignn_b200.so.
#include "common.cuh"

extern "C" int alignn_abi_version(void) { return ALIGNN_ABI_VERSION; }

extern "C" const char *alignn_error_string(int code) {
    switch (code) {
        case ALIGNN_OK: return "ok";
        case ALIGNN_ERR_BAD_ARG: return "bad argument (null pointer, negative size, misaligned buffer or invalid probability)";
        case ALIGNN_ERR_BAD_SHAPE: return "unsupported shape (hidden % heads != 0, hidden too large, or sizes beyond int32 indexing)";
        case ALIGNN_ERR_WORKSPACE: return "workspace too small (see alignn_plan_workspace_bytes)";
        case ALIGNN_ERR_BAD_DTYPE: return "dtype must be ALIGNN_F32 or ALIGNN_BF16";
        default: break;
    }
    if (code >= ALIGNN_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(code - ALIGNN_ERR_CUDA_BASE));
    return "unknown error";
}
