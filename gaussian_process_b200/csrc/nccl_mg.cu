// Multi-GPU pieces (one process per GPU).  NCCL is dlopen'ed at run time from the torch-bundled
// libnccl.so.2 so that libgpx.so itself has no link-time dependency on it.
#include <dlfcn.h>
#include "common.cuh"

extern "C" int gpx_nccl_unique_id(void* id128) {
    (void)id128;
    gpx_set_error("gpx_nccl_unique_id: not implemented yet");
    return GPX_E_NCCL;
}
extern "C" int gpx_nccl_init(gpx_handle h, const void* id128, int rank, int world) {
    (void)h; (void)id128; (void)rank; (void)world;
    gpx_set_error("gpx_nccl_init: not implemented yet");
    return GPX_E_NCCL;
}
extern "C" int gpx_potrf_mg(gpx_handle h, double* Aloc, int64_t n, int64_t ldl, int64_t nb, double* panel, double* dinv) {
    (void)h; (void)Aloc; (void)n; (void)ldl; (void)nb; (void)panel; (void)dinv;
    gpx_set_error("gpx_potrf_mg: not implemented yet");
    return GPX_E_NCCL;
}
