// nccl_shim.cu -- NCCL reached through dlopen so that single-GPU use has no NCCL dependency
// (placeholder: multi-GPU selection lands with the sharding milestone).
#include "common.cuh"
using namespace ssdr;
extern "C" {
int ssdr_fps_f32_sharded(const float*, size_t, size_t, size_t, size_t, int32_t, size_t, int32_t*, void*, void*) { return set_error(SSDR_ERR_UNSUPPORTED, "sharded FPS not built yet"); }
int ssdr_nccl_unique_id(void*) { return set_error(SSDR_ERR_UNSUPPORTED, "nccl shim not built yet"); }
int ssdr_nccl_comm_init(void**, int, const void*, int) { return set_error(SSDR_ERR_UNSUPPORTED, "nccl shim not built yet"); }
int ssdr_nccl_comm_destroy(void*) { return set_error(SSDR_ERR_UNSUPPORTED, "nccl shim not built yet"); }
}
