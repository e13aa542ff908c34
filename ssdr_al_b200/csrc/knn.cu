// knn.cu -- k nearest neighbours (placeholder until the kernels land in the next milestone).
#include "common.cuh"
using namespace ssdr;
extern "C" {
int ssdr_knn(const float*, size_t, size_t, const float*, size_t, size_t, int64_t*) { return set_error(SSDR_ERR_UNSUPPORTED, "knn not built yet"); }
int ssdr_knn_batch(const float*, size_t, size_t, size_t, const float*, size_t, size_t, int64_t*) { return set_error(SSDR_ERR_UNSUPPORTED, "knn not built yet"); }
int ssdr_knn_batch_dev(const float*, size_t, size_t, const float*, size_t, size_t, int64_t*, void*, ssdr_knn_stats*) { return set_error(SSDR_ERR_UNSUPPORTED, "knn not built yet"); }
int ssdr_knn_batch_dev_i32(const float*, size_t, size_t, const float*, size_t, size_t, int32_t*, void*, ssdr_knn_stats*) { return set_error(SSDR_ERR_UNSUPPORTED, "knn not built yet"); }
}
