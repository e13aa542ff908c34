// TEST INFRASTRUCTURE ONLY -- extern "C" forwarding shim around the UNMODIFIED reference KNN
// (/root/reference/SSDR_AL_s3dis/utils/nearest_neighbors/knn_.cxx:22-135, declared in knn_.h:2-17).
// The reference sources are compiled where they lie (see oracle/Makefile); nothing is copied.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the result.
#include <cstddef>
#include "knn_.h"

extern "C" {
void ref_knn(const float* pts, size_t npts, size_t dim, const float* q, size_t nq, size_t K, long* out, int omp) {
    if (omp) cpp_knn_omp(pts, npts, dim, q, nq, K, out);
    else     cpp_knn(pts, npts, dim, q, nq, K, out);
}
void ref_knn_batch(const float* pts, size_t B, size_t npts, size_t dim, const float* q, size_t nq, size_t K,
                   long* out, int omp) {
    if (omp) cpp_knn_batch_omp(pts, B, npts, dim, q, nq, K, out);
    else     cpp_knn_batch(pts, B, npts, dim, q, nq, K, out);
}
}
