// TEST INFRASTRUCTURE ONLY -- extern "C" shim around the UNMODIFIED reference grid subsampling
// (/root/reference/SSDR_AL_s3dis/utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106).
// It performs the same vector marshalling as the reference's CPython wrapper (wrapper.cpp:202-221), which
// itself no longer compiles against numpy 2.x. Reference sources are compiled where they lie (oracle/Makefile).
#include <cstring>
#include <vector>
#include "grid_subsampling/grid_subsampling.h"

struct RefGridResult {
    std::vector<PointXYZ> pts;
    std::vector<float> feats;
    std::vector<int> cls;
};

extern "C" {
void* ref_grid_run(const float* pts, const float* feats, const int* cls, size_t N, size_t fdim, size_t ldim,
                   float dl, size_t* M) {
    std::vector<PointXYZ> op((const PointXYZ*)pts, (const PointXYZ*)pts + N);
    std::vector<float> of;
    std::vector<int> oc;
    if (feats) of.assign(feats, feats + N * fdim);
    if (cls) oc.assign(cls, cls + N * ldim);
    RefGridResult* r = new RefGridResult();
    grid_subsampling(op, r->pts, of, r->feats, oc, r->cls, dl, 0);
    *M = r->pts.size();
    return r;
}
void ref_grid_fetch(void* h, float* pts, float* feats, int* cls) {
    RefGridResult* r = (RefGridResult*)h;
    if (pts) std::memcpy(pts, r->pts.data(), r->pts.size() * sizeof(PointXYZ));
    if (feats && !r->feats.empty()) std::memcpy(feats, r->feats.data(), r->feats.size() * sizeof(float));
    if (cls && !r->cls.empty()) std::memcpy(cls, r->cls.data(), r->cls.size() * sizeof(int));
}
void ref_grid_free(void* h) { delete (RefGridResult*)h; }
}
