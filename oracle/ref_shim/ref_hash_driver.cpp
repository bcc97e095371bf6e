// ref_hash_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own haloc::Hash (src/hash.cpp, compiled
// where it lies): getHash (:57-85) and match (:190-206).  The reference seeds its projection vectors with time(NULL)
// (:95); this library interposes time() with a constant so that a run is reproducible, lets the reference build its
// vectors, and hands them out so that the oracle and the CUDA kernel can be checked on the SAME vectors.
#include <stdint.h>
#include <string.h>
#include <time.h>
#include "uvip_cv_standin.hpp"  // everything hash.h pulls in, before the access override below
#include <Eigen/Dense>
#define private public          // r_ (the projection vectors) is private in include/hash.h
#include "hash.h"
#undef private

extern "C" {

time_t time(time_t* t) { if (t) *t = (time_t)20261017; return (time_t)20261017; }       // haloc::Hash::initProjections :95

void* refh_create(int num_proj)
{
    haloc::Hash* h = new haloc::Hash();
    haloc::Hash::Params p; p.num_proj = num_proj;
    h->setParams(p);
    return h;
}
void refh_destroy(void* h) { delete (haloc::Hash*)h; }
// getHash on `rows` x 32 descriptors; returns the hash length (num_proj * 32)
int refh_get_hash(void* h_, const uint8_t* desc, int rows, float* out)
{
    haloc::Hash* h = (haloc::Hash*)h_;
    cv::Mat d(rows > 0 ? rows : 1, 32, CV_8UC1);
    if (rows > 0) memcpy(d.data, desc, (size_t)rows * 32);
    if (rows == 0) d.rows = 0;
    const std::vector<float> v = h->getHash(d);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return (int)v.size();
}
int refh_projection_length(void* h_) { haloc::Hash* h = (haloc::Hash*)h_; return h->r_.empty() ? 0 : (int)h->r_[0].size(); }
void refh_get_projections(void* h_, float* out)
{
    haloc::Hash* h = (haloc::Hash*)h_;
    for (size_t i = 0; i < h->r_.size(); i++) memcpy(out + i * h->r_[0].size(), h->r_[i].data(), h->r_[i].size() * sizeof(float));
}
float refh_match(void* h_, const float* a, const float* b, int n)
{
    return ((haloc::Hash*)h_)->match(std::vector<float>(a, a + n), std::vector<float>(b, b + n));
}

}  // extern "C"
