// ORBextractor.h — drop-in for the reference's include/ORBextractor.h (chintha/U-VIP-SLAM, :45-94): the same class
// in namespace USLAM, same constructor and operator() signature, same getters; the body forwards to the C-ABI of
// libuvip_orb.so (include/uvip_orb.h), i.e. to the sm_100a kernels.  Tracking.cc keeps calling
//     (*mpORBextractor)(img0pyr[0], cv::Mat(), pts0_ext, New_Descriptors, grid_2d, min_px_dist, FullDetect, num_featsneeded);
// (src/Tracking.cc:946) unchanged.  Build with -DUVIP_WITH_OPENCV against real OpenCV/Eigen, or without it against
// the stand-in types of uvip_compat.h (used by this repository's own tests, where OpenCV C++ is not installed).
//
// Error behaviour mirrors the reference: empty image -> silent return, outputs untouched (src/ORBextractor.cc:852-853);
// wrong type -> assert (:856); any C-ABI failure is thrown as std::runtime_error where the reference would have let a
// cv::Exception propagate.  There is no CPU fallback.
#pragma once
#include <cassert>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef UVIP_WITH_OPENCV
#include <opencv2/core/core.hpp>
#include <Eigen/Core>
#else
#include "uvip_compat.h"
#endif
#include "../../include/uvip_orb.h"

namespace USLAM {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures = 1000, float scaleFactor = 1.2f, int nlevels = 8, int scoreType = HARRIS_SCORE, int fastTh = 7)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), scoreType(scoreType), fastTh(fastTh) {}
    ~ORBextractor() { if (handle_) uvip_extractor_destroy(handle_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (include/ORBextractor.h:56-58)
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors,
                    Eigen::MatrixXi& grid_2d, int& min_px_dist, bool FullDetect, int num_featsneeded)
    {
        (void)mask;                                            // built but never consulted on the live path
        const cv::Mat img = image.getMat();
        if (img.empty()) return;
        assert(img.type() == CV_8UC1);
        ensure(img.cols, img.rows);
        const int n_in = FullDetect ? 0 : (int)keypoints.size();
        // incoming keypoints counted in steps of 256: cap is part of the call shape the library keys its captured CUDA graph on
        const int cap = nfeatures + 8 * nlevels + 64 + (n_in + 255) / 256 * 256;
        std::vector<uvip_keypoint> kp((size_t)cap);
        static_assert(sizeof(uvip_keypoint) == sizeof(cv::KeyPoint), "cv::KeyPoint layout");
        if (n_in) std::memcpy(kp.data(), keypoints.data(), sizeof(uvip_keypoint) * (size_t)n_in);
        std::vector<unsigned char> desc((size_t)cap * 32);
        int n = n_in;
        const int rc = uvip_extract(handle_, img.data, img.cols, img.rows, (int)img.step, kp.data(), &n, cap, desc.data(),
                                    grid_2d.data(), (int)grid_2d.rows(), (int)grid_2d.cols(), min_px_dist, FullDetect ? 1 : 0, num_featsneeded);
        if (rc != UVIP_OK) throw std::runtime_error(std::string("uvip_extract: ") + uvip_last_error());
        keypoints.resize((size_t)n);                           // cleared and replaced (:928-929,959)
        if (n) std::memcpy(static_cast<void*>(keypoints.data()), kp.data(), sizeof(uvip_keypoint) * (size_t)n);
        if (n == 0) descriptors.release();                     // :920-921
        else {
            descriptors.create(n, 32, CV_8U);
            cv::Mat d = descriptors.getMat();
            for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), desc.data() + (size_t)i * 32, 32);
        }
    }
    // upstream 4-argument convenience form (FullDetect = true)
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors)
    {
        Eigen::MatrixXi none(1, 1); int d = 1;
        (*this)(image, mask, keypoints, descriptors, none, d, true, 0);
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }

protected:
    void ensure(int w, int h)
    {
        if (handle_ && w <= max_w_ && h <= max_h_) return;
        if (handle_) { uvip_extractor_destroy(handle_); handle_ = nullptr; }
        uvip_extractor_params p;
        p.nfeatures = nfeatures; p.scale_factor = (float)scaleFactor; p.nlevels = nlevels; p.score_type = scoreType; p.fast_th = fastTh;
        p.retry_th = 0; p.cell = 0; p.device = 0; p.max_width = w; p.max_height = h; p.max_batch = 1;
        if (uvip_extractor_create(&p, &handle_) != UVIP_OK) throw std::runtime_error(std::string("uvip_extractor_create: ") + uvip_last_error());
        max_w_ = w; max_h_ = h;
    }

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int scoreType;
    int fastTh;
    uvip_extractor* handle_ = nullptr;
    int max_w_ = 0, max_h_ = 0;
};

}  // namespace USLAM
