// grid_standin.h — TEST INFRASTRUCTURE.  The keypoint grid behind the stand-in GetFeaturesInArea of slam_standin.h /
// mappoint_deps_standin.h: forwards to the C oracle's restatement (src/FrameKTL.cc:250-264,359-436, src/KeyFrame.cc:952-992).
#ifndef UVIP_GRID_STANDIN_H
#define UVIP_GRID_STANDIN_H
#include <vector>
#include "uvip_cv_standin.hpp"
#ifndef FRAME_GRID_ROWS
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
#endif
namespace USLAM {
// CSR over cols*rows cells as the oracle builds it
struct GridStandin {
    std::vector<float> kx, ky; std::vector<int32_t> octave, cell_start, cell_items;
    float minX, minY, inv_w, inv_h; int cols, rows;
    void build(const std::vector<cv::KeyPoint>& k, float mnMinX, float mnMaxX, float mnMinY, float mnMaxY)
    {
        cols = FRAME_GRID_COLS; rows = FRAME_GRID_ROWS; minX = mnMinX; minY = mnMinY;
        inv_w = (float)cols / (mnMaxX - mnMinX); inv_h = (float)rows / (mnMaxY - mnMinY);       // src/FrameKTL.cc:150-151
        const int n = (int)k.size();
        kx.resize(n); ky.resize(n); octave.resize(n); cell_start.assign(cols * rows + 1, 0); cell_items.assign(n ? n : 1, 0);
        for (int i = 0; i < n; i++) { kx[i] = k[i].pt.x; ky[i] = k[i].pt.y; octave[i] = k[i].octave; }
        uo_grid_build(kx.data(), ky.data(), n, minX, minY, inv_w, inv_h, cols, rows, cell_start.data(), cell_items.data());
    }
    std::vector<size_t> area(float x, float y, float r, int minLevel, int maxLevel) const
    {
        std::vector<int32_t> out(kx.size() ? kx.size() : 1);
        const int n = uo_features_in_area(kx.data(), ky.data(), octave.data(), cell_start.data(), cell_items.data(), minX, minY, inv_w, inv_h,
                                          cols, rows, x, y, r, minLevel, maxLevel, out.data(), (int)out.size());
        return std::vector<size_t>(out.begin(), out.begin() + n);
    }
};

}  // namespace USLAM
#endif
