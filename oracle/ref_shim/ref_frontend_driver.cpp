// ref_frontend_driver.cpp — TEST INFRASTRUCTURE.  A "mini front-end" made of the reference's REAL classes compiled together,
// unmodified: FrameKTL (src/FrameKTL.cc), MapPoint (src/MapPoint.cc), ORBmatcher (src/ORBmatcher.cc), ORBextractor
// (src/ORBextractor.cc) — only KeyFrame / Map / IMU / Converter are stand-ins (frame_deps_standin.h).  It replays what
// Tracking::SearchLocalPoints does for one frame (src/Tracking.cc:2196-2228): build the frame's keypoint grid
// (FrameKTL::compute_descriptors :250-264 with PosInGrid :426-436), run FrameKTL::isInFrustum (:299-357, with the real
// MapPoint::PredictScale) on every map point, then ORBmatcher::SearchByProjection(F, vpMapPoints, th) (:49-125, which walks
// the real grid through the real FrameKTL::GetFeaturesInArea :359-424).  This pins the one piece the other _ref libraries had
// to restate: the frame grid.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "ORBmatcher.h"
#include "ORBextractor.h"

using namespace USLAM;

std::set<MapPoint*> KeyFrame::GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mapPoints) if (p && !p->isBad()) s.insert(p); return s; }

extern "C" {

// frame: nk keypoints (x, y, octave, angle) + descriptors; bounds = {minX, maxX, minY, maxY}; intr = {fx, fy, cx, cy};
// Tcw 4x4 row-major.  map points: world position, the level and camera centre (ref_Ow) of their single observation, descriptor.
// out: inview / u / v / level / viewcos per map point (what isInFrustum left in the MapPoint), owner[nk] = claiming map point
// or -1, the real grid as CSR (cell = ix * 48 + iy), and the return value of SearchByProjection.
int reff_search_local_points(int nk, const float* kxyoa, const uint8_t* kdesc, const int32_t* bounds, const float* intr, int nlevels,
                             float scale_factor, const float* Tcw, int np, const float* pos, const int32_t* obs_level, const float* ref_Ow,
                             const uint8_t* pdesc, float th, float nnratio, float cos_limit,
                             int32_t* inview, float* u, float* v, int32_t* level, float* viewcos, int32_t* owner,
                             int32_t* cell_start, int32_t* cell_items)
{
    ORBextractor extractor(1000, scale_factor, nlevels, ORBextractor::FAST_SCORE, 20);
    FrameKTL F;
    F.mpORBextractor = &extractor; F.mpORBvocabulary = 0;
    FrameKTL::mnMinX = bounds[0]; FrameKTL::mnMaxX = bounds[1]; FrameKTL::mnMinY = bounds[2]; FrameKTL::mnMaxY = bounds[3];
    FrameKTL::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(bounds[1] - bounds[0]);     // src/FrameKTL.cc:85-86
    FrameKTL::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(bounds[3] - bounds[2]);
    FrameKTL::fx = intr[0]; FrameKTL::fy = intr[1]; FrameKTL::cx = intr[2]; FrameKTL::cy = intr[3];
    F.mvKeysUn.resize((size_t)nk);
    for (int i = 0; i < nk; i++) F.mvKeysUn[(size_t)i] = cv::KeyPoint(kxyoa[4 * i], kxyoa[4 * i + 1], 31.f, kxyoa[4 * i + 3], 0.f, (int)kxyoa[4 * i + 2], -1);
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors = cv::Mat(nk > 0 ? nk : 1, 32, CV_8UC1);
    if (nk > 0) memcpy(F.mDescriptors.data, kdesc, (size_t)nk * 32);
    F.mvpMapPoints.assign((size_t)nk, (MapPoint*)0);
    F.SetN(nk);
    F.compute_descriptors();                       // scale tables + the keypoint grid
    F.mfLogScaleFactor = log(F.mfScaleFactor);     // src/FrameKTL.cc:97 (image constructor)
    cv::Mat T(4, 4, CV_32F); memcpy(T.data, Tcw, 64);
    F.SetPose(T);
    // the real grid, cell by cell
    int n_items = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
        for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            cell_start[ix * FRAME_GRID_ROWS + iy] = n_items;
            for (size_t k = 0; k < F.mGrid[ix][iy].size(); k++) cell_items[n_items++] = (int32_t)F.mGrid[ix][iy][k];
        }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = n_items;
    // map points: one observation each in a stand-in keyframe
    Map map;
    KeyFrame kf;
    kf.mnScaleLevels = nlevels; kf.mfLogScaleFactor = F.mfLogScaleFactor; kf.mvScaleFactors = F.mvScaleFactors;
    kf.keysUn.resize((size_t)np); kf.descriptors = cv::Mat(np > 0 ? np : 1, 32, CV_8UC1);
    if (np > 0) memcpy(kf.descriptors.data, pdesc, (size_t)np * 32);
    kf.Ow = cv::Mat(3, 1, CV_32F); memcpy(kf.Ow.data, ref_Ow, 12);
    std::vector<MapPoint*> mps((size_t)np);
    for (int i = 0; i < np; i++) {
        kf.keysUn[(size_t)i].octave = obs_level[i];
        cv::Mat P(3, 1, CV_32F); memcpy(P.data, pos + 3 * (size_t)i, 12);
        MapPoint* mp = new MapPoint(P, &kf, &map);
        mp->AddObservation(&kf, (size_t)i);
        mp->ComputeDistinctiveDescriptors();
        mp->UpdateNormalAndDepth();
        mps[(size_t)i] = mp;
    }
    for (int i = 0; i < np; i++) {                 // Tracking::SearchLocalPoints, src/Tracking.cc:2206-2215
        MapPoint* mp = mps[(size_t)i];
        const bool in = F.isInFrustum(mp, cos_limit);
        inview[i] = in ? 1 : 0; u[i] = mp->mTrackProjX; v[i] = mp->mTrackProjY; level[i] = mp->mnTrackScaleLevel; viewcos[i] = mp->mTrackViewCos;
    }
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(F, mps, th);
    for (int k = 0; k < nk; k++) {
        owner[k] = -1;
        if (F.mvpMapPoints[(size_t)k]) for (int i = 0; i < np; i++) if (mps[(size_t)i] == F.mvpMapPoints[(size_t)k]) { owner[k] = i; break; }
    }
    for (int i = 0; i < np; i++) delete mps[(size_t)i];
    return n;
}

}  // extern "C"
