// ref_mappoint_driver.cpp — TEST INFRASTRUCTURE.  The reference's REAL MapPoint (src/MapPoint.cc) over stand-in KeyFrame / Map:
// MapPoint::ComputeDistinctiveDescriptors (:197-270) on ragged observation lists.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "MapPoint.h"

using namespace USLAM;

extern "C" {

// desc: all observed descriptors, list p = rows start[p]..start[p+1]-1.  For every list the chosen descriptor (what the
// reference stores in mDescriptor) is written to out_desc + 32 p; chosen[p] = 1 when the list was not empty.
// The reference walks std::map<KeyFrame*, size_t>, i.e. observations ordered by keyframe ADDRESS; the keyframes of a list are
// laid out in one array here, so that order is the list order.
void refp_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int npoints, uint8_t* out_desc, int32_t* chosen)
{
    Map map;
    const float pos[3] = {0.f, 0.f, 1.f};
    for (int p = 0; p < npoints; p++) {
        const int n = start[p + 1] - start[p];
        chosen[p] = 0;
        std::vector<KeyFrame> kfs((size_t)(n > 0 ? n : 1));
        for (size_t i = 0; i < kfs.size(); i++) { kfs[i].mnId = (unsigned long)i; kfs[i].mnScaleLevels = 8; kfs[i].mvScaleFactors.assign(8, 1.f); }
        cv::Mat P(3, 1, CV_32F); memcpy(P.data, pos, 12);
        MapPoint mp(P, &kfs[0], &map);
        for (int i = 0; i < n; i++) {
            kfs[(size_t)i].descriptors = cv::Mat(1, 32, CV_8UC1);
            memcpy(kfs[(size_t)i].descriptors.data, desc + (size_t)(start[p] + i) * 32, 32);
            mp.AddObservation(&kfs[(size_t)i], 0);
        }
        mp.ComputeDistinctiveDescriptors();
        const cv::Mat d = mp.GetDescriptor();
        if (n > 0 && !d.empty()) { memcpy(out_desc + (size_t)p * 32, d.data, 32); chosen[p] = 1; }
    }
}

}  // extern "C"
