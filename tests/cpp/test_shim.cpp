// C++ host-shim test: drives USLAM::ORBextractor / USLAM::ORBmatcher (u-vip-slam_b200/host) exactly the way
// Tracking.cc does (src/Tracking.cc:893-964, :2177-2231) and dumps the results for the pytest harness, which
// compares them with the CPU oracle.  usage: test_shim <in.bin> <out.bin>
//   in.bin : int32 w, h, nfeatures, fastTh, min_px_dist, num_needed ; w*h bytes image
//   out.bin: [full detect] int32 n ; n*28 B keypoints ; n*32 B descriptors ; [grid path] same + grid ints ;
//            [matcher] int32 nmatches ; nq int32 (keypoint index per map point or -1)
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>
#include "../../u-vip-slam_b200/host/ORBextractor.h"
#include "../../u-vip-slam_b200/host/ORBmatcher.h"

struct MockMapPoint {
    bool mbTrackInView = true; int mnTrackScaleLevel = 0; float mTrackViewCos = 0.9f, mTrackProjX = 0, mTrackProjY = 0;
    cv::Mat desc; bool bad = false;
    float pos[3] = {0, 0, 1}, minDist = 1.0f;
    bool isBad() const { return bad; }
    cv::Mat GetDescriptor() const { return desc; }
    cv::Mat GetWorldPos() { return cv::Mat(3, 4, CV_8UC1, pos, 4); }              // 3x1 float viewed through the 8-bit stand-in (ptr<float>(r)[0])
    float GetMinDistanceInvariance() const { return minDist; }
};
typedef std::map<unsigned, std::vector<unsigned> > FeatureVector;     // DBoW2::FeatureVector
struct MockFrame {
    std::vector<cv::KeyPoint> mvKeysUn, mvKeys; cv::Mat mDescriptors; std::vector<MockMapPoint*> mvpMapPoints; std::vector<float> mvScaleFactors;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mnScaleLevels = 8; float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    FeatureVector mFeatVec;
    float pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; float fx = 1, fy = 1, cx = 0, cy = 0;
    cv::Mat mTcw;                                                                 // 4x4 float viewed through the 8-bit stand-in
    MockFrame() : mTcw(4, 16, CV_8UC1, pose, 16) {}
};
struct MockKeyFrame {
    std::vector<MockMapPoint*> mps; FeatureVector fv; cv::Mat desc; std::vector<cv::KeyPoint> keys;
    std::vector<MockMapPoint*> GetMapPointMatches() { return mps; }
    FeatureVector GetFeatureVector() { return fv; }
    cv::Mat GetDescriptor(size_t i) { return desc.row((int)i); }
    cv::KeyPoint GetKeyPointUn(size_t i) const { return keys[i]; }
    std::vector<cv::KeyPoint> GetKeyPointsUn() const { return keys; }
    cv::Mat GetDescriptors() const { return desc; }
};

static void put(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("write"); exit(2); } }

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: test_shim in.bin out.bin\n"); return 2; }
    FILE* fi = fopen(argv[1], "rb"); if (!fi) { perror("in"); return 2; }
    int hdr[6]; if (fread(hdr, 4, 6, fi) != 6) return 2;
    const int w = hdr[0], h = hdr[1], nfeatures = hdr[2], fastTh = hdr[3]; int min_px_dist = hdr[4]; const int num_needed = hdr[5];
    std::vector<unsigned char> pix((size_t)w * h);
    if (fread(pix.data(), 1, pix.size(), fi) != pix.size()) return 2;
    fclose(fi);
    FILE* fo = fopen(argv[2], "wb"); if (!fo) { perror("out"); return 2; }
    try {
        USLAM::ORBextractor ex(nfeatures, 1.2f, 8, USLAM::ORBextractor::FAST_SCORE, fastTh);
        cv::Mat img(h, w, CV_8UC1, pix.data(), (size_t)w);
        // 1. full detection, as in the NOT_INITIALIZED / LOST states (Tracking.cc:942-946)
        std::vector<cv::KeyPoint> kps; cv::Mat desc;
        Eigen::MatrixXi grid = Eigen::MatrixXi::Zero(h / min_px_dist + 2, w / min_px_dist + 2);
        ex(img, cv::Mat(), kps, desc, grid, min_px_dist, true, 0);
        int n = (int)kps.size();
        put(fo, &n, 4); put(fo, kps.data(), (size_t)n * 28);
        for (int i = 0; i < n; i++) put(fo, desc.ptr(i), 32);
        // 2. replenishing detection while WORKING: existing points mark the grid (Tracking.cc:901-909)
        std::vector<cv::KeyPoint> kps2(kps.begin(), kps.begin() + (n < 20 ? n : 20));
        for (auto& k : kps2) { k.pt.x = (float)(int)k.pt.x; k.pt.y = (float)(int)k.pt.y; k.octave = 0; }
        for (const auto& k : kps2) grid((int)(k.pt.y / min_px_dist), (int)(k.pt.x / min_px_dist))++;
        cv::Mat desc2;
        ex(img, cv::Mat(), kps2, desc2, grid, min_px_dist, false, num_needed);
        int n2 = (int)kps2.size();
        put(fo, &n2, 4); put(fo, kps2.data(), (size_t)n2 * 28);
        for (int i = 0; i < n2; i++) put(fo, desc2.ptr(i), 32);
        int gr = grid.rows(), gc = grid.cols();
        put(fo, &gr, 4); put(fo, &gc, 4); put(fo, grid.data(), (size_t)gr * gc * 4);
        // empty image: silent return, outputs untouched
        std::vector<cv::KeyPoint> keep(3); cv::Mat keepd;
        ex(cv::Mat(), cv::Mat(), keep, keepd, grid, min_px_dist, true, 0);
        int untouched = keep.size() == 3 ? 1 : 0; put(fo, &untouched, 4);
        // 3. SearchByProjection: project every detected keypoint back as a map point with a slightly wrong position
        MockFrame F;
        F.mvKeysUn = kps; F.mDescriptors = desc; F.mvpMapPoints.assign((size_t)n, nullptr);
        F.mvScaleFactors.resize(8); F.mvScaleFactors[0] = 1.0f;
        for (int i = 1; i < 8; i++) F.mvScaleFactors[i] = F.mvScaleFactors[i - 1] * ex.GetScaleFactor();
        F.mfGridElementWidthInv = 64.0f / (float)w; F.mfGridElementHeightInv = 48.0f / (float)h;
        std::vector<MockMapPoint> mps((size_t)n); std::vector<MockMapPoint*> vp;
        for (int i = 0; i < n; i++) {
            mps[i].mTrackProjX = kps[i].pt.x + ((i % 5) - 2) * 0.7f; mps[i].mTrackProjY = kps[i].pt.y + ((i % 3) - 1) * 0.9f;
            mps[i].mnTrackScaleLevel = kps[i].octave; mps[i].mTrackViewCos = (i & 1) ? 0.9990f : 0.9f;
            mps[i].mbTrackInView = (i % 11) != 0; mps[i].bad = (i % 13) == 0;
            mps[i].desc = desc.row(i);
            vp.push_back(&mps[i]);
        }
        USLAM::ORBmatcher matcher(0.8f);
        const int nm = matcher.SearchByProjection(F, vp, 1.0f);
        put(fo, &nm, 4);
        std::vector<int> owner((size_t)n, -1);
        for (int i = 0; i < n; i++) if (F.mvpMapPoints[i]) owner[i] = (int)(F.mvpMapPoints[i] - mps.data());
        put(fo, owner.data(), (size_t)n * 4);
        int dd = USLAM::ORBmatcher::DescriptorDistance(desc.row(0), desc.row(n > 1 ? 1 : 0));
        put(fo, &dd, 4);
        // 4. SearchByBoW(KeyFrame*, Frame&): the keyframe holds the same features; node id = (x / 64) + 16 * (y / 64) puts
        //    ~30 features into a node, a few map points are missing / bad
        MockKeyFrame KF; MockFrame F2;
        KF.desc = desc; KF.keys = kps; F2.mDescriptors = desc; F2.mvKeys = kps; F2.mvpMapPoints.assign((size_t)n, nullptr);
        std::vector<MockMapPoint> kfmps((size_t)n);
        for (int i = 0; i < n; i++) {
            const unsigned node = (unsigned)(kps[i].pt.x / 64) + 16u * (unsigned)(kps[i].pt.y / 64);
            KF.fv[node].push_back((unsigned)i);
            if (i % 17 != 3) F2.mFeatVec[node + (i % 29 == 0 ? 1000u : 0u)].push_back((unsigned)i);
            kfmps[i].bad = (i % 19) == 0;
            KF.mps.push_back((i % 7 == 0) ? nullptr : &kfmps[i]);
        }
        USLAM::ORBmatcher bow(0.9f, true);
        std::vector<MockMapPoint*> vm;
        const int nb = bow.SearchByBoW(&KF, F2, vm);
        put(fo, &nb, 4);
        std::vector<int> bowner((size_t)n, -1);
        for (int i = 0; i < n; i++) if (vm[i]) bowner[i] = (int)(vm[i] - kfmps.data());
        put(fo, bowner.data(), (size_t)n * 4);
        // 6 (written after 5 below). SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist): relocalisation
        // geometry.  The keyframe's map point i is keypoint i back-projected at depth 2 + (i % 7) * 0.5 through the camera,
        // then seen by a slightly rotated / translated current pose.
        MockFrame F3;
        F3.mvKeysUn = kps; F3.mvKeys = kps; F3.mDescriptors = desc; F3.mvpMapPoints.assign((size_t)n, nullptr);
        F3.mvScaleFactors = F.mvScaleFactors; F3.mfGridElementWidthInv = F.mfGridElementWidthInv; F3.mfGridElementHeightInv = F.mfGridElementHeightInv;
        F3.mnMaxX = w; F3.mnMaxY = h; F3.fx = 458.0f; F3.fy = 457.0f; F3.cx = 367.0f; F3.cy = 248.0f;
        {
            const float c = 0.99995f, sn = 0.0099998f;                            // ~0.573 degrees about the optical axis
            const float P[16] = {c, -sn, 0, 0.01f, sn, c, 0, -0.02f, 0, 0, 1, 0.03f, 0, 0, 0, 1};
            for (int i = 0; i < 16; i++) F3.pose[i] = P[i];
        }
        std::vector<MockMapPoint> mps3((size_t)n);
        MockKeyFrame KF3; KF3.desc = desc; KF3.keys = kps;
        std::set<MockMapPoint*> found;
        for (int i = 0; i < n; i++) {
            const float z = 2.0f + (float)(i % 7) * 0.5f;
            mps3[i].pos[0] = (kps[i].pt.x - F3.cx) / F3.fx * z; mps3[i].pos[1] = (kps[i].pt.y - F3.cy) / F3.fy * z; mps3[i].pos[2] = z;
            mps3[i].minDist = z / F3.mvScaleFactors[(size_t)kps[i].octave] * 1.05f;
            mps3[i].desc = desc.row(i); mps3[i].bad = (i % 23) == 0;
            KF3.mps.push_back((i % 9 == 0) ? nullptr : &mps3[i]);
            if (i % 10 == 1) found.insert(&mps3[i]);
            if (i % 31 == 5) F3.mvpMapPoints[i] = &mps3[0];                        // keypoints that already carry a map point
        }
        USLAM::ORBmatcher reloc(0.9f, true);
        const int nr = reloc.SearchByProjection(F3, &KF3, found, 10.0f, 100);
        std::vector<int> rowner((size_t)n, -1);
        for (int i = 0; i < n; i++) if (F3.mvpMapPoints[i]) rowner[i] = (int)(F3.mvpMapPoints[i] - mps3.data());
        // 5. MapPoint::ComputeDistinctiveDescriptors batched: list p = descriptor rows p, p+7, p+14, ... (1 + p % 9 of them)
        const int np = n < 64 ? n : 64;
        std::vector<std::vector<cv::Mat> > lists((size_t)np);
        for (int p = 0; p < np; p++)
            for (int j = 0; j < 1 + p % 9; j++) lists[(size_t)p].push_back(desc.row((p + 7 * j) % n));
        std::vector<int> best;
        bow.ComputeDistinctiveDescriptors(lists, best);
        put(fo, &np, 4);
        put(fo, best.data(), (size_t)np * 4);
        put(fo, &nr, 4);
        put(fo, rowner.data(), (size_t)n * 4);
        // 7. SearchByBoW(KeyFrame*, KeyFrame*): KF (section 4) against a second keyframe with the same features, other gaps
        MockKeyFrame KFb; KFb.desc = desc; KFb.keys = kps;
        std::vector<MockMapPoint> mpsb((size_t)n);
        for (int i = 0; i < n; i++) {
            const unsigned node = (unsigned)(kps[i].pt.x / 64) + 16u * (unsigned)(kps[i].pt.y / 64);
            if (i % 15 != 4) KFb.fv[node + (i % 37 == 0 ? 2000u : 0u)].push_back((unsigned)i);
            mpsb[i].bad = (i % 21) == 2;
            KFb.mps.push_back((i % 6 == 1) ? nullptr : &mpsb[i]);
        }
        std::vector<MockMapPoint*> m12;
        const int nkk = bow.SearchByBoW(&KF, &KFb, m12);
        put(fo, &nkk, 4);
        std::vector<int> o12((size_t)n, -1);
        for (int i = 0; i < n; i++) if (m12[i]) o12[i] = (int)(m12[i] - mpsb.data());
        put(fo, o12.data(), (size_t)n * 4);
    } catch (const std::exception& e) {
        fprintf(stderr, "shim error: %s\n", e.what());
        fclose(fo);
        return 3;
    }
    fclose(fo);
    return 0;
}
