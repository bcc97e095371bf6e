// Host-thread concurrency test of the drop-in shim (the matcher is called from the Tracking, LocalMapping and
// LoopClosing threads — src/Tracking.cc:2222, src/LocalMapping.cc:1230, src/LoopClosing.cc:373 — each through a STACK-LOCAL
// ORBmatcher, while Tracking also runs the extractor).  Three matcher threads + one extractor thread run ITER iterations
// concurrently; every iteration's result must equal the serial run made first.  Also measures what the reference's call shape
// costs through the shim: construct + SearchByProjection + destruct at the fork's real sizes (400-1000 keypoints, a few
// hundred map points), printed as microseconds per call.
//   usage: test_threads [iterations]      exit 0 = all equal, 3 = no CUDA device (no CPU fallback), 1 = mismatch
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "../../u-vip-slam_b200/host/ORBextractor.h"
#include "../../u-vip-slam_b200/host/ORBmatcher.h"

struct MockMapPoint {
    bool mbTrackInView = true; int mnTrackScaleLevel = 0; float mTrackViewCos = 0.9f, mTrackProjX = 0, mTrackProjY = 0;
    std::vector<unsigned char> d; bool bad = false;
    bool isBad() const { return bad; }
    cv::Mat GetDescriptor() const { return cv::Mat(1, 32, CV_8UC1, const_cast<unsigned char*>(d.data()), 32); }
};
struct MockFrame {
    std::vector<cv::KeyPoint> mvKeysUn; cv::Mat mDescriptors; std::vector<MockMapPoint*> mvpMapPoints; std::vector<float> mvScaleFactors;
    int mnMinX = 0, mnMinY = 0; float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
};

static unsigned long long rng_state(unsigned long long& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return s >> 33; }

struct Scene {
    MockFrame F; std::vector<MockMapPoint> pts; std::vector<MockMapPoint*> vp;
    void build(unsigned long long seed, int nk, int np, int W, int H)
    {
        unsigned long long s = seed;
        F.mvScaleFactors.resize(8); F.mvScaleFactors[0] = 1.f; for (int l = 1; l < 8; l++) F.mvScaleFactors[l] = F.mvScaleFactors[l - 1] * 1.2f;
        F.mnMinX = 0; F.mnMinY = 0; F.mfGridElementWidthInv = 64.f / W; F.mfGridElementHeightInv = 48.f / H;
        F.mvKeysUn.resize((size_t)nk); F.mDescriptors.create(nk, 32, CV_8U); F.mvpMapPoints.assign((size_t)nk, (MockMapPoint*)0);
        for (int i = 0; i < nk; i++) {
            cv::KeyPoint& k = F.mvKeysUn[(size_t)i];
            k.pt.x = 16.f + (float)(rng_state(s) % (unsigned)(W - 32)); k.pt.y = 16.f + (float)(rng_state(s) % (unsigned)(H - 32));
            k.octave = (int)(rng_state(s) % 4); k.angle = (float)(rng_state(s) % 360);
            for (int b = 0; b < 32; b++) F.mDescriptors.ptr(i)[b] = (unsigned char)rng_state(s);
        }
        pts.resize((size_t)np); vp.resize((size_t)np);
        for (int p = 0; p < np; p++) {
            MockMapPoint& m = pts[(size_t)p];
            const int k = (int)(rng_state(s) % (unsigned)nk);
            m.mTrackProjX = F.mvKeysUn[(size_t)k].pt.x + (float)((int)(rng_state(s) % 5) - 2);
            m.mTrackProjY = F.mvKeysUn[(size_t)k].pt.y + (float)((int)(rng_state(s) % 5) - 2);
            m.mnTrackScaleLevel = F.mvKeysUn[(size_t)k].octave + (int)(rng_state(s) % 2);
            m.mTrackViewCos = (p & 1) ? 0.9990f : 0.9f;
            m.d.assign(F.mDescriptors.ptr(k), F.mDescriptors.ptr(k) + 32);
            for (int fl = 0; fl < (int)(rng_state(s) % 20); fl++) m.d[rng_state(s) % 32] ^= (unsigned char)(1u << (rng_state(s) % 8));
            m.bad = (p % 37) == 5; m.mbTrackInView = (p % 11) != 3;
            vp[(size_t)p] = &m;
        }
    }
    // one call exactly as Tracking::SearchLocalPoints makes it (src/Tracking.cc:2222-2231): stack-local matcher, th by state
    std::vector<int> run(float th, int* nmatches)
    {
        for (auto& p : F.mvpMapPoints) p = 0;
        USLAM::ORBmatcher matcher(0.8f);
        *nmatches = matcher.SearchByProjection(F, vp, th);
        std::vector<int> owner(F.mvpMapPoints.size(), -1);
        for (size_t k = 0; k < owner.size(); k++) if (F.mvpMapPoints[k]) owner[k] = (int)(F.mvpMapPoints[k] - &pts[0]);
        return owner;
    }
};

static void synth_image(std::vector<unsigned char>& img, int W, int H, unsigned long long seed)
{
    unsigned long long s = seed;
    img.assign((size_t)W * H, 0);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) img[(size_t)y * W + x] = (unsigned char)(96 + ((x / 16 + y / 16) & 1) * 24 + (int)(rng_state(s) % 7));
    for (int r = 0; r < 300; r++) {
        const int x0 = (int)(rng_state(s) % (unsigned)W), y0 = (int)(rng_state(s) % (unsigned)H), w = 6 + (int)(rng_state(s) % 43), h = 6 + (int)(rng_state(s) % 43);
        const unsigned char g = (unsigned char)rng_state(s);
        for (int y = y0; y < y0 + h && y < H; y++) for (int x = x0; x < x0 + w && x < W; x++) img[(size_t)y * W + x] = g;
    }
}

int main(int argc, char** argv)
{
    const int ITER = argc > 1 ? atoi(argv[1]) : 40;
    if (uvip_device_count() == 0) { fprintf(stderr, "no CUDA device available; libuvip_orb has no CPU fallback\n"); return 3; }
    const int W = 752, H = 480;
    try {
        // three matcher scenes at the fork's real sizes (Tracking / LocalMapping / LoopClosing), one extractor image
        Scene sc[3];
        sc[0].build(11, 1000, 600, W, H); sc[1].build(22, 700, 300, W, H); sc[2].build(33, 400, 900, W, H);
        const float ths[3] = {1.f, 3.f, 5.f};
        std::vector<unsigned char> pix; synth_image(pix, W, H, 7);
        cv::Mat img(H, W, CV_8UC1, pix.data(), (size_t)W);
        USLAM::ORBextractor ex(1000, 1.2f, 8, USLAM::ORBextractor::FAST_SCORE, 20);
        int mpd = 20;
        // serial run: the expected results
        std::vector<int> want[3]; int wantn[3];
        for (int t = 0; t < 3; t++) { want[t] = sc[t].run(ths[t], &wantn[t]); if (wantn[t] < 50) { fprintf(stderr, "scene %d matched only %d\n", t, wantn[t]); return 1; } }
        std::vector<cv::KeyPoint> wkps; cv::Mat wdesc; Eigen::MatrixXi g0 = Eigen::MatrixXi::Zero(H / mpd + 2, W / mpd + 2);
        ex(img, cv::Mat(), wkps, wdesc, g0, mpd, true, 0);
        if (wkps.size() < 500) { fprintf(stderr, "extractor found only %zu keypoints\n", wkps.size()); return 1; }
        // concurrent run
        std::atomic<int> bad(0);
        std::vector<std::thread> th;
        for (int t = 0; t < 3; t++)
            th.emplace_back([&, t]() {
                try {
                    for (int it = 0; it < ITER; it++) {
                        int n = 0; const std::vector<int> got = sc[t].run(ths[t], &n);
                        if (n != wantn[t] || got != want[t]) bad++;
                    }
                } catch (const std::exception& e) { fprintf(stderr, "matcher thread %d: %s\n", t, e.what()); bad += 1000; }
            });
        th.emplace_back([&]() {
            try {
                for (int it = 0; it < ITER; it++) {
                    std::vector<cv::KeyPoint> kps; cv::Mat desc; Eigen::MatrixXi g = Eigen::MatrixXi::Zero(H / mpd + 2, W / mpd + 2);
                    ex(img, cv::Mat(), kps, desc, g, mpd, true, 0);
                    bool same = kps.size() == wkps.size();
                    for (size_t i = 0; same && i < kps.size(); i++)
                        same = std::memcmp(&kps[i], &wkps[i], sizeof(cv::KeyPoint)) == 0 && std::memcmp(desc.ptr((int)i), wdesc.ptr((int)i), 32) == 0;
                    if (!same) bad++;
                }
            } catch (const std::exception& e) { fprintf(stderr, "extractor thread: %s\n", e.what()); bad += 1000; }
        });
        for (auto& t : th) t.join();
        // the reference's call shape, single thread: construct + search + destruct per call (handle is thread-local: no CUDA allocation)
        double us[3];
        for (int t = 0; t < 3; t++) {
            int n = 0; sc[t].run(ths[t], &n);
            const auto t0 = std::chrono::steady_clock::now();
            for (int it = 0; it < 50; it++) sc[t].run(ths[t], &n);
            us[t] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / 50;
        }
        printf("{\"mismatches\": %d, \"iterations\": %d, \"threads\": 4, \"matches\": [%d, %d, %d], \"keypoints\": %zu, "
               "\"construct_search_destruct_us\": {\"1000kp_600mp\": %.1f, \"700kp_300mp\": %.1f, \"400kp_900mp\": %.1f}}\n",
               bad.load(), ITER, wantn[0], wantn[1], wantn[2], wkps.size(), us[0], us[1], us[2]);
        return bad.load() ? 1 : 0;
    } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 2; }
}
