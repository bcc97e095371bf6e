// test_shim_m8.cpp — the drop-in shim's keyframe-side searches (u-vip-slam_b200/host/ORBmatcher.h: Fuse x2,
// SearchByProjection(KF, Scw, ...), SearchBySim3) driven through the SAME scene code and the SAME stand-in SLAM types the
// reference's own ORBmatcher.cc is compiled against (oracle/ref_shim/m8_scene.h), so that the two result bundles can be
// compared byte for byte.  Prebuilt into oracle/_ref/test_shim_m8 by `make -C oracle ref` (slam_standin.h uses DBoW2's
// FeatureVector from the reference tree, which does not exist on the GPU box).
//   test_shim_m8 <scene bundle> <result bundle> <which>      exit 0 ok, 3 no CUDA device ("no CPU fallback"), 2 I/O
#include <stdio.h>
#include <stdexcept>
#include "slam_standin.h"
#include "../../u-vip-slam_b200/host/ORBmatcher.h"
#include "m8_scene.h"

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s scene out which\n", argv[0]); return 2; }
    m8::Bundle in, out;
    if (!in.load(argv[1]) || (int)in.a.size() < m8::A_COUNT) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    try {
        const int r = m8::run<USLAM::ORBmatcher>(atoi(argv[3]), in, out);
        if (!out.save(argv[2])) return 2;
        printf("%d\n", r);
    } catch (const std::runtime_error& e) {
        fprintf(stderr, "shim: %s (no CPU fallback)\n", e.what());
        return 3;
    }
    return 0;
}
