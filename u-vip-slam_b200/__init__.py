"""u-vip-slam_b200 — B200-native ORB front-end (ORBextractor + ORBmatcher descriptor path) for U-VIP-SLAM.

The directory name has hyphens (it is the repository's package directory); import it through
`__graft_entry__.load_package()`, which registers it as module `uvip_slam_b200`.

Layout:  csrc/   hand-written sm_100a CUDA kernels + the C-ABI (include/uvip_orb.h) -> libuvip_orb.so
         host/   C++ drop-in shim classes USLAM::ORBextractor / USLAM::ORBmatcher over the C-ABI
         capi.py ctypes binding of the C-ABI;  frontend.py Python mirror of the two reference classes
         synth.py deterministic synthetic inputs;  build.py nvcc recipe
There is no CPU fallback: every compute entry point goes through libuvip_orb.so and raises if it is missing."""
from . import synth  # noqa: F401
from . import build  # noqa: F401
from . import capi  # noqa: F401
from . import frontend  # noqa: F401
from . import sharding  # noqa: F401
from .capi import UvipError, KP_DTYPE  # noqa: F401
from .frontend import ORBextractor, ORBmatcher, ORBVocabulary, KLTTracker  # noqa: F401
