// stand-in for <opencv2/core/eigen.hpp>: src/hash.cpp includes it and uses nothing from it
