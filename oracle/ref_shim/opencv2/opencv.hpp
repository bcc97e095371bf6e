// stand-in for <opencv2/opencv.hpp>: see uvip_cv_standin.hpp (test infrastructure)
#include "../uvip_cv_standin.hpp"
