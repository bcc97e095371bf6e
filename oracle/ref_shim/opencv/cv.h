// stand-in for <opencv/cv.h>: see uvip_cv_standin.hpp (test infrastructure)
#include "../uvip_cv_standin.hpp"
