// stand-in for <opencv2/features2d/features2d.hpp>: see uvip_cv_standin.hpp (test infrastructure)
#include "../../uvip_cv_standin.hpp"
