// stand-in for <opencv2/highgui/highgui.hpp>: see uvip_cv_standin.hpp (test infrastructure)
#include "../../uvip_cv_standin.hpp"
