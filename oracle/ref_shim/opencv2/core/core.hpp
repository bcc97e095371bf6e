// stand-in for <opencv2/core/core.hpp>: see uvip_cv_standin.hpp (test infrastructure)
#include "../../uvip_cv_standin.hpp"
