// stand-in for <opencv2/core/core.hpp> where the OpenCV SDK is absent: see ../../../cv_compat.h
#include "../../../cv_compat.h"
