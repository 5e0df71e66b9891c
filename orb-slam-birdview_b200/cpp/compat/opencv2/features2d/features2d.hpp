// stand-in for <opencv2/features2d/features2d.hpp> where the OpenCV SDK is absent: see ../../../cv_compat.h
#include "../../../cv_compat.h"
