/* ORACLE -- TEST INFRASTRUCTURE ONLY: stand-in for <opencv2/opencv.hpp>, see ../cv_standin.hpp */
#include "../cv_standin.hpp"
