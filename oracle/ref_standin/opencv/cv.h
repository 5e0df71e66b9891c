/* ORACLE -- TEST INFRASTRUCTURE ONLY: stand-in for <opencv/cv.h>, see ../cv_standin.hpp */
#include "../cv_standin.hpp"
