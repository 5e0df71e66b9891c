/* ORACLE -- TEST INFRASTRUCTURE ONLY: stand-in for <opencv2/features2d/features2d.hpp>, see ../../cv_standin.hpp */
#include "../../cv_standin.hpp"
