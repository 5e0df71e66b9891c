// Thirdparty/DBoW2/DBoW2/FeatureVector.h of the reference: node id -> indices of the local features under that node.
// (The reference's own header is used instead where the reference tree is on the include path.)
#ifndef ORBB200_COMPAT_FEATURE_VECTOR_H
#define ORBB200_COMPAT_FEATURE_VECTOR_H
#if defined(__has_include)
#if __has_include("Thirdparty/DBoW2/DBoW2/FeatureVector.h")
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#define ORBB200_HAVE_DBOW2_FEATURE_VECTOR
#endif
#endif
#ifndef ORBB200_HAVE_DBOW2_FEATURE_VECTOR
#include <map>
#include <vector>
namespace DBoW2
{
class FeatureVector : public std::map<unsigned int, std::vector<unsigned int> > {};
}
#endif
#endif
