#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
struct BowVector : std::map<unsigned, double> {};
struct FeatureVector : std::map<unsigned, std::vector<unsigned>> {};
template <class D, class F> class TemplatedVocabulary {};
}
