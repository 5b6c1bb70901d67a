#pragma once
#include <map>
#include <vector>
namespace fbow {
struct BoWVector : std::map<unsigned, float> {};
struct BoWFeatVector : std::map<unsigned, std::vector<unsigned>> {};
class Vocabulary {};
}
