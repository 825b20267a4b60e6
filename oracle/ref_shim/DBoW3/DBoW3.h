// oracle/ref_shim/DBoW3/DBoW3.h — TEST INFRASTRUCTURE ONLY: DBoW3::FeatureVector as Features/matcher.cpp walks it (an ordered map from
// vocabulary node id to the indices of the features under that node; DBoW3's own class derives from exactly this std::map).
#pragma once
#include <map>
#include <vector>
namespace DBoW3 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};
class BowVector : public std::map<unsigned int, double> {};
class Vocabulary;
}  // namespace DBoW3
