// oracle/ref_shim/DBoW3/DBoW3.h — TEST INFRASTRUCTURE ONLY: DBoW3::FeatureVector as Features/matcher.cpp walks it (an ordered map from
// vocabulary node id to the indices of the features under that node; DBoW3's own class derives from exactly this std::map).
#pragma once
#include <cstdlib>
#include <map>
#include <vector>
namespace DBoW3 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};
class BowVector : public std::map<unsigned int, double> {};
// DBoW3::Vocabulary::transform (Frame::ComputeBoW): vocabulary queries are out of scope; stops the process if reached
class Vocabulary {
public:
    template <typename D> void transform(const D&, BowVector&, FeatureVector&, int) { std::abort(); }
};
}  // namespace DBoW3
