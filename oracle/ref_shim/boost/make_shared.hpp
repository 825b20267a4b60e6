// oracle/ref_shim/boost/make_shared.hpp — TEST INFRASTRUCTURE ONLY: boost::make_shared<T>() for the shim's pcl::PointCloud<T>::Ptr.
#pragma once
#include <memory>
namespace boost {
template <typename T, typename... A> std::shared_ptr<T> make_shared(A&&... a) { return std::make_shared<T>(static_cast<A&&>(a)...); }
}  // namespace boost
