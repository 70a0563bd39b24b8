// TEST INFRASTRUCTURE ONLY -- stand-in for libigl's igl/serialize.h (libigl is not vendored in the
// reference tree: external/libigl/ is an empty submodule).  ffat_solver.h:71,236 derive FFAT_Map from
// igl::Serializable and call igl::serialize/deserialize only in the legacy SaveToFile/LoadFromFile
// members (ffat_solver.h:503-508,1066-1071), which the synthesis path never reaches.
#pragma once
#include <string>
namespace igl {
struct Serializable {
    virtual ~Serializable() {}
    virtual void InitSerialization() = 0;
    template <typename T> void Add(T&, const std::string&) {}
};
template <typename T> bool serialize(const T&, const std::string&, const std::string&, bool = false) { return false; }
template <typename T> bool deserialize(T&, const std::string&, const std::string&) { return false; }
}  // namespace igl
