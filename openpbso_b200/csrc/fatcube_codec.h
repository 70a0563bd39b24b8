// Hand-written proto3 wire codec for ffat_map.proto (reference ffat_map.proto:12-51).
// There is no protoc / libprotobuf in the build image (and the reference vendors only protobuf
// headers), so the loader speaks the wire format directly.  Host-only C++.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace pbso {

// The fields FFAT_Map_Serialize_Double keeps (ffat_map_serialize.h:55-79).
struct FatcubeMap {
    // ffat_map_t_1 (shell #2)
    double cellsize = 0;
    std::vector<std::vector<double>> lowcorners;   // mat: one vec per face
    std::vector<std::vector<int>> n_elements;      // mat_i: one vec_i per face
    std::vector<int> strides;
    std::vector<double> center1, bboxlow, bboxtop;
    // ffat_map_t_3
    double k = 0;
    std::vector<double> center3;
    bool is_compressed = false;
    std::vector<std::vector<double>> psi;          // column-major: one vec per column
    int modeid = 0;
};

// Parses a serialized ffat_map_double.  Accepts packed and unpacked repeated scalars, skips unknown
// fields, merges repeated occurrences of a submessage (last-wins for scalars), as protobuf does.
bool fatcube_decode(const uint8_t* data, size_t size, FatcubeMap& out, std::string& err);
// Canonical proto3 encoding: ascending field numbers, packed repeated scalars, zero-valued scalars
// (incl. -0.0, as protobuf 3.7.1 generated code tests `!= 0`) omitted, every submessage FFAT_Map_Serialize_Double::Save touches is emitted even when empty.
void fatcube_encode(const FatcubeMap& m, std::string& out);

// ListDirFiles(dirname, names, contains) (io.cpp:18-35): readdir order, skips dot entries, keeps
// entries whose FULL PATH contains `contains`.  Returns false when the directory cannot be opened.
bool list_dir_files(const char* dirname, std::vector<std::string>& names, const char* contains);

}  // namespace pbso
