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

// ---- the LEGACY .fatcube format: libigl's igl::serialize of the FFAT_Map<T,3> object (ffat_solver.h:978-991 lists the
// members, :1066-1071 Save / Load, :1075-1085 LoadAll; format: external/libigl/include/igl/serialize.h) ------------------
// A file is a stream of chunks  [string name][string type][u64 size][data]  (string = [u64 length][bytes]; `type` is the
// writer's typeid().name(), which a reader that knows the members has no use for).  The map is ONE chunk named
// "serial_map_ch3" whose data is [u64 inner size][chunks of the members]; int = 4 bytes, double = 8, bool = 1; an Eigen matrix =
// [i64 rows][i64 cols][column-major data]; std::vector = [u64 count][elements] where an int is raw, a std::pair<int,int> is 8
// bytes, a matrix is as above and a nested serializable object (the three shells in "maps") is [u64 inner size][chunks].
// GetMapVal reads shell 2 only (ffat_solver.h:1188-1203), which is what the run-time map keeps.
bool legacy_fatcube_sniff(const uint8_t* data, size_t size);              // starts with the "serial_map_ch3" chunk header?
bool legacy_fatcube_decode(const uint8_t* data, size_t size, FatcubeMap& out, std::string& err);
// Writes what FFAT_Map<T,3>::Save would for a map that holds `m` (all three shells get shell 2's geometry: only shell 2 is
// ever read back; type strings are g++'s, as in a file written by the reference built with GCC).
void legacy_fatcube_encode(const FatcubeMap& m, std::string& out);

// ListDirFiles(dirname, names, contains) (io.cpp:18-35): readdir order, skips dot entries, keeps
// entries whose FULL PATH contains `contains`.  Returns false when the directory cannot be opened.
bool list_dir_files(const char* dirname, std::vector<std::string>& names, const char* contains);

}  // namespace pbso
