// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), ffat_map_serialize.h
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: FFAT_Map_Serialize (reference ffat_map_serialize.h:80-331).  `.fatcube` files are
// protobuf messages (ffat_map.proto:12-51); the wire codec is inside libpbso_b200 (no protoc / libprotobuf
// needed).  Save / Load / LoadAll / Check keep the reference's signatures and ownership.
#ifndef FFAT_MAP_SERIALIZE
#define FFAT_MAP_SERIALIZE
#include <cstring>
#include <map>
#include <memory>
#include <vector>
#include "ffat_solver.h"
#include "io.h"

namespace Gpu_Wavesolver {
struct FFAT_Map_Serialize_Double {
    static void Save(const char* filename, const FFAT_Map<double, 3>& map) {
        assert(map._set && "FFAT map is empty");
        pbso_mirror::check(pbso_ffat_save_file(map._set.get(), map.modeId, filename), "FFAT_Map_Serialize::Save");
    }
    static void Load(const char* filename, FFAT_Map<double, 3>& map) {
        pbso_ffat* h = nullptr;
        pbso_mirror::check(pbso_ffat_load_file(filename, &h), "FFAT_Map_Serialize::Load");
        std::shared_ptr<pbso_ffat> set(h, [](pbso_ffat* p) { pbso_ffat_destroy(p); });
        int id = 0;
        pbso_mirror::check(pbso_ffat_mode_ids(h, &id), "FFAT_Map_Serialize::Load");
        Fill(set, id, map);
    }
    // Every "*.fatcube" in dirname, keyed by modeId (later files overwrite earlier ones with the same id).
    static std::map<int, FFAT_Map<double, 3>>* LoadAll(const char* dirname) {
        auto* out = new std::map<int, FFAT_Map<double, 3>>();
        std::vector<std::string> filenames;
        ListDirFiles(dirname, filenames, ".fatcube");
        for (const auto& f : filenames) {
            FFAT_Map<double, 3> m;
            Load(f.c_str(), m);
            (*out)[m.modeId] = m;
        }
        return out;
    }
    template <typename T>
    static bool MatchBits(const T* data1, const T* data2, const int size) {
        return std::memcmp(data1, data2, sizeof(T) * (size_t)size) == 0;
    }
    // Bitwise equality of everything Save writes (reference :281-329).
    static bool Check(const FFAT_Map<double, 3>& map1, const FFAT_Map<double, 3>& map2) {
        double g1[32], g2[32]; int i1[18], i2[18]; int n1, n2, c1, c2, z1, z2;
        pbso_mirror::check(pbso_ffat_get_map(map1._set.get(), map1.modeId, g1, i1, &n1, &c1, &z1, nullptr), "Check");
        pbso_mirror::check(pbso_ffat_get_map(map2._set.get(), map2.modeId, g2, i2, &n2, &c2, &z2, nullptr), "Check");
        bool match = MatchBits(g1, g2, 32) && MatchBits(i1, i2, 18) && n1 == n2 && c1 == c2 && z1 == z2;
        match &= map1.modeId == map2.modeId;
        if (match) match &= MatchBits(map1._Psi.data(), map2._Psi.data(), n1 * c1);
        return match;
    }
private:
    static void Fill(const std::shared_ptr<pbso_ffat>& set, int id, FFAT_Map<double, 3>& map) { FFAT_Map<double, 3>::FillFromSet(set, id, map); }
};
typedef FFAT_Map_Serialize_Double FFAT_Map_Serialize;
}  // namespace Gpu_Wavesolver
#endif
