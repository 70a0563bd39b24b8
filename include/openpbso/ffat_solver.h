// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), ffat_solver.h (run-time subset)
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: ffat_solver.h -- the far-field acoustic transfer (FFAT) cube map that
// ModalSolver::computeTransfer evaluates (reference ffat_solver.h:235-295, 1180-1206) and its construction from
// shell pressure samples (constructor :944-989, Solve :1007-1069, ReadNElementsFile :1100-1118).  Visualisation,
// resampling and JPEG compression (the rest of the reference header) are not part of this path.  A map's data
// lives on the host (for GetData / Check) and in a device set (pbso_ffat) that kernel K3 reads; GetMapVal
// evaluates and Solve fits on the B200 (kernels K3 / K6).
#ifndef FFAT_SOLVER_H
#define FFAT_SOLVER_H
#include <cassert>
#include <complex>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include "Eigen/Dense"
#include "io.h"
#include "pbso_check.h"

namespace Gpu_Wavesolver {
struct FFAT_Map_Serialize_Double;
typedef FFAT_Map_Serialize_Double FFAT_Map_Serialize;
template <typename T, int M> class FFAT_Map { /* only M = 3 is on the synthesis path */ };

template <typename T>
class FFAT_Map<T, 3> {
public:
    typedef Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> FFAT_MatrixXd;
    typedef Eigen::Matrix<T, 3, 1> FFAT_Vector3d;
    typedef Eigen::Matrix<std::complex<T>, Eigen::Dynamic, 1> FFAT_VectorXcd;

    FFAT_Map() = default;
    // Shell geometry from the cube-map mesh (reference :944-989): V holds 4 vertices per quad, shells back to back;
    // N_elements[shell][face] = quads along the two in-face axes.  Shell 2 is the one the run-time map keeps.
    FFAT_Map(const int& modeId_, const T cellSize, const Eigen::Matrix<T, Eigen::Dynamic, 3>& V,
             const std::vector<std::vector<std::pair<int, int>>>& N_elements)
        : modeId(modeId_), _cellSize(cellSize) {
        const int N_shells = (int)N_elements.size();
        assert(N_shells > 1 && "need more than 1 shell");
        std::vector<int> ne; std::vector<double> v((size_t)V.rows() * 3);
        for (const auto& shell : N_elements) {
            assert(shell.size() == 6 && "N_elements wrong size");
            for (const auto& p : shell) { ne.push_back(p.first); ne.push_back(p.second); }
        }
        for (int i = 0; i < (int)V.rows(); ++i) for (int j = 0; j < 3; ++j) v[(size_t)i * 3 + j] = (double)V(i, j);
        pbso_ffat_fitter* h = nullptr;
        pbso_mirror::check(pbso_ffat_fitter_create((double)cellSize, v.data(), (int)V.rows(), ne.data(), N_shells, &h), "FFAT_Map::FFAT_Map");
        _fitter = std::shared_ptr<pbso_ffat_fitter>(h, [](pbso_ffat_fitter* p) { pbso_ffat_fitter_destroy(p); });
        double geom[32];
        pbso_mirror::check(pbso_ffat_fitter_shell(h, 2, geom, nullptr), "FFAT_Map::FFAT_Map");
        _center << (T)geom[28], (T)geom[29], (T)geom[30];
    }
    // Least-squares fit of Psi from the shells' Dirichlet pressure (reference :1007-1069); a repeated k returns at once.
    void Solve(const T& k, const FFAT_VectorXcd& dirichletPressure, const bool& powerScaling = false) {
        if (_k == k) return;
        assert(_fitter && "Solve started before proper initialization");
        int n_total = 0, n_dir = 0;
        pbso_mirror::check(pbso_ffat_fitter_info(_fitter.get(), nullptr, &n_total, &n_dir, nullptr), "FFAT_Map::Solve");
        assert(dirichletPressure.size() == 2 * n_total && "Dirichlet pressure wrong size.");
        std::vector<double> p((size_t)4 * n_total);
        for (int i = 0; i < 2 * n_total; ++i) { p[2 * (size_t)i] = (double)dirichletPressure(i).real(); p[2 * (size_t)i + 1] = (double)dirichletPressure(i).imag(); }
        const double kd = (double)k;
        std::vector<double> psi((size_t)n_dir);
        pbso_mirror::check(pbso_ffat_fitter_solve(_fitter.get(), 1, &kd, p.data(), powerScaling ? 1 : 0, psi.data(), nullptr), "FFAT_Map::Solve");
        _Psi.resize(n_dir, 1);
        for (int i = 0; i < n_dir; ++i) _Psi(i, 0) = (T)psi[(size_t)i];
        _k = k;
        // the run-time map (what FFAT_Map_Serialize::Save keeps): shell 2 + Psi + k
        double geom[32]; int igeom[18];
        pbso_mirror::check(pbso_ffat_fitter_shell(_fitter.get(), 2, geom, igeom), "FFAT_Map::Solve");
        geom[31] = kd;
        pbso_ffat* h = nullptr;
        pbso_mirror::check(pbso_ffat_create(1, &modeId, geom, igeom, psi.data(), n_dir, nullptr, &h), "FFAT_Map::Solve");
        _set = std::shared_ptr<pbso_ffat>(h, [](pbso_ffat* q) { pbso_ffat_destroy(q); });
        _single.reset();
    }
    // The LEGACY file form: igl::serialize of the whole object (reference :1066-1085; FFAT_Map_Serialize in ffat_map_serialize.h is
    // the protobuf form that replaced it).  Save writes it (shell 2 three times: GetMapVal reads shell 2 only); Load / LoadAll
    // read it -- and the protobuf form too, the library tells them apart by the first chunk header.  No libigl involved: the
    // format is restated in libpbso_b200 (csrc/fatcube_codec.h).
    static void Save(const char* filename, const FFAT_Map<T, 3>& map) {
        assert(map._set && "FFAT map is empty");
        pbso_mirror::check(pbso_ffat_save_legacy_file(map._set.get(), map.modeId, filename), "FFAT_Map::Save");
    }
    static void Load(const char* filename, FFAT_Map<T, 3>& map) {
        pbso_ffat* h = nullptr;
        pbso_mirror::check(pbso_ffat_load_file(filename, &h), "FFAT_Map::Load");
        std::shared_ptr<pbso_ffat> set(h, [](pbso_ffat* p) { pbso_ffat_destroy(p); });
        int id = 0;
        pbso_mirror::check(pbso_ffat_mode_ids(h, &id), "FFAT_Map::Load");
        FillFromSet(set, id, map);
    }
    static std::map<int, FFAT_Map<T, 3>>* LoadAll(const char* dirname) {
        std::vector<std::string> filenames;
        ListDirFiles(dirname, filenames, ".fatcube");
        auto* map = new std::map<int, FFAT_Map<T, 3>>();
        for (const auto& filename : filenames) {
            FFAT_Map<T, 3> map_;
            Load(filename.c_str(), map_);
            (*map)[map_.modeId] = map_;
        }
        return map;
    }
    // One line per shell: "Nx Ny" for the six faces (reference :1100-1118).
    static void ReadNElementsFile(const char* filename, std::vector<std::vector<std::pair<int, int>>>& N_elements) {
        std::ifstream stream(filename);
        assert(stream && "File not exist");
        std::string line;
        N_elements.clear();
        while (std::getline(stream, line)) {
            std::istringstream iss(line);
            std::vector<std::pair<int, int>> v(6);
            for (int ii = 0; ii < 6; ++ii) { std::pair<int, int> p; iss >> p.first >> p.second; v[ii] = p; }
            N_elements.push_back(v);
        }
    }
    inline FFAT_Vector3d GetCenter() const { return _center; }
    inline T GetCellSize() const { return _cellSize; }
    inline const FFAT_MatrixXd& GetData() const { return _Psi; }
    // |Psi(direction of p) / (k |p - centre|)| with texel-centre bilinear interpolation on the cube map.
    T GetMapVal(const FFAT_Vector3d& p, const bool getCompressed = false) const {
        if (getCompressed) assert(_is_compressed && "asking for compressed values without compression");
        assert(_set && "FFAT map is empty");
        const double pos[3] = {(double)p(0), (double)p(1), (double)p(2)};
        double out = 0;
        pbso_ffat* one = single();
        pbso_mirror::check(pbso_ffat_eval(one, 1, pos, 1, getCompressed ? 1 : 0, &out), "FFAT_Map::GetMapVal");
        return (T)out;
    }
    // 8-bit quantisation of the map, face by face (reference :1125-1178): _compressed_Psi = round(Psi 255/maxAmp) maxAmp/255.
    // and _is_compressed = true; returns maxAmp over all faces.  The reference writes every face as a JPEG file of the given
    // quality and reads it back in between (OpenCV); this build has no image codec, so output_template / quality are unused
    // and the bytes are kept as quantised -- a caller with a codec runs it between pbso_ffat_quantise and
    // pbso_ffat_set_compressed_u8 (include/pbso_b200.h).  The device then reads ONE byte per texel for
    // GetMapVal(p, true) / computeTransfer.
    T Compress(const char* output_template = "tmp-%u-%u-amp.jpg", const int quality = 65) {
        (void)output_template; (void)quality;
        assert(_set && "FFAT map is empty");
        double g = 0;
        pbso_mirror::check(pbso_ffat_compress(_set.get(), modeId, &g), "FFAT_Map::Compress");
        _is_compressed = true;
        _single.reset();
        return (T)g;
    }
    int modeId = 0;

private:
    T _k = -1;
    T _cellSize = 0;
    FFAT_Vector3d _center;
    FFAT_MatrixXd _Psi;               // column 0 is what GetMapVal reads (reference :1203)
    bool _is_compressed = false;
    std::shared_ptr<pbso_ffat> _set;  // set this map was loaded into (keyed by modeId)
    std::shared_ptr<pbso_ffat_fitter> _fitter;    // shell geometry, when built from a mesh
    mutable std::shared_ptr<pbso_ffat> _single;   // this map alone, re-keyed to id 0, built on first GetMapVal

    // host copy of one map of a loaded set (what both Load functions fill in)
    static void FillFromSet(const std::shared_ptr<pbso_ffat>& set, int id, FFAT_Map<T, 3>& map) {
        double geom[32]; int igeom[18]; int n = 0, cols = 0, comp = 0;
        pbso_mirror::check(pbso_ffat_get_map(set.get(), id, geom, igeom, &n, &cols, &comp, nullptr), "FFAT_Map");
        std::vector<double> psi((size_t)n * cols);
        pbso_mirror::check(pbso_ffat_get_map(set.get(), id, nullptr, nullptr, nullptr, nullptr, nullptr, psi.data()), "FFAT_Map");
        map._Psi.resize(n, cols);
        for (int c = 0; c < cols; ++c) for (int r = 0; r < n; ++r) map._Psi(r, c) = (T)psi[(size_t)c * n + r];
        map._cellSize = (T)geom[0];
        map._center << (T)geom[28], (T)geom[29], (T)geom[30];
        map._k = (T)geom[31];
        map._is_compressed = comp != 0;
        map.modeId = id;
        map._set = set;
        map._single.reset();
    }
    pbso_ffat* single() const {
        if (!_single) {
            double geom[32]; int igeom[18]; int n = 0, cols = 0, comp = 0;
            pbso_mirror::check(pbso_ffat_get_map(_set.get(), modeId, geom, igeom, &n, &cols, &comp, nullptr), "FFAT_Map");
            std::vector<double> psi((size_t)n * cols);
            pbso_mirror::check(pbso_ffat_get_map(_set.get(), modeId, nullptr, nullptr, nullptr, nullptr, nullptr, psi.data()), "FFAT_Map");
            const int id0 = 0; const unsigned char c = (unsigned char)comp;
            pbso_ffat* h = nullptr;
            std::vector<unsigned char> q8((size_t)n); double amp[6];
            const bool bytes = comp && pbso_ffat_get_compressed(_set.get(), modeId, q8.data(), amp, nullptr) == PBSO_OK;   // compressed in memory
            const unsigned char c1 = bytes ? 0 : c;
            pbso_mirror::check(pbso_ffat_create(1, &id0, geom, igeom, psi.data(), n, &c1, &h), "FFAT_Map");
            _single = std::shared_ptr<pbso_ffat>(h, [](pbso_ffat* p) { pbso_ffat_destroy(p); });
            if (bytes) pbso_mirror::check(pbso_ffat_set_compressed_u8(h, 0, q8.data(), n, amp), "FFAT_Map");
        }
        return _single.get();
    }
    friend FFAT_Map_Serialize;
};
}  // namespace Gpu_Wavesolver
#endif
