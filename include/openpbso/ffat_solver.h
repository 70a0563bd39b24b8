// openpbso drop-in: run-time subset of ffat_solver.h -- the far-field acoustic transfer (FFAT) cube map that
// ModalSolver::computeTransfer evaluates (reference ffat_solver.h:235-295, 1180-1206).  Map construction /
// fitting / visualisation (the other ~1000 lines of the reference header) is offline preprocessing and is not
// part of this path.  A map's data lives on the host (for GetData / Check) and in a device set (pbso_ffat)
// that kernel K3 reads; GetMapVal evaluates on the B200.
#ifndef FFAT_SOLVER_H
#define FFAT_SOLVER_H
#include <cassert>
#include <map>
#include <memory>
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

    FFAT_Map() = default;
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
    int modeId = 0;

private:
    T _k = -1;
    T _cellSize = 0;
    FFAT_Vector3d _center;
    FFAT_MatrixXd _Psi;               // column 0 is what GetMapVal reads (reference :1203)
    bool _is_compressed = false;
    std::shared_ptr<pbso_ffat> _set;  // set this map was loaded into (keyed by modeId)
    mutable std::shared_ptr<pbso_ffat> _single;   // this map alone, re-keyed to id 0, built on first GetMapVal

    pbso_ffat* single() const {
        if (!_single) {
            double geom[32]; int igeom[18]; int n = 0, cols = 0, comp = 0;
            pbso_mirror::check(pbso_ffat_get_map(_set.get(), modeId, geom, igeom, &n, &cols, &comp, nullptr), "FFAT_Map");
            std::vector<double> psi((size_t)n * cols);
            pbso_mirror::check(pbso_ffat_get_map(_set.get(), modeId, nullptr, nullptr, nullptr, nullptr, nullptr, psi.data()), "FFAT_Map");
            const int id0 = 0; const unsigned char c = (unsigned char)comp;
            pbso_ffat* h = nullptr;
            pbso_mirror::check(pbso_ffat_create(1, &id0, geom, igeom, psi.data(), n, &c, &h), "FFAT_Map");
            _single = std::shared_ptr<pbso_ffat>(h, [](pbso_ffat* p) { pbso_ffat_destroy(p); });
        }
        return _single.get();
    }
    friend FFAT_Map_Serialize;
};
}  // namespace Gpu_Wavesolver
#endif
