// =============================================================================
// pbso_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++17 / IEEE-double restatement of the openpbso modal-synthesis hot
// path.  Nothing in the product (openpbso_b200/,
// include/) may include, link or call this file; only tests/, the smoke check
// in __graft_entry__.py and the cpu_baseline / --impl reference legs of
// bench.py use it, and there only as the checker or the timed CPU baseline.
//
// Parity status: the reference ships no tests or golden vectors for this path
// (SURVEY.md section 4).  Rows marked [pinned:_ref] below are additionally
// cross-checked in tests/test_oracle_vs_ref.py against the reference's OWN
// headers compiled from /root/reference (oracle/Makefile target `ref` ->
// oracle/_ref/libpbso_ref.so; Eigen is absent in the image so those headers
// are compiled against the minimal shim include/openpbso/eigen_shim/Eigen/Dense).
// Rows without that mark would be "parity unpinned" (anchored only on analytic
// known-answer tests, tests/test_oracle_kat.py); at the end of round 1 every row
// carries the mark except libigl's per_vertex_normals (oracle.py): libigl
// needs the real Eigen, which is absent, so it cannot be compiled here.
//
// Each function cites the reference file:line it follows (paths relative to
// /root/reference).  No Eigen: every Eigen expression on the path is an
// element-wise loop or a dot product and is restated as a plain loop.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <list>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

namespace orc {

static const int SAMPLE_RATE = 44100;  // config.h:13

// -----------------------------------------------------------------------------
// ModalIntegrator   [pinned:_ref]
// -----------------------------------------------------------------------------
// modal_integrator.h:47-70  Build(): (rho, omega^2, alpha, beta) -> (a, b)
static void build_ab(double density, const double* omegaSquared, int N,
                     double alpha, double beta, double* a, double* b) {
    for (int ii = 0; ii < N; ++ii) {
        double omega = std::sqrt(omegaSquared[ii] / density);      // :63
        double xi = 0.5 * (alpha / omega + beta * omega);           // :64
        a[ii] = 2.0 * xi * omega;                                   // :65
        b[ii] = std::pow(omega, 2);                                 // :66
    }
}

// modal_integrator.h:86-100  ctor: DyRT IIR coefficients incl. the 1E9 scale
static void coeffs(int N, double h, const double* a, const double* b,
                   double* c1, double* c2, double* c3) {
    for (int ii = 0; ii < N; ++ii) {
        double epsilon = std::exp(-a[ii] / 2 * h);                  // :89
        double theta = h * std::sqrt(b[ii] - a[ii] * a[ii] / 4.0);  // :90
        double gamma = std::asin(a[ii] / (2.0 * std::sqrt(b[ii]))); // :91
        double omega = std::sqrt(b[ii]);                            // :92
        double omega_d = std::sqrt(b[ii] - std::pow(a[ii], 2) / 4.0);  // :93
        c1[ii] = 2.0 * epsilon * std::cos(theta);                   // :95
        c2[ii] = -std::pow(epsilon, 2);                             // :96
        c3[ii] = 2.0 * (epsilon * std::cos(theta + gamma) -
                        std::pow(epsilon, 2) * std::cos(2.0 * theta + gamma));  // :97
        c3[ii] /= (3.0 * omega * omega_d);                          // :98
        c3[ii] *= 1E9;                                              // :99
    }
}

struct Integrator {
    int N;
    double h;
    std::vector<double> a, b, c1, c2, c3;
    std::vector<double> q[3];   // modal_integrator.h:24 three-slot ring
    int ptr = 0;                // :29
    Integrator(int N_, double h_, const double* a_, const double* b_)
        : N(N_), h(h_), a(a_, a_ + N_), b(b_, b_ + N_), c1(N_), c2(N_), c3(N_) {
        for (auto& v : q) v.assign(N, 0.0);                         // :76-78
        coeffs(N, h, a.data(), b.data(), c1.data(), c2.data(), c3.data());
    }
    // modal_integrator.h:103-123  Step(Q) / Step()
    const double* step(const double* Q) {
        double* q_k = q[(ptr + 1) % 3].data();                      // :105
        const double* q_km1 = q[(ptr) % 3].data();                  // :106
        const double* q_km2 = q[(ptr + 2) % 3].data();              // :107
        if (Q) {
            for (int i = 0; i < N; ++i)
                q_k[i] = c1[i] * q_km1[i] + c2[i] * q_km2[i] + c3[i] * Q[i];  // :109-110
        } else {
            for (int i = 0; i < N; ++i)
                q_k[i] = c1[i] * q_km1[i] + c2[i] * q_km2[i];       // :120
        }
        ptr = (ptr + 1) % 3;                                        // :111
        return q_k;
    }
};

// -----------------------------------------------------------------------------
// Forces  (forces.h)   [pinned:_ref]
// -----------------------------------------------------------------------------
enum ForceType { PointForceT = 0, GaussianForceT = 1, AutoregressiveForceT = 2 };  // forces.h:12-16

struct Force {
    virtual bool Add(double* spread, int BUF) = 0;                  // forces.h:21
    virtual Force* clone() const = 0;
    virtual ~Force() = default;
};
struct PointForce : Force {                                         // forces.h:26-31
    bool used = false;
    bool Add(double* spread, int) override {                        // :81-90
        if (used) return false;
        spread[0] += 1.;
        used = true;
        return true;
    }
    Force* clone() const override { return new PointForce(*this); }
};
struct GaussianForce : Force {                                      // forces.h:33-48
    double _width;
    int _widthSamples;
    int _count = 0;
    int _center;
    int _cutoff = 5;
    explicit GaussianForce(double width) : _width(width) {
        _widthSamples = std::max(1, (int)(_width / 1000000. * SAMPLE_RATE));  // :44
        _center = (int)((_cutoff - 0.5) * _widthSamples);           // :45
    }
    bool Add(double* spread, int BUF) override {                    // :92-105
        if (_width == 0 || _count >= _cutoff * 2 * _widthSamples) return false;
        for (int ii = 0; ii < BUF; ++ii) {
            const double p =
                -0.5 * std::pow((double)(_count + ii - _center) / (double)_widthSamples, 2);
            spread[ii] += std::exp(p);
        }
        _count += BUF;
        return true;
    }
    Force* clone() const override { return new GaussianForce(*this); }
};
struct AutoregressiveForce : Force {                                // forces.h:60-79
    std::vector<double> _buf{0, 0, 0};
    int _bufLen = 3;
    int _bufIdx = 0;
    std::vector<double> _a{0.783, 0.116};
    double _sigma = 0.00148;
    double _mu = 0.142;
    std::default_random_engine _generator;                          // :69 default seed
    std::normal_distribution<double> _distribution;                 // :70
    double GetMuEffective() {                                       // :107-117
        double mu_tilde = 0.0;
        for (int ii = 0; ii < 2; ++ii)
            mu_tilde += _a.at(ii) * _buf.at((_bufIdx + _bufLen - ii - 1) % _bufLen);
        mu_tilde += _sigma * _distribution(_generator);
        _buf.at(_bufIdx) = mu_tilde;
        _bufIdx = (_bufIdx + 1) % _bufLen;
        return _mu + mu_tilde;
    }
    bool Add(double* spread, int BUF) override {                    // :119-128
        for (int ii = 0; ii < BUF; ++ii) spread[ii] += GetMuEffective();
        return true;
    }
    void SetParam(double a0, double a1, double sigma, double mu) {  // :130-137
        _buf = {0, 0, 0};
        _a = {a0, a1};
        _sigma = sigma;
        _mu = mu;
    }
    Force* clone() const override { return new AutoregressiveForce(*this); }
};

// -----------------------------------------------------------------------------
// ModalSolver   (modal_solver.h)   [pinned:_ref -- the reference's own ModalSolver<double,BUF>::step and queues,
//  buffer by buffer on the committed force script, cfg1 and a cfg5-style batch: tests/test_oracle_vs_ref.py]
// -----------------------------------------------------------------------------
struct ForceMessage {                                               // modal_solver.h:27-77
    std::vector<double> data;
    int forceType = PointForceT;
    std::unique_ptr<Force> force;
    bool sustainedForceStart = false, sustainedForceEnd = false, clearAllForces = false;
    ForceMessage() : force(new PointForce()) {}
    ForceMessage(const ForceMessage& t)                             // :39-47 deep copy
        : data(t.data), forceType(t.forceType), force(t.force->clone()),
          sustainedForceStart(t.sustainedForceStart),
          sustainedForceEnd(t.sustainedForceEnd), clearAllForces(t.clearAllForces) {}
    ForceMessage& operator=(const ForceMessage& t) {                // :48-59
        if (&t == this) return *this;
        data = t.data; forceType = t.forceType;
        sustainedForceStart = t.sustainedForceStart;
        sustainedForceEnd = t.sustainedForceEnd;
        clearAllForces = t.clearAllForces;
        force.reset(t.force->clone());
        return *this;
    }
};
struct ArParam { double a0, a1, sigma, mu; };                      // forces.h:50-55

// moodycamel::ReaderWriterQueue(maxSize) holds ceilToPow2(maxSize+1)-1 items
// without allocating (external/readerwriterqueue.h:101); try_enqueue fails
// beyond that.
static size_t rwq_capacity(size_t maxSize) {
    size_t x = maxSize + 1, p = 1;
    while (p < x) p <<= 1;
    return p - 1;
}

struct Solver {
    int N, BUF;
    std::shared_ptr<Integrator> integ;
    std::deque<ForceMessage> q_force;  size_t cap_force = rwq_capacity(512);   // modal_solver.h:129
    std::deque<std::vector<double>> q_trans; size_t cap_trans = rwq_capacity(1);  // :131
    std::deque<ArParam> q_arprm; size_t cap_arprm = rwq_capacity(1);           // :133
    std::list<ForceMessage> active;                                  // :118
    std::vector<double> space, time_, latest_transfer, qnorm, sound; // :119-120,113,114,110
    bool useTransfer = true, useTransferCache = true, sustained = false;  // :137-138,125
    Solver(int N_, int BUF_) : N(N_), BUF(BUF_), space(N_, 0.0), time_(BUF_, 0.0),
        latest_transfer(N_, 1E7), qnorm(N_, 0.0), sound(BUF_, 0.0) {}  // :89-92,134-141

    // modal_solver.h:181-276.  Returns 1 if a sound buffer was produced, 0 if
    // the step returned early (clearAllForces, :186-189).
    int step() {
        if (!q_force.empty()) {                                      // :184
            ForceMessage mess = q_force.front(); q_force.pop_front();
            if (mess.clearAllForces) { active.clear(); return 0; }   // :186-189
            if (mess.sustainedForceStart) {                          // :190-194
                active.clear(); sustained = true; active.push_back(mess);
            }
            if (!sustained) active.push_back(mess);                  // :195-196
            else active.begin()->data = mess.data;                   // :197-200
            if (mess.sustainedForceEnd) { active.clear(); sustained = false; }  // :201-204
        }
        std::fill(time_.begin(), time_.end(), 0.0);                  // :206
        if (!sustained) {                                            // :207-221
            std::fill(space.begin(), space.end(), 0.0);
            auto it = active.begin();
            while (it != active.end()) {
                bool added = it->force->Add(time_.data(), BUF);
                if (!added) { active.erase(it++); }
                else {
                    for (int i = 0; i < N; ++i) space[i] += it->data[i];
                    ++it;
                }
            }
        } else {                                                     // :222-240
            auto it = active.begin();
            if (it->forceType == AutoregressiveForceT) {
                if (!q_arprm.empty()) {
                    ArParam p = q_arprm.front(); q_arprm.pop_front();
                    static_cast<AutoregressiveForce*>(it->force.get())
                        ->SetParam(p.a0, p.a1, p.sigma, p.mu);
                }
            }
            it->force->Add(time_.data(), BUF);
            space = it->data;
        }
        bool use = useTransfer;                                      // :243-248 (lock always free here)
        if (use) {
            if (!q_trans.empty()) { latest_transfer = q_trans.front(); q_trans.pop_front(); }  // :250-252
        } else {
            latest_transfer.assign(N, 1E7);                          // :254, 89-92
        }
        useTransferCache = use;                                      // :256
        std::fill(qnorm.begin(), qnorm.end(), 0.0);                  // :262
        std::vector<double> Q(N);
        const int nt = (int)latest_transfer.size();
        for (int ii = 0; ii < BUF; ++ii) {                           // :263-271
            for (int m = 0; m < N; ++m) Q[m] = space[m] * time_[ii];
            const double* q = integ->step(Q.data());
            double s = 0.0;
            for (int m = 0; m < nt; ++m) s += q[m] * latest_transfer[m];
            sound[ii] = s;
            for (int m = 0; m < N; ++m) qnorm[m] += q[m] * q[m];
        }
        for (int m = 0; m < N; ++m) qnorm[m] = std::sqrt(qnorm[m]);  // :272
        return 1;
    }
};

// -----------------------------------------------------------------------------
// Impulse projection U^T f   (tools/real_time_modal_sound.cpp:236-295)
// [pinned: the reference tool's own GetModalForceVertex/Face, cut out at test time and compiled in place:
//  tests/test_oracle_vs_ref.py::test_projection_matches_the_reference_tool_functions]
// U is mode-major: U[m*nDOF + d]  (ModeData.h:23-24, 61-83)
// -----------------------------------------------------------------------------
static void project_vertex(int forceDim, const double* U, int nDOF, int vid,
                           const double vn[3], double* out) {        // :268-280
    for (int mm = 0; mm < forceDim; ++mm) {
        const double* mode = U + (size_t)mm * nDOF;
        out[mm] = vn[0] * mode[vid * 3 + 0] + vn[1] * mode[vid * 3 + 1] +
                  vn[2] * mode[vid * 3 + 2];
    }
}
static void project_face(int forceDim, const double* U, int nDOF, const int vids[3],
                         const double coords[3], const double vn[3], double* out) {  // :236-251
    for (int mm = 0; mm < forceDim; ++mm) {
        const double* mode = U + (size_t)mm * nDOF;
        double acc = 0.0;
        for (int jj = 0; jj < 3; ++jj) {
            acc += vn[0] * mode[vids[jj] * 3 + 0] * coords[jj] +
                   vn[1] * mode[vids[jj] * 3 + 1] * coords[jj] +
                   vn[2] * mode[vids[jj] * 3 + 2] * coords[jj];
        }
        out[mm] = acc;
    }
}

// -----------------------------------------------------------------------------
// FFAT map evaluation   (ffat_solver.h)   [pinned:_ref -- the reference's own LoadAll + |GetMapVal| at ~1000 probe
//  positions incl. ties, tests/test_oracle_vs_ref.py, and the reference-generated fixture tests/golden/ffat_eval.npz;
//  KATs (texel-centre identity, 1/r law, continuity) in tests/test_oracle_kat.py]
// -----------------------------------------------------------------------------
struct FFATMap {                 // fields kept by ffat_map_serialize.h:55-79
    double k;                    // FFAT_Map<T,3>::_k
    double center3[3];           // FFAT_Map<T,3>::_center
    double cellSize;             // FFAT_Map<T,1>::_cellSize  (shell #2)
    double lowCorners[6][3];
    int    nElem[6][2];
    int    strides[6];
    double center1[3], bboxLow[3], bboxTop[3];
    const double* Psi;           // column 0 of _Psi (length D)
    int D;
};

// ffat_solver.h:676-712  Intersect
static void ffat_intersect(const FFATMap& M, const double p[3], double surf[3], int ind[3]) {
    double d[3], t_enter[3];
    for (int i = 0; i < 3; ++i) {
        d[i] = M.center1[i] - p[i];                                  // :681
        double tmin = (M.bboxLow[i] - p[i]) / d[i];                  // :682
        double tmax = (M.bboxTop[i] - p[i]) / d[i];                  // :683
        t_enter[i] = std::min(tmin, tmax);                           // :684 (Eigen array min: a<b?a:b order below)
    }
    // Eigen's maxCoeff is a plain running max                       // :685
    double t_en = t_enter[0];
    for (int i = 1; i < 3; ++i) if (t_enter[i] > t_en) t_en = t_enter[i];
    for (int i = 0; i < 3; ++i) surf[i] = p[i] + t_en * d[i];        // :686
    double minDist = std::numeric_limits<double>::max();             // :688
    ind[0] = 0;
    for (int dd = 0; dd < 3; ++dd) {                                 // :689-698
        if (std::abs(M.bboxLow[dd] - surf[dd]) < minDist) {
            minDist = std::abs(M.bboxLow[dd] - surf[dd]);
            ind[0] = dd * 2 + 1;
        }
        if (std::abs(M.bboxTop[dd] - surf[dd]) < minDist) {
            minDist = std::abs(M.bboxTop[dd] - surf[dd]);
            ind[0] = dd * 2;
        }
    }
    int dk = ind[0] / 2, di = (dk + 1) % 3, dj = (dk + 2) % 3;       // :699-702
    auto Clamp = [](int x, int l, int h) { return std::min(std::max(x, l), h); };
    ind[1] = (int)std::floor((surf[di] - M.lowCorners[ind[0]][di]) / M.cellSize);  // :706-707
    ind[2] = (int)std::floor((surf[dj] - M.lowCorners[ind[0]][dj]) / M.cellSize);  // :708-709
    ind[1] = Clamp(ind[1], 0, M.nElem[ind[0]][0] - 1);               // :710
    ind[2] = Clamp(ind[2], 0, M.nElem[ind[0]][1] - 1);               // :711
}

// ffat_solver.h:736-803  Interpolate -> 4 (face,x,y) indices + bilinear weights
static void ffat_interpolate(const FFATMap& M, const double surf[3], const int nn[3],
                             int idx[4][3], double co[4]) {
    int dk = nn[0] / 2, di = (dk + 1) % 3, dj = (dk + 2) % 3;
    int x, y, xp, yp;
    double tx, ty;
    const int Nx = M.nElem[nn[0]][0], Ny = M.nElem[nn[0]][1];
    const double* low = M.lowCorners[nn[0]];
    const double h = M.cellSize;
    double x_float = (surf[di] - (low[di] + 0.5 * h)) / h;           // :757
    double y_float = (surf[dj] - (low[dj] + 0.5 * h)) / h;           // :758
    x = (int)std::floor(x_float);
    y = (int)std::floor(y_float);
    if (x < 0) { x = 0; xp = 0; tx = 0; }                            // :763-776
    else if (x >= 0 && x < Nx - 1) { xp = x + 1; tx = x_float - (double)x; }
    else { x = Nx - 1; xp = Nx - 1; tx = 0; }
    if (y < 0) { y = 0; yp = 0; ty = 0; }                            // :777-790
    else if (y >= 0 && y < Ny - 1) { yp = y + 1; ty = y_float - (double)y; }
    else { y = Ny - 1; yp = Ny - 1; ty = 0; }
    tx = std::min(std::max(tx, 0.0), 1.0);                           // :791
    ty = std::min(std::max(ty, 0.0), 1.0);                           // :792
    int f = nn[0];
    idx[0][0] = f; idx[0][1] = x;  idx[0][2] = y;                    // :795-798
    idx[1][0] = f; idx[1][1] = xp; idx[1][2] = y;
    idx[2][0] = f; idx[2][1] = x;  idx[2][2] = yp;
    idx[3][0] = f; idx[3][1] = xp; idx[3][2] = yp;
    co[0] = (1.0 - tx) * (1.0 - ty);                                 // :799-802
    co[1] = tx * (1.0 - ty);
    co[2] = (1.0 - tx) * ty;
    co[3] = tx * ty;
}

// ffat_solver.h:1180-1206 GetMapVal, :141-144 GetDataQuadStride, :899-906 Reconstruct
static double ffat_getmapval(const FFATMap& M, const double p[3]) {
    double surf[3]; int ind[3];
    ffat_intersect(M, p, surf, ind);
    int idx[4][3]; double co[4];
    ffat_interpolate(M, surf, ind, idx, co);
    double psi0 = 0.0;
    for (int kk = 0; kk < 4; ++kk) {
        int id = M.strides[idx[kk][0]] + idx[kk][1] * M.nElem[idx[kk][0]][1] + idx[kk][2];  // :141-144
        psi0 += co[kk] * M.Psi[id];                                  // :1203
    }
    double dx = p[0] - M.center3[0], dy = p[1] - M.center3[1], dz = p[2] - M.center3[2];
    double r = std::sqrt(dx * dx + dy * dy + dz * dz);               // :1205 (p-_center).norm()
    const double kr = M.k * r;                                       // :904
    return std::abs(psi0 / kr);                                      // :905
}

// -----------------------------------------------------------------------------
// FFAT map CONSTRUCTION   [pinned:_ref, with the shim's JacobiSVD -- see
// tests/test_oracle_vs_ref.py::test_ffat_fit_*]
// -----------------------------------------------------------------------------
// ffat_solver.h:399-428  FFAT_Map<T,1>::FFAT_Map(modeId, cellSize, V, N_elements): V holds 4 vertices per
// quad (CubemapMesh, :334-397); only the first vertex of each face's first quad is read (the low corner).
// _bboxLow/_bboxTop are never initialised by the reference (:420-427 min/max against whatever the members
// hold: undefined behaviour).  `bbox_init` is that starting value: 0 is what a zero-filled frame -- and the
// Eigen shim's value-initialised Matrix -- gives; it yields the true bounding box whenever the box straddles
// the origin, which is the only case the reference can be said to define.
static void ffat_shell_from_vertices(double cellSize, const double* V /*[rows][3]*/, const int nElem[6][2],
                                     double bbox_init, FFATMap& M, int& n_total) {
    M.cellSize = cellSize;
    int sum = 0;
    for (int f = 0; f < 6; ++f) {
        const int N = nElem[f][0] * nElem[f][1];                     // :408
        for (int d = 0; d < 3; ++d) M.lowCorners[f][d] = V[(size_t)(sum * 4) * 3 + d];   // :412-413
        M.nElem[f][0] = nElem[f][0]; M.nElem[f][1] = nElem[f][1];
        M.strides[f] = sum;                                          // :414
        sum += N;                                                    // :415
    }
    n_total = sum;                                                   // :418
    M.center1[0] = (M.lowCorners[0][0] + M.lowCorners[1][0]) / 2.0;  // :419-422
    M.center1[1] = (M.lowCorners[2][1] + M.lowCorners[3][1]) / 2.0;
    M.center1[2] = (M.lowCorners[4][2] + M.lowCorners[5][2]) / 2.0;
    for (int j = 0; j < 3; ++j) { M.bboxLow[j] = bbox_init; M.bboxTop[j] = bbox_init; }
    for (int f = 0; f < 6; ++f)                                      // :423-428
        for (int j = 0; j < 3; ++j) {
            M.bboxLow[j] = std::min(M.bboxLow[j], M.lowCorners[f][j]);
            M.bboxTop[j] = std::max(M.bboxTop[j], M.lowCorners[f][j]);
        }
}

struct FFATFit {                 // FFAT_Map<T,3> before Solve (ffat_solver.h:944-989)
    std::vector<FFATMap> shells;
    std::vector<int> strides;    // FFAT_Map<T,3>::_strides: quads before shell s
    int n_total = 0, n_dir = 0;
    double center3[3];
};
static void ffat_fit_geometry(double cellSize, const double* V, const int* nElem /*[S][6][2]*/, int S,
                              double bbox_init, FFATFit& F) {
    F.shells.resize(S); F.strides.resize(S);
    int offset = 0;                                                  // rows of V consumed (:969-977)
    F.n_total = 0;
    for (int s = 0; s < S; ++s) {
        int ne[6][2]; std::memcpy(ne, nElem + (size_t)s * 12, sizeof(ne));
        int sum = 0;
        ffat_shell_from_vertices(cellSize, V + (size_t)offset * 3, ne, bbox_init, F.shells[s], sum);
        F.strides[s] = F.n_total;                                    // :963
        F.n_total += sum;                                            // :964
        offset += sum * 4;                                           // :976
    }
    std::memcpy(F.center3, F.shells[2].center1, sizeof(F.center3));  // :982
    F.n_dir = 0;                                                     // :983-986
    for (int f = 0; f < 6; ++f) F.n_dir += F.shells[2].nElem[f][0] * F.shells[2].nElem[f][1];
}

// ffat_solver.h:1007-1069 FFAT_Map<T,3>::Solve -> :872-897 FFAT_Solver<T,3>::Solve -> :909-929 Scaling.
// pressure: complex interleaved (re, im), 2 * n_total entries (two triangles per quad; the even one is read).
// R_out / Pabs_out (optional): n_dir x S row-major, the radii and |interpolated pressure| per shell.
static double ffat_fit_solve(const FFATFit& F, double k, const double* pressure, bool powerScaling,
                             double* Psi, double* R_out, double* Pabs_out) {
    const int S = (int)F.shells.size();
    const FFATMap& outer = F.shells[2];
    std::vector<double> R((size_t)F.n_dir * S), Pre((size_t)F.n_dir * S), Pim((size_t)F.n_dir * S);
    int offset = 0;
    for (int dd = 0; dd < 6; ++dd) {                                 // :1021-1063
        const int dk = dd / 2, di = (dk + 1) % 3, dj = (dk + 2) % 3;
        const int dim1 = outer.nElem[dd][0], dim2 = outer.nElem[dd][1];
        double ijk[3]; ijk[dk] = 0;
        for (int ii = 0; ii < dim1; ++ii) {
            ijk[di] = 0.5 + ii;
            for (int jj = 0; jj < dim2; ++jj) {
                ijk[dj] = 0.5 + jj;
                double pos0[3];
                for (int d = 0; d < 3; ++d) pos0[d] = outer.lowCorners[dd][d] + ijk[d] * outer.cellSize;   // :1035-1036
                for (int ss = 0; ss < S; ++ss) {
                    const FFATMap& sh = F.shells[ss];
                    double pos[3]; int posind[3];
                    ffat_intersect(sh, pos0, pos, posind);           // :1043
                    const double dx = pos[0] - F.center3[0], dy = pos[1] - F.center3[1], dz = pos[2] - F.center3[2];
                    const size_t row = (size_t)(offset + ii * dim2 + jj) * S + ss;
                    R[row] = std::sqrt(dx * dx + dy * dy + dz * dz); // :1046
                    int idx[4][3]; double co[4];
                    ffat_interpolate(sh, pos, posind, idx, co);      // :1051
                    double pre = 0, pim = 0;
                    for (int kk = 0; kk < 4; ++kk) {                 // :1052-1057
                        const int q = sh.strides[idx[kk][0]] + idx[kk][1] * sh.nElem[idx[kk][0]][1] + idx[kk][2];
                        const size_t e = (size_t)2 * F.strides[ss] + (size_t)2 * q;
                        pre += co[kk] * pressure[2 * e]; pim += co[kk] * pressure[2 * e + 1];
                    }
                    Pre[row] = pre; Pim[row] = pim;
                }
            }
        }
        offset += dim1 * dim2;
    }
    // FFAT_Solver<T,3>::Solve: per direction, least squares of basis * psi = |p| with basis_s = 1/(k r_s)
    // (:881-895).  Eigen's JacobiSVD of an S x 1 matrix: sigma = |basis|, u = basis/sigma, v = 1, so
    // solve(b) = (u . b) / sigma.
    for (int ii = 0; ii < F.n_dir; ++ii) {
        double ss2 = 0, ub = 0;
        for (int s = 0; s < S; ++s) {
            const double kr = R[(size_t)ii * S + s] * k;             // :882
            const double basis = 1.0 / kr;                           // :883 (pow(kr,1) == kr)
            const double p2 = std::hypot(Pre[(size_t)ii * S + s], Pim[(size_t)ii * S + s]);   // :885 std::abs(complex)
            ss2 += basis * basis; ub += basis * p2;
            if (Pabs_out) Pabs_out[(size_t)ii * S + s] = p2;
        }
        const double sigma = std::sqrt(ss2);
        Psi[ii] = (ub / sigma) / sigma;
    }
    if (R_out) std::memcpy(R_out, R.data(), sizeof(double) * R.size());
    double scale = 1.0;
    if (powerScaling) {                                              // :909-929 (shell 0 only)
        double numer = 0, denom = 0;
        for (int ii = 0; ii < F.n_dir; ++ii) {
            const double kr = k * R[(size_t)ii * S];
            const double pa = std::hypot(Pre[(size_t)ii * S], Pim[(size_t)ii * S]);
            numer += std::pow(pa, 2);
            denom += std::pow(Psi[ii] / kr, 2);
        }
        scale = std::sqrt(numer / denom);
        for (int ii = 0; ii < F.n_dir; ++ii) Psi[ii] *= scale;
    }
    return scale;
}

// -----------------------------------------------------------------------------
// ModeData / ModalMaterial helpers   [pinned:_ref]
// -----------------------------------------------------------------------------
// ModeData.h:120-148 numModesAudible incl. its cache quirk (the cache is only
// filled by the loop branch, :143-147)
struct AudibleCache { int N = -1; double freqThres = 22100., density = -1; };
static int num_modes_audible(const double* omega2, int n, double density, double audibleFreq,
                             AudibleCache& c) {
    if (density == c.density && c.freqThres == audibleFreq && c.N >= 0) return c.N;  // :123-127
    auto Freq = [&](double os) { return std::sqrt(os / density) / (2. * M_PI); };     // :128-130
    if (n == 0 || Freq(omega2[0]) > audibleFreq) return 0;                            // :131-133
    if (Freq(omega2[n - 1]) <= audibleFreq) return n;                                 // :134-136
    int ii;
    for (ii = 0; ii < n; ++ii) if (Freq(omega2[ii]) > audibleFreq) break;             // :137-142
    c.N = ii; c.density = density; c.freqThres = audibleFreq;                         // :143-145
    return c.N;
}

}  // namespace orc

// =============================================================================
// C entry points (ctypes) -- thin, no logic
// =============================================================================
using namespace orc;
extern "C" {

void orc_build_ab(double density, const double* omega2, int N, double alpha, double beta,
                  double* a, double* b) { build_ab(density, omega2, N, alpha, beta, a, b); }
void orc_coeffs(int N, double h, const double* a, const double* b, double* c1, double* c2,
                double* c3) { coeffs(N, h, a, b, c1, c2, c3); }

void* orc_integrator_create(int N, double h, const double* a, const double* b) {
    return new std::shared_ptr<Integrator>(new Integrator(N, h, a, b));
}
void orc_integrator_destroy(void* p) { delete static_cast<std::shared_ptr<Integrator>*>(p); }
void orc_integrator_step(void* p, const double* Q, double* q_out) {
    Integrator& I = **static_cast<std::shared_ptr<Integrator>*>(p);
    const double* q = I.step(Q);
    if (q_out) std::memcpy(q_out, q, sizeof(double) * I.N);
}
// direct-form state as (q_{k-1}, q_{k-2})
void orc_integrator_get_state(void* p, double* q1, double* q2) {
    Integrator& I = **static_cast<std::shared_ptr<Integrator>*>(p);
    std::memcpy(q1, I.q[I.ptr % 3].data(), sizeof(double) * I.N);
    std::memcpy(q2, I.q[(I.ptr + 2) % 3].data(), sizeof(double) * I.N);
}

void* orc_solver_create(int N, int BUF, void* integ) {
    Solver* s = new Solver(N, BUF);
    s->integ = *static_cast<std::shared_ptr<Integrator>*>(integ);
    return s;
}
void orc_solver_destroy(void* s) { delete static_cast<Solver*>(s); }
// type: 0 point, 1 gaussian(width_us), 2 autoregressive.  flags: bit0 sustainedStart,
// bit1 sustainedEnd, bit2 clearAllForces.  Returns 1 on success, 0 if the queue is full.
int orc_solver_enqueue_force(void* sp, const double* data, int type, double width_us, int flags) {
    Solver* s = static_cast<Solver*>(sp);
    if (s->q_force.size() >= s->cap_force) return 0;
    ForceMessage m;
    m.data.assign(data, data + s->N);
    m.forceType = type;
    if (type == GaussianForceT) m.force.reset(new GaussianForce(width_us));
    else if (type == AutoregressiveForceT) m.force.reset(new AutoregressiveForce());
    else m.force.reset(new PointForce());
    m.sustainedForceStart = flags & 1; m.sustainedForceEnd = flags & 2; m.clearAllForces = flags & 4;
    s->q_force.push_back(m);
    return 1;
}
int orc_solver_enqueue_trans(void* sp, const double* data, int n) {
    Solver* s = static_cast<Solver*>(sp);
    if (s->q_trans.size() >= s->cap_trans) return 0;
    s->q_trans.emplace_back(data, data + n);
    return 1;
}
int orc_solver_enqueue_arprm(void* sp, double a0, double a1, double sigma, double mu) {
    Solver* s = static_cast<Solver*>(sp);
    if (s->q_arprm.size() >= s->cap_arprm) return 0;
    s->q_arprm.push_back({a0, a1, sigma, mu});
    return 1;
}
void orc_solver_set_use_transfer(void* sp, int use) { static_cast<Solver*>(sp)->useTransfer = use; }
int orc_solver_step(void* sp, double* sound, double* qnorm) {
    Solver* s = static_cast<Solver*>(sp);
    int produced = s->step();
    if (produced) {
        if (sound) std::memcpy(sound, s->sound.data(), sizeof(double) * s->BUF);
        if (qnorm) std::memcpy(qnorm, s->qnorm.data(), sizeof(double) * s->N);
    }
    return produced;
}
// (space, time) the force state machine handed to the hot loop in the last step (modal_solver.h:206-240)
void orc_solver_last_force(void* sp, double* space, double* time) {
    Solver* s = static_cast<Solver*>(sp);
    std::memcpy(space, s->space.data(), sizeof(double) * s->N);
    std::memcpy(time, s->time_.data(), sizeof(double) * s->BUF);
}
void orc_solver_latest_transfer(void* sp, double* trans) {
    Solver* s = static_cast<Solver*>(sp);
    std::memcpy(trans, s->latest_transfer.data(), sizeof(double) * s->latest_transfer.size());
}
int orc_solver_num_active(void* sp) { return (int)static_cast<Solver*>(sp)->active.size(); }

// Temporal force profiles on their own (forces.h:81-128): fills n_buf buffers of BUF
// samples; alive[b] = Add's return value for buffer b.
void orc_force_profile(int type, double width_us, int BUF, int n_buf, double* out, int* alive) {
    std::unique_ptr<Force> f;
    if (type == GaussianForceT) f.reset(new GaussianForce(width_us));
    else if (type == AutoregressiveForceT) f.reset(new AutoregressiveForce());
    else f.reset(new PointForce());
    for (int b = 0; b < n_buf; ++b) {
        double* o = out + (size_t)b * BUF;
        std::fill(o, o + BUF, 0.0);
        alive[b] = f->Add(o, BUF) ? 1 : 0;
    }
}

void orc_project_vertex(int forceDim, const double* U, int nDOF, int vid, const double* vn,
                        double* out) { project_vertex(forceDim, U, nDOF, vid, vn, out); }
void orc_project_face(int forceDim, const double* U, int nDOF, const int* vids,
                      const double* coords, const double* vn, double* out) {
    project_face(forceDim, U, nDOF, vids, coords, vn, out);
}
// Dense restatement used only for the batched-projection parity check: Y[b][m] = sum_d U[m][d] F[b][d]
// (F is B x K, one dense load vector per impulse).  Same contraction as GetModalForceVertex with a dense f.
void orc_project_dense(int M, int K, int B, const double* U, const double* F, double* Y) {
    for (int b = 0; b < B; ++b) {
        const double* f = F + (size_t)b * K;
        for (int m = 0; m < M; ++m) {
            const double* u = U + (size_t)m * K;
            double acc = 0.0;
            for (int d = 0; d < K; ++d) acc += u[d] * f[d];
            Y[(size_t)b * M + m] = acc;
        }
    }
}

// geom: 1 + 18 + 3 + 3 + 3 + 3 + 1 doubles = {cellSize, lowCorners[6][3], center1, bboxLow,
// bboxTop, center3, k}; igeom: nElem[6][2], strides[6] = 18 ints.
static void fill_map(FFATMap& M, const double* geom, const int* igeom, const double* Psi, int D) {
    M.cellSize = geom[0];
    std::memcpy(M.lowCorners, geom + 1, sizeof(double) * 18);
    std::memcpy(M.center1, geom + 19, sizeof(double) * 3);
    std::memcpy(M.bboxLow, geom + 22, sizeof(double) * 3);
    std::memcpy(M.bboxTop, geom + 25, sizeof(double) * 3);
    std::memcpy(M.center3, geom + 28, sizeof(double) * 3);
    M.k = geom[31];
    std::memcpy(M.nElem, igeom, sizeof(int) * 12);
    std::memcpy(M.strides, igeom + 12, sizeof(int) * 6);
    M.Psi = Psi; M.D = D;
}
// modal_solver.h:286-315 computeTransfer: out[m + l*n_maps] = |GetMapVal_m(pos_l)| (col-major N x L,
// tools/real_time_modal_sound.cpp:921-927).  geom/igeom/Psi are per-map contiguous blocks.
void orc_ffat_eval(int n_maps, const double* geom, const int* igeom, const double* Psi, int D,
                   const double* pos, int L, double* out) {
    for (int l = 0; l < L; ++l)
        for (int m = 0; m < n_maps; ++m) {
            FFATMap M;
            fill_map(M, geom + (size_t)m * 32, igeom + (size_t)m * 18, Psi + (size_t)m * D, D);
            out[(size_t)l * n_maps + m] = std::abs(ffat_getmapval(M, pos + 3 * l));
        }
}
void orc_ffat_intersect(const double* geom, const int* igeom, const double* p, double* surf,
                        int* ind) {
    FFATMap M; fill_map(M, geom, igeom, nullptr, 0);
    ffat_intersect(M, p, surf, ind);
}
void orc_ffat_interpolate(const double* geom, const int* igeom, const double* surf, const int* nn,
                          int* idx12, double* co4) {
    FFATMap M; fill_map(M, geom, igeom, nullptr, 0);
    int idx[4][3];
    ffat_interpolate(M, surf, nn, idx, co4);
    std::memcpy(idx12, idx, sizeof(idx));
}

// FFAT map construction.  geom_out [S][32] / igeom_out [S][18] use the record layout of fill_map (k = -1,
// centre3 = shell 2's centre); shell_strides [S]; counts[2] = {N_elements_total, N_directions}.
void orc_ffat_fit_geometry(double cellSize, const double* V, const int* nElem, int S, double bbox_init,
                           double* geom_out, int* igeom_out, int* shell_strides, int* counts) {
    FFATFit F; ffat_fit_geometry(cellSize, V, nElem, S, bbox_init, F);
    for (int s = 0; s < S; ++s) {
        const FFATMap& M = F.shells[s];
        double* g = geom_out + (size_t)s * 32; int* ig = igeom_out + (size_t)s * 18;
        g[0] = M.cellSize; std::memcpy(g + 1, M.lowCorners, sizeof(double) * 18);
        std::memcpy(g + 19, M.center1, 24); std::memcpy(g + 22, M.bboxLow, 24); std::memcpy(g + 25, M.bboxTop, 24);
        std::memcpy(g + 28, F.center3, 24); g[31] = -1.0;
        std::memcpy(ig, M.nElem, sizeof(int) * 12); std::memcpy(ig + 12, M.strides, sizeof(int) * 6);
        shell_strides[s] = F.strides[s];
    }
    counts[0] = F.n_total; counts[1] = F.n_dir;
}
// Solve for n_maps modes sharing the geometry: k[n_maps], pressure [n_maps][2*n_total] complex interleaved,
// psi_out [n_maps][n_dir], scale_out [n_maps] or NULL, R_out [n_dir][S] or NULL (same for every map).
void orc_ffat_fit_solve(int S, const double* geom, const int* igeom, const int* shell_strides, int n_maps,
                        const double* k, const double* pressure, int power_scaling, double* psi_out,
                        double* scale_out, double* R_out, double* Pabs_out) {
    FFATFit F; F.shells.resize(S); F.strides.assign(shell_strides, shell_strides + S);
    for (int s = 0; s < S; ++s) fill_map(F.shells[s], geom + (size_t)s * 32, igeom + (size_t)s * 18, nullptr, 0);
    std::memcpy(F.center3, geom + 2 * 32 + 28, 24);
    F.n_dir = 0; F.n_total = 0;
    for (int f = 0; f < 6; ++f) F.n_dir += F.shells[2].nElem[f][0] * F.shells[2].nElem[f][1];
    for (int s = 0; s < S; ++s) for (int f = 0; f < 6; ++f) F.n_total += F.shells[s].nElem[f][0] * F.shells[s].nElem[f][1];
    for (int m = 0; m < n_maps; ++m) {
        const double sc = ffat_fit_solve(F, k[m], pressure + (size_t)m * 4 * F.n_total, power_scaling != 0,
                                         psi_out + (size_t)m * F.n_dir, m == 0 ? R_out : nullptr,
                                         Pabs_out ? Pabs_out + (size_t)m * F.n_dir * S : nullptr);
        if (scale_out) scale_out[m] = sc;
    }
}

int orc_num_modes_audible(const double* omega2, int n, double density, double freq, double* cache3) {
    AudibleCache c; c.N = (int)cache3[0]; c.freqThres = cache3[1]; c.density = cache3[2];
    int r = num_modes_audible(omega2, n, density, freq, c);
    cache3[0] = c.N; cache3[1] = c.freqThres; cache3[2] = c.density;
    return r;
}

// ModalMaterial.h:35-55 Read: skip leading '#' lines, then "density E nu alpha beta".
// out5 = {density, youngsModulus, poissonRatio, alpha, beta}; returns 0 if the file is missing.
int orc_material_read(const char* filename, double* out5) {
    std::ifstream stream(filename);
    if (!stream) return 0;
    std::string line;
    while (std::getline(stream, line)) { if (line[0] != '#') break; }
    std::istringstream iss(line);
    double v[5] = {0, 0, 0, 0, 0};
    iss >> v[0]; iss >> v[1]; iss >> v[2]; iss >> v[3]; iss >> v[4];
    std::memcpy(out5, v, sizeof(v));
    return 1;
}
// ModeData.h:61-83 read: int nDOF, int nModes, double w2[nModes], double U[nModes][nDOF]
int orc_modes_read_header(const char* filename, int* nDOF, int* nModes) {
    std::ifstream fin(filename, std::ios::binary);
    if (!fin.good()) return 0;
    fin.read((char*)nDOF, sizeof(int));
    fin.read((char*)nModes, sizeof(int));
    return fin.good() ? 1 : 0;
}
int orc_modes_read(const char* filename, double* omega2, double* U) {
    std::ifstream fin(filename, std::ios::binary);
    if (!fin.good()) return 0;
    int nDOF, nModes;
    fin.read((char*)&nDOF, sizeof(int));
    fin.read((char*)&nModes, sizeof(int));
    fin.read((char*)omega2, sizeof(double) * nModes);
    for (int i = 0; i < nModes; ++i) fin.read((char*)(U + (size_t)i * nDOF), sizeof(double) * nDOF);
    return fin.good() ? 1 : 0;
}
// ModeData.h:87-107 write
int orc_modes_write(const char* filename, int nDOF, int nModes, const double* omega2,
                    const double* U) {
    std::ofstream fout(filename, std::ios::binary);
    if (!fout.good()) return 0;
    fout.write((const char*)&nDOF, sizeof(int));
    fout.write((const char*)&nModes, sizeof(int));
    fout.write((const char*)omega2, sizeof(double) * nModes);
    for (int i = 0; i < nModes; ++i)
        fout.write((const char*)(U + (size_t)i * nDOF), sizeof(double) * nDOF);
    return 1;
}

// -----------------------------------------------------------------------------
// Whole-job CPU baseline helper: renders n_obj independent objects for n_buf buffers each with
// the reference loop (Solver::step), one PointForce per object at buffer imp_buf[o] with modal
// load space[o][N] and transfer trans[o][N]; mixes into mix[n_buf*BUF] (+=).  Used by bench.py's
// cpu_baseline / --impl reference legs (threads call it on disjoint object ranges).
// -----------------------------------------------------------------------------
void orc_batch_render(int n_obj, int N, int BUF, int n_buf, double h, const double* a,
                      const double* b, const double* space, const double* trans,
                      const int* imp_buf, double* mix) {
    std::vector<double> y(BUF);
    for (int o = 0; o < n_obj; ++o) {
        auto integ = std::make_shared<Integrator>(N, h, a + (size_t)o * N, b + (size_t)o * N);
        Solver s(N, BUF);
        s.integ = integ;
        s.q_trans.emplace_back(trans + (size_t)o * N, trans + (size_t)(o + 1) * N);
        for (int bi = 0; bi < n_buf; ++bi) {
            if (bi == imp_buf[o]) {
                ForceMessage m; m.data.assign(space + (size_t)o * N, space + (size_t)(o + 1) * N);
                s.q_force.push_back(m);
            }
            s.step();
            double* out = mix + (size_t)bi * BUF;
            for (int i = 0; i < BUF; ++i) out[i] += s.sound[i];
        }
    }
}

}  // extern "C"
