// ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the REFERENCE's own headers (read in place
// from /root/reference via -I, never copied) against the Eigen shim and exposes them to ctypes so
// that tests/test_oracle_vs_ref.py can pin oracle/pbso_oracle.cpp to them.  Built by oracle/Makefile
// into oracle/_ref/libpbso_ref.so (git-ignored).  The reference's arithmetic statements are its own;
// only the Eigen container underneath (element-wise loops, a dot product) is the shim's, and the
// protobuf message classes under FFAT_Map_Serialize are oracle/ref_stubs/ffat_map.pb.h.
// io.cpp (ListDirFiles) is compiled alongside, also in place.
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>
#include "Eigen/Dense"
#include "config.h"            // /root/reference/config.h
#include "modal_integrator.h"  // /root/reference/modal_integrator.h
#include "forces.h"            // /root/reference/forces.h
#include "ModeData.h"          // /root/reference/ModeData.h
#include "ModalMaterial.h"     // /root/reference/ModalMaterial.h
// modal_solver.h pulls in ffat_solver.h and ffat_map_serialize.h; their absent third-party includes
// (libigl viewer/serialize, protoc-generated ffat_map.pb.h) are satisfied by oracle/ref_stubs/.
#include "modal_solver.h"      // /root/reference/modal_solver.h

typedef ModalIntegrator<double> Integ;
typedef Integ::ModalVec Vec;

template <int BUF>
static void force_profile(int type, double width_us, int n_buf, double* out, int* alive) {
    std::unique_ptr<Force<double, BUF>> f;
    if (type == 1) f.reset(new GaussianForce<double, BUF>(width_us));
    else if (type == 2) f.reset(new AutoregressiveForce<double, BUF>());
    else f.reset(new PointForce<double, BUF>());
    Eigen::Matrix<double, BUF, 1> spread;
    for (int b = 0; b < n_buf; ++b) {
        spread.setZero();
        alive[b] = f->Add(spread) ? 1 : 0;
        std::memcpy(out + (size_t)b * BUF, spread.data(), sizeof(double) * BUF);
    }
}

extern "C" {
void* ref_integrator_build(double density, const double* omega2, int n_omega, double alpha,
                           double beta, double h, int N) {
    std::vector<double> os(omega2, omega2 + n_omega);
    return Integ::Build(density, os, alpha, beta, h, N);
}
void* ref_integrator_create(int N, double h, const double* a, const double* b) {
    Vec va, vb; va.resize(N); vb.resize(N);
    for (int i = 0; i < N; ++i) { va(i) = a[i]; vb(i) = b[i]; }
    return new Integ(N, h, va, vb);
}
void ref_integrator_destroy(void* p) { delete static_cast<Integ*>(p); }
void ref_integrator_step(void* p, int N, const double* Q, double* q_out) {
    Integ* I = static_cast<Integ*>(p);
    if (Q) {
        Vec vq; vq.resize(N);
        for (int i = 0; i < N; ++i) vq(i) = Q[i];
        const Vec& q = I->Step(vq);
        std::memcpy(q_out, q.data(), sizeof(double) * N);
    } else {
        const Vec& q = I->Step();
        std::memcpy(q_out, q.data(), sizeof(double) * N);
    }
}
// BUF is a template parameter in the reference: instantiate the sizes the tests use.
int ref_force_profile(int type, double width_us, int BUF, int n_buf, double* out, int* alive) {
    switch (BUF) {
        case 64:  force_profile<64>(type, width_us, n_buf, out, alive); return 1;
        case 256: force_profile<256>(type, width_us, n_buf, out, alive); return 1;
        case 513: force_profile<513>(type, width_us, n_buf, out, alive); return 1;
        default: return 0;
    }
}
int ref_num_modes_audible(const double* omega2, int n, double density, double freq, int repeat,
                          int* results) {
    ModeData<double> md;
    md._omegaSquared.assign(omega2, omega2 + n);
    for (int r = 0; r < repeat; ++r) results[r] = md.numModesAudible(density, freq);
    return md._N_modesAudible;
}
int ref_modes_roundtrip(const char* in_file, const char* out_file, int* nDOF, int* nModes,
                        double* omega2_first_last) {
    ModeData<double> md;
    md.read(in_file);
    *nDOF = md.numDOF(); *nModes = md.numModes();
    omega2_first_last[0] = md.omegaSquared(0);
    omega2_first_last[1] = md.omegaSquared(md.numModes() - 1);
    md.write(out_file);
    return 1;
}
int ref_material_read(const char* filename, double* out5) {
    std::unique_ptr<ModalMaterial<double>> m(ModalMaterial<double>::Read(filename));
    if (!m) return 0;
    out5[0] = m->density; out5[1] = m->youngsModulus; out5[2] = m->poissonRatio;
    out5[3] = m->alpha; out5[4] = m->beta;
    return 1;
}
double ref_material_xi(double alpha, double beta, double omega) {
    ModalMaterial<double> m; m.alpha = alpha; m.beta = beta; return m.xi(omega);
}
double ref_material_omega_di(double alpha, double beta, double omega) {
    ModalMaterial<double> m; m.alpha = alpha; m.beta = beta; return m.omega_di(omega);
}
}

// ---------------------------------------------------------------------------------------------
// ModalSolver<double,BUF>::step and its queues (modal_solver.h:100-399), FFAT runtime query
// (ffat_solver.h:676-803, 1180-1206) and the .fatcube loader (ffat_map_serialize.h:90-279),
// all the reference's own code.
namespace {
struct SolverBase {
    virtual ~SolverBase() {}
    virtual int enqueue_force(const double* d, int type, double width_us, int flags) = 0;
    virtual int enqueue_trans(const double* d, int n) = 0;
    virtual int enqueue_arprm(double a0, double a1, double sigma, double mu) = 0;
    virtual void set_use_transfer(int u) = 0;
    virtual int step(double* y, double* qnorm) = 0;
    virtual void latest_transfer(double* out) = 0;
    virtual void read_maps(const char* dir) = 0;
    virtual int compute_transfer(const double* pos, double* out) = 0;
    virtual int compute_transfer_enqueue(const double* pos) = 0;
};
template <int BUF>
struct SolverT : SolverBase {
    int N;
    ModalSolver<double, BUF> s;
    SolverT(int N_, std::shared_ptr<Integ> integ) : N(N_), s(N_) { s.setIntegrator(integ); }
    int enqueue_force(const double* d, int type, double width_us, int flags) override {
        ForceMessage<double, BUF> m;
        m.data.resize(N);
        for (int i = 0; i < N; ++i) m.data(i) = d[i];
        if (type == 1) { m.forceType = ForceType::GaussianForce; m.force.reset(new GaussianForce<double, BUF>(width_us)); }
        else if (type == 2) { m.forceType = ForceType::AutoregressiveForce; m.force.reset(new AutoregressiveForce<double, BUF>()); }
        else { m.forceType = ForceType::PointForce; m.force.reset(new PointForce<double, BUF>()); }
        m.sustainedForceStart = flags & 1; m.sustainedForceEnd = flags & 2; m.clearAllForces = flags & 4;
        return s.enqueueForceMessage(m) ? 1 : 0;
    }
    int enqueue_trans(const double* d, int n) override {
        TransMessage<double> t; t.N = n; t.data.resize(n);
        for (int i = 0; i < n; ++i) t.data(i) = d[i];
        return s.enqueueTransMessage(t) ? 1 : 0;
    }
    int enqueue_arprm(double a0, double a1, double sigma, double mu) override {
        AutoregressiveForceParam<double> p; p.a[0] = a0; p.a[1] = a1; p.sigma = sigma; p.mu = mu;
        return s.enqueueArprmMessage(p) ? 1 : 0;
    }
    void set_use_transfer(int u) override { s.setUseTransfer(u != 0); }
    int step(double* y, double* qnorm) override {
        s.step();
        SoundMessage<double, BUF> snd;
        if (!s.dequeueSoundMessage(snd)) return 0;      // clearAllForces returns before rendering
        std::memcpy(y, snd.data.data(), sizeof(double) * BUF);
        Eigen::Matrix<double, -1, 1> q = s.getQBufferNorm();
        if (qnorm) std::memcpy(qnorm, q.data(), sizeof(double) * N);
        return 1;
    }
    void latest_transfer(double* out) override {
        const TransMessage<double>& t = s.getLatestTransfer();
        std::memcpy(out, t.data.data(), sizeof(double) * t.data.size());
    }
    void read_maps(const char* dir) override { s.readFFATMaps(dir); }
    int compute_transfer(const double* pos, double* out) override {
        return s.computeTransfer(Eigen::Matrix<double, 3, 1>(pos[0], pos[1], pos[2]), out) ? 1 : 0;
    }
    int compute_transfer_enqueue(const double* pos) override {
        return s.computeTransfer(Eigen::Matrix<double, 3, 1>(pos[0], pos[1], pos[2])) ? 1 : 0;
    }
};
typedef std::map<int, Gpu_Wavesolver::FFAT_Map<double, 3>> MapSet;
}  // namespace

extern "C" {
// The solver shares ownership of the integrator the way tools/real_time_modal_sound.cpp:332-345 does;
// here the bridge keeps the raw pointer alive (no-op deleter), the caller destroys it separately.
void* ref_solver_create(int N, int BUF, void* integ) {
    std::shared_ptr<Integ> sp(static_cast<Integ*>(integ), [](Integ*) {});
    switch (BUF) {
        case 64:  return static_cast<SolverBase*>(new SolverT<64>(N, sp));
        case 256: return static_cast<SolverBase*>(new SolverT<256>(N, sp));
        case 513: return static_cast<SolverBase*>(new SolverT<513>(N, sp));
        default: return nullptr;
    }
}
void ref_solver_destroy(void* p) { delete static_cast<SolverBase*>(p); }
int ref_solver_enqueue_force(void* p, const double* d, int type, double width_us, int flags) { return static_cast<SolverBase*>(p)->enqueue_force(d, type, width_us, flags); }
int ref_solver_enqueue_trans(void* p, const double* d, int n) { return static_cast<SolverBase*>(p)->enqueue_trans(d, n); }
int ref_solver_enqueue_arprm(void* p, double a0, double a1, double sigma, double mu) { return static_cast<SolverBase*>(p)->enqueue_arprm(a0, a1, sigma, mu); }
void ref_solver_set_use_transfer(void* p, int u) { static_cast<SolverBase*>(p)->set_use_transfer(u); }
int ref_solver_step(void* p, double* y, double* qnorm) { return static_cast<SolverBase*>(p)->step(y, qnorm); }
void ref_solver_latest_transfer(void* p, double* out) { static_cast<SolverBase*>(p)->latest_transfer(out); }
void ref_solver_read_ffat_maps(void* p, const char* dir) { static_cast<SolverBase*>(p)->read_maps(dir); }
int ref_solver_compute_transfer(void* p, const double* pos, double* out) { return static_cast<SolverBase*>(p)->compute_transfer(pos, out); }
int ref_solver_compute_transfer_enqueue(void* p, const double* pos) { return static_cast<SolverBase*>(p)->compute_transfer_enqueue(pos); }

// FFAT_Map_Serialize::LoadAll (ffat_map_serialize.h:268-279) -> handle; returns the number of maps.
void* ref_ffat_load_all(const char* dir, int* n_maps) {
    MapSet* m = Gpu_Wavesolver::FFAT_Map_Serialize::LoadAll(dir);
    *n_maps = (int)m->size();
    return m;
}
void ref_ffat_free(void* h) { delete static_cast<MapSet*>(h); }
// out[l*n + i] = |maps.at(i).GetMapVal(pos_l)|  (modal_solver.h:286-315); returns 0 on a missing modeId.
int ref_ffat_eval(void* h, const double* pos, int L, int use_compressed, double* out) {
    MapSet& m = *static_cast<MapSet*>(h);
    const int n = (int)m.size();
    try {
        for (int l = 0; l < L; ++l) {
            Eigen::Matrix<double, 3, 1> p(pos[3 * l], pos[3 * l + 1], pos[3 * l + 2]);
            for (int i = 0; i < n; ++i) out[(size_t)l * n + i] = std::abs(m.at(i).GetMapVal(p, use_compressed != 0));
        }
    } catch (const std::out_of_range&) { return 0; }
    return 1;
}
// The LEGACY .fatcube format (libigl's igl::serialize of the FFAT_Map<T,3> object, ffat_solver.h:978-991, 1066-1085), with the
// reference's own code and libigl's own serialize.h compiled in place:
//   ref_ffat_legacy_from_fatcube  FFAT_Map_Serialize::Load(protobuf file) -> FFAT_Map<T,3>::Save(legacy file)
//   ref_ffat_legacy_fit_save      constructor + Solve (as ref_ffat_fit) -> FFAT_Map<T,3>::Save(legacy file): all three shells inside
//   ref_ffat_legacy_load_all      FFAT_Map<T,3>::LoadAll(dir) (the legacy loader) -> the same map set handle ref_ffat_eval takes
int ref_ffat_legacy_from_fatcube(const char* in_file, const char* out_file) {
    Gpu_Wavesolver::FFAT_Map<double, 3> map;
    Gpu_Wavesolver::FFAT_Map_Serialize::Load(in_file, map);
    Gpu_Wavesolver::FFAT_Map<double, 3>::Save(out_file, map);
    return map.modeId;
}
int ref_ffat_legacy_fit_save(int mode_id, double cell_size, const double* V, int n_rows, const int* n_elements, int n_shells,
                             double k, const double* pressure, int power_scaling, const char* legacy_file) {
    typedef Gpu_Wavesolver::FFAT_Map<double, 3> Map3;
    Eigen::Matrix<double, Eigen::Dynamic, 3> Vm; Vm.resize(n_rows, 3);
    for (int i = 0; i < n_rows; ++i) for (int j = 0; j < 3; ++j) Vm(i, j) = V[(size_t)i * 3 + j];
    std::vector<std::vector<std::pair<int, int>>> ne(n_shells, std::vector<std::pair<int, int>>(6));
    int n_total = 0;
    for (int s = 0; s < n_shells; ++s) for (int f = 0; f < 6; ++f) {
        ne[s][f] = std::make_pair(n_elements[(s * 6 + f) * 2], n_elements[(s * 6 + f) * 2 + 1]);
        n_total += ne[s][f].first * ne[s][f].second;
    }
    Map3 map(mode_id, cell_size, Vm, ne);
    Map3::FFAT_VectorXcd P; P.resize(2 * n_total);
    for (int i = 0; i < 2 * n_total; ++i) P(i) = std::complex<double>(pressure[2 * i], pressure[2 * i + 1]);
    map.Solve(k, P, power_scaling != 0);
    Map3::Save(legacy_file, map);
    return (int)map.GetData().rows();
}
void* ref_ffat_legacy_load_all(const char* dir, int* n_maps) {
    MapSet* m = Gpu_Wavesolver::FFAT_Map<double, 3>::LoadAll(dir);
    *n_maps = (int)m->size();
    return m;
}
// Load one file and Save it again with the reference's own code (byte-level round trip of the codec).
int ref_ffat_load_save(const char* in_file, const char* out_file, int* mode_id, double* k_cell) {
    Gpu_Wavesolver::FFAT_Map<double, 3> map;
    Gpu_Wavesolver::FFAT_Map_Serialize::Load(in_file, map);
    *mode_id = map.modeId;
    k_cell[0] = map.GetData().size() ? map.GetData()(0, 0) : 0.0;
    k_cell[1] = (double)map.GetData().rows();
    k_cell[2] = (double)map.GetData().cols();
    Gpu_Wavesolver::FFAT_Map_Serialize::Save(out_file, map);
    return 1;
}
// Gpu_Wavesolver::ListDirFiles (io.cpp:18-35): count of entries with the extension (sorted list length).
int ref_list_dir_files(const char* dir, const char* ext, char* joined, int cap) {
    std::vector<std::string> names;
    Gpu_Wavesolver::ListDirFiles(dir, names, ext);
    std::string j;
    for (auto& s : names) { j += s; j += '\n'; }
    if (joined && cap > 0) { std::strncpy(joined, j.c_str(), cap - 1); joined[cap - 1] = 0; }
    return (int)names.size();
}

// cfg5-style offline batch with the reference's own classes: one ModalSolver<double,256> +
// ModalIntegrator per object, a single PointForce at buffer imp_buf[o], sound dequeued after every
// step and mixed down (the loop tools/real_time_modal_sound.cpp:527-535 runs on its sim thread).
void ref_batch_render(int n_obj, int N, int n_buf, double h, const double* a, const double* b,
                      const double* space, const double* trans, const int* imp_buf, double* mix) {
    const int BUF = 256;
    for (int o = 0; o < n_obj; ++o) {
        Vec va, vb; va.resize(N); vb.resize(N);
        for (int i = 0; i < N; ++i) { va(i) = a[(size_t)o * N + i]; vb(i) = b[(size_t)o * N + i]; }
        std::shared_ptr<Integ> integ(new Integ(N, h, va, vb));
        SolverT<BUF> s(N, integ);
        s.enqueue_trans(trans + (size_t)o * N, N);
        std::vector<double> y(BUF);
        for (int bi = 0; bi < n_buf; ++bi) {
            if (bi == imp_buf[o]) s.enqueue_force(space + (size_t)o * N, 0, 0.0, 0);
            if (s.step(y.data(), nullptr)) {
                double* out = mix + (size_t)bi * BUF;
                for (int i = 0; i < BUF; ++i) out[i] += y[i];
            }
        }
    }
}

// FFAT map construction: the reference's own FFAT_Map<double,3>(modeId, cellSize, V, N_elements) constructor
// (ffat_solver.h:944-989 -> shells :399-428), Solve (:1007-1069 -> FFAT_Solver<T,3>::Solve :872-897,
// Scaling :909-929) and, optionally, FFAT_Map_Serialize::Save of the result.  V arrives row-major [rows][3].
// The JacobiSVD underneath is the shim's (closed form for a one-column matrix).
int ref_ffat_fit(int mode_id, double cell_size, const double* V, int n_rows, const int* n_elements, int n_shells,
                 double k, const double* pressure, int power_scaling, double* psi_out, double* centre_out,
                 const char* save_to) {
    typedef Gpu_Wavesolver::FFAT_Map<double, 3> Map3;
    Eigen::Matrix<double, Eigen::Dynamic, 3> Vm; Vm.resize(n_rows, 3);
    for (int i = 0; i < n_rows; ++i) for (int j = 0; j < 3; ++j) Vm(i, j) = V[(size_t)i * 3 + j];
    std::vector<std::vector<std::pair<int, int>>> ne(n_shells, std::vector<std::pair<int, int>>(6));
    int n_total = 0;
    for (int s = 0; s < n_shells; ++s) for (int f = 0; f < 6; ++f) {
        ne[s][f] = std::make_pair(n_elements[(s * 6 + f) * 2], n_elements[(s * 6 + f) * 2 + 1]);
        n_total += ne[s][f].first * ne[s][f].second;
    }
    Map3 map(mode_id, cell_size, Vm, ne);
    Map3::FFAT_VectorXcd P; P.resize(2 * n_total);
    for (int i = 0; i < 2 * n_total; ++i) P(i) = std::complex<double>(pressure[2 * i], pressure[2 * i + 1]);
    map.Solve(k, P, power_scaling != 0);
    const auto& Psi = map.GetData();
    for (int i = 0; i < (int)Psi.rows(); ++i) psi_out[i] = Psi(i, 0);
    const auto c = map.GetCenter();
    for (int j = 0; j < 3; ++j) centre_out[j] = c(j);
    if (save_to) Gpu_Wavesolver::FFAT_Map_Serialize::Save(save_to, map);
    return (int)Psi.rows();
}

// The reference's own producer of the cube-map mesh: FFAT_Map<double,1>::CubemapMesh (ffat_solver.h:334-397).  Cells
// bboxLow_r..bboxTop_r (inclusive) of a grid with the given low corner and cell size; returns the number of vertices
// written (4 per quad, V_out row-major) and N_elements [6][2].
int ref_cubemap_mesh(const int* bbox_low_r, const int* bbox_top_r, double cell_size, const double* grid_low, const int* dim,
                     double* V_out, int max_rows, int* n_elements_out, int* data_indices_out) {
    typedef Gpu_Wavesolver::FFAT_Map<double, 1> Map1;
    Eigen::Vector3i lo, hi, dm; Eigen::Matrix<double, 3, 1> gl;
    for (int j = 0; j < 3; ++j) { lo(j) = bbox_low_r[j]; hi(j) = bbox_top_r[j]; dm(j) = dim[j]; gl(j) = grid_low[j]; }
    std::vector<Eigen::Matrix<double, 3, 1>> V; std::vector<Eigen::Vector3i> F; std::vector<int> idx;
    std::vector<std::pair<int, int>> ne;
    Map1::CubemapMesh(lo, hi, cell_size, gl, dm, V, F, idx, ne);
    if ((int)V.size() > max_rows) return -(int)V.size();
    for (size_t i = 0; i < V.size(); ++i) for (int j = 0; j < 3; ++j) V_out[i * 3 + j] = V[i](j);
    for (int f = 0; f < 6; ++f) { n_elements_out[2 * f] = ne[f].first; n_elements_out[2 * f + 1] = ne[f].second; }
    if (data_indices_out) for (size_t i = 0; i < idx.size(); ++i) data_indices_out[i] = idx[i];
    return (int)V.size();
}
}
