// ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the REFERENCE's own headers (read in place
// from /root/reference via -I, never copied) against the Eigen shim and exposes them to ctypes so
// that tests/test_oracle_vs_ref.py can pin oracle/pbso_oracle.cpp to them.  Built by oracle/Makefile
// into oracle/_ref/libpbso_ref.so (git-ignored).  The reference's arithmetic statements are its own;
// only the Eigen container underneath (element-wise loops, a dot product) is the shim's.
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
