// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), modal_integrator.h
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: ModalIntegrator<T> (reference modal_integrator.h:19-123) over the C ABI.
// Solves  q'' + a q' + b q = f  per mode with the DyRT two-pole IIR; coefficients (kernel K2), state and
// Step() live on the B200.  Same constructor / Build / Step signatures and the same ownership: Build returns
// a raw new'd pointer, Step returns a reference that stays valid until the next Step.
#ifndef MODAL_INTEGRATOR_H
#define MODAL_INTEGRATOR_H
#include <cassert>
#include <vector>
#include "Eigen/Dense"
#include "pbso_check.h"

template <typename T>
class ModalIntegrator {
public:
    typedef Eigen::Matrix<T, Eigen::Dynamic, 1> ModalVec;
private:
    pbso_integrator* _h = nullptr;
    int _N = 0;
    ModalVec _q_k;                   // host mirror of the newest ring slot, what Step() hands back
    std::vector<double> _in, _out;   // double staging for T != double
public:
    ModalIntegrator(const int N, const T h, const ModalVec& a, const ModalVec& b) : _N(N) {
        assert(a.size() == N && "Vec a has wrong size");
        assert(b.size() == N && "Vec b has wrong size");
        std::vector<double> da(N), db(N);
        for (int i = 0; i < N; ++i) { da[i] = (double)a(i); db[i] = (double)b(i); }
        // an empty integrator (every mode culled by freq_threshold, tools/...cpp:309-345) is legal in the
        // reference and renders silence: it owns no device handle
        if (N > 0) pbso_mirror::check(pbso_integrator_create(N, (double)h, da.data(), db.data(), &_h), "ModalIntegrator");
        _q_k.setZero(N); _in.resize(N); _out.resize(N);
    }
    ~ModalIntegrator() { pbso_integrator_destroy(_h); }
    ModalIntegrator(const ModalIntegrator&) = delete;
    ModalIntegrator& operator=(const ModalIntegrator&) = delete;

    // (density, omega^2, alpha, beta) -> integrator for the first N modes; N < 0 takes all (reference :47-70)
    static ModalIntegrator<T>* Build(const T density, const std::vector<T> omegaSquared, const T alpha,
                                     const T beta, const T h, int N = -1) {
        if (N < 0) N = omegaSquared.size();
        else assert(N <= (int)omegaSquared.size() && "N for modal integrator invalid");
        ModalVec a, b; a.resize(N); b.resize(N);
        for (int ii = 0; ii < N; ++ii) {
            const T omega = sqrt(omegaSquared.at(ii) / density);
            const T xi = (T)0.5 * (alpha / omega + beta * omega);
            a(ii) = (T)2 * xi * omega;
            b(ii) = omega * omega;
        }
        return new ModalIntegrator<T>(N, h, a, b);
    }
    const ModalVec& Step(const ModalVec& Q) {
        assert(Q.size() == _N && "input force incorrect dimension");
        if (_N == 0) return _q_k;
        for (int i = 0; i < _N; ++i) _in[i] = (double)Q(i);
        pbso_mirror::check(pbso_integrator_step(_h, _in.data(), _out.data()), "ModalIntegrator::Step");
        for (int i = 0; i < _N; ++i) _q_k(i) = (T)_out[i];
        return _q_k;
    }
    const ModalVec& Step() {
        if (_N == 0) return _q_k;
        pbso_mirror::check(pbso_integrator_step(_h, nullptr, _out.data()), "ModalIntegrator::Step");
        for (int i = 0; i < _N; ++i) _q_k(i) = (T)_out[i];
        return _q_k;
    }
    // ---- B200 side ----
    pbso_integrator* handle() const { return _h; }
    int size() const { return _N; }
};
#endif
