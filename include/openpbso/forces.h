// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), forces.h
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: temporal force profiles (reference forces.h:12-137).  Host-side by design: a profile is
// BUF_SIZE numbers per buffer and feeds the device kernel through ModalSolver::step.  The autoregressive
// force keeps the reference's generator types (std::default_random_engine + std::normal_distribution,
// default-seeded) so that its sample stream is the same under the same standard library.
#ifndef FORCES_H
#define FORCES_H
#include <algorithm>
#include <cmath>
#include <random>
#include <vector>
#include "Eigen/Dense"
#include "config.h"

enum class ForceType { PointForce = 0, GaussianForce = 1, AutoregressiveForce = 2 };

template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
class Force {
public:
    // Adds this force's contribution for the next buffer; false once the force has died out.
    virtual bool Add(Eigen::Matrix<T, BUF_SIZE, 1>& forceSpread) = 0;
    virtual ~Force() = default;
};

// Unit impulse on sample 0 of the first buffer it sees (reference :81-90).
template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
class PointForce : public Force<T, BUF_SIZE> {
    bool used = false;
public:
    bool Add(Eigen::Matrix<T, BUF_SIZE, 1>& forceSpread) override {
        if (used) return false;
        forceSpread(0) += 1.;
        used = true;
        return true;
    }
};

// Gaussian pulse of `width` microseconds, 10 sigma support, may span buffers (reference :33-48, 92-105).
template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
class GaussianForce : public Force<T, BUF_SIZE> {
    T _width;
    int _widthSamples;
    int _count = 0;
    int _center;
    int _cutoff = 5;
public:
    GaussianForce(const T width) : _width(width) {
        _widthSamples = std::max(1, (int)(_width / 1000000. * SAMPLE_RATE));
        _center = (int)((_cutoff - 0.5) * _widthSamples);
    }
    bool Add(Eigen::Matrix<T, BUF_SIZE, 1>& forceSpread) override {
        if (_width == 0 || _count >= _cutoff * 2 * _widthSamples) return false;
        for (int ii = 0; ii < BUF_SIZE; ++ii) {
            const T z = (T)(_count + ii - _center) / (T)_widthSamples;
            forceSpread(ii) += std::exp(-(T)0.5 * std::pow(z, 2));
        }
        _count += BUF_SIZE;
        return true;
    }
};

template <typename T>
struct AutoregressiveForceParam {
    std::vector<T> a = {0.783, 0.116};
    T sigma = 0.00148;
    T mu = 0.142;
};

// AR(2) scraping force, Pai et al. 2001 (reference :60-79, 107-137).
template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
class AutoregressiveForce : public Force<T, BUF_SIZE> {
    std::vector<T> _buf;
    const int _bufLen;
    int _bufIdx = 0;
    std::vector<T> _a;
    T _sigma;
    T _mu;
    std::default_random_engine _generator;
    std::normal_distribution<T> _distribution;
    T GetMuEffective() {
        T mu_tilde = (T)0.0;
        for (int ii = 0; ii < 2; ++ii) mu_tilde += _a.at(ii) * _buf.at((_bufIdx + _bufLen - ii - 1) % _bufLen);
        mu_tilde += _sigma * _distribution(_generator);
        _buf.at(_bufIdx) = mu_tilde;
        _bufIdx = (_bufIdx + 1) % _bufLen;
        return _mu + mu_tilde;
    }
public:
    AutoregressiveForce() : _buf{0, 0, 0}, _bufLen(_buf.size()), _a{0.783, 0.116}, _sigma(0.00148), _mu(0.142) {}
    bool Add(Eigen::Matrix<T, BUF_SIZE, 1>& forceSpread) override {
        for (int ii = 0; ii < BUF_SIZE; ++ii) forceSpread(ii) += GetMuEffective();
        return true;
    }
    void SetParam(const AutoregressiveForceParam<T>& param) {
        _buf = {0, 0, 0};
        _a = param.a; _sigma = param.sigma; _mu = param.mu;
    }
};
#endif
