// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), modal_solver.h
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: ModalSolver<T, BUF_SIZE> (reference modal_solver.h:22-399) -- the per-buffer scheduler.
// Message types, queue capacities and the force / transfer state machine of step() follow the reference; the
// hot loop (BUF_SIZE integrator steps + transfer-weighted modal sum + per-mode RMS, :261-272) is one launch of
// kernel K1 through pbso_render_buffer, and computeTransfer is one launch of kernel K3 for all modes.
// Differences from the reference that a caller can observe: none in values beyond FP64 rounding; the data
// race between getLatestTransfer() and step() is gone (the solver owns its copy); a missing GPU throws.
#ifndef MODAL_SOLVER_H
#define MODAL_SOLVER_H
#include <cassert>
#include <iostream>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "config.h"
#include "Eigen/Dense"
#include "external/readerwriterqueue.h"
#include "modal_integrator.h"
#include "ffat_solver.h"
#include "ffat_map_serialize.h"
#include "forces.h"

template <typename T>
struct DataMessage {
    Eigen::Matrix<T, -1, 1> data;
};

template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
struct ForceMessage {
    Eigen::Matrix<T, -1, 1> data;                     // modal load U^T f
    ForceType forceType = ForceType::PointForce;
    std::unique_ptr<Force<T, BUF_SIZE>> force;        // temporal profile
    bool sustainedForceStart = false;
    bool sustainedForceEnd = false;
    bool clearAllForces = false;
    ForceMessage() : force(new PointForce<T, BUF_SIZE>()) {}
    ForceMessage(const ForceMessage& o) { *this = o; }
    ForceMessage& operator=(const ForceMessage& o) {   // deep copy incl. the profile's internal state
        if (&o == this) return *this;
        data = o.data; forceType = o.forceType;
        sustainedForceStart = o.sustainedForceStart; sustainedForceEnd = o.sustainedForceEnd;
        clearAllForces = o.clearAllForces;
        switch (o.forceType) {
            case ForceType::PointForce:
                force.reset(new PointForce<T, BUF_SIZE>(*static_cast<PointForce<T, BUF_SIZE>*>(o.force.get()))); break;
            case ForceType::GaussianForce:
                force.reset(new GaussianForce<T, BUF_SIZE>(*static_cast<GaussianForce<T, BUF_SIZE>*>(o.force.get()))); break;
            case ForceType::AutoregressiveForce:
                force.reset(new AutoregressiveForce<T, BUF_SIZE>(*static_cast<AutoregressiveForce<T, BUF_SIZE>*>(o.force.get()))); break;
            default:
                std::cout << static_cast<int>(o.forceType) << std::endl;
                assert(false && "unrecognized force type");
        }
        return *this;
    }
};

template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
struct SoundMessage {
    Eigen::Matrix<T, BUF_SIZE, 1> data;
};

template <typename T>
struct TransMessage {
    bool useCompressed = false;
    int N = 0;
    Eigen::Matrix<T, -1, 1> data;
    void setToUnit() { data.setOnes(N); data *= 1E7; }
    explicit TransMessage() = default;
    explicit TransMessage(const int N_) : N(N_) { setToUnit(); }
};

template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
class ModalSolver {
    typedef Gpu_Wavesolver::FFAT_Map<T, 3> FFAT_Map;
    moodycamel::ReaderWriterQueue<ForceMessage<T, BUF_SIZE>> _queue_force;
    moodycamel::ReaderWriterQueue<SoundMessage<T, BUF_SIZE>> _queue_sound;
    moodycamel::ReaderWriterQueue<TransMessage<T>> _queue_trans;
    moodycamel::ReaderWriterQueue<DataMessage<T>> _queue_qnorm;
    moodycamel::ReaderWriterQueue<AutoregressiveForceParam<T>> _queue_arprm;
    SoundMessage<T, BUF_SIZE> _mess_sound;
    ForceMessage<T, BUF_SIZE> _mess_force;
    TransMessage<T> _mess_trans;
    TransMessage<T> _latest_transfer;
    DataMessage<T> _mess_qnorm;
    std::shared_ptr<ModalIntegrator<T>> _integrator;
    const int _N_modes;
    std::shared_ptr<pbso_ffat> _ffat_maps;            // the whole directory as one device set
    int _N_ffat_maps = 0;
    std::list<ForceMessage<T, BUF_SIZE>> _activeForces;
    Eigen::Matrix<T, -1, 1> _forceSpreadBufferSpace;
    Eigen::Matrix<T, BUF_SIZE, 1> _forceSpreadBufferTime;
    std::mutex _useTransferMutex;
    bool _useTransfer = true;
    bool _useTransferCache = true;
    bool _sustainedForces = false;
    bool _transferDirty = true;                       // device copy of _latest_transfer is stale
    bool _latestIsUnit = true;                        // _latest_transfer currently holds setToUnit()'s values
    std::vector<double> _stage_space, _stage_time, _stage_trans, _stage_y, _stage_qnorm;

public:
    explicit ModalSolver(const int N_modes)
        : _queue_force(512), _queue_sound(2), _queue_trans(1), _queue_qnorm(2), _queue_arprm(1),
          _mess_trans(N_modes), _latest_transfer(N_modes), _N_modes(N_modes) {
        _forceSpreadBufferSpace.resize(_N_modes);
        _mess_qnorm.data.setZero(_N_modes);
        _stage_space.resize(_N_modes); _stage_time.resize(BUF_SIZE); _stage_y.resize(BUF_SIZE); _stage_qnorm.resize(_N_modes);
    }
    inline void setIntegrator(std::shared_ptr<ModalIntegrator<T>> integrator) { _integrator = integrator; _transferDirty = true; }
    inline const TransMessage<T>& getLatestTransfer() { return _latest_transfer; }
    inline void setUseTransfer(const bool s) { std::lock_guard<std::mutex> g(_useTransferMutex); _useTransfer = s; }
    inline Eigen::Matrix<T, -1, 1> getQBufferNorm() {
        DataMessage<T> mess;
        if (_queue_qnorm.try_dequeue(mess)) return mess.data;
        return Eigen::Matrix<T, -1, 1>::Zero(_N_modes);
    }

    // One audio buffer (reference :181-276).
    void step() {
        // -- at most one force message per buffer --
        if (dequeueForceMessage(_mess_force)) {
            if (_mess_force.clearAllForces) { _activeForces.clear(); return; }     // no buffer is produced
            if (_mess_force.sustainedForceStart) { _activeForces.clear(); _sustainedForces = true; _activeForces.push_back(_mess_force); }
            if (!_sustainedForces) _activeForces.push_back(_mess_force);
            else _activeForces.begin()->data = _mess_force.data;
            if (_mess_force.sustainedForceEnd) { _activeForces.clear(); _sustainedForces = false; }
        }
        // -- rank-1 force of this buffer: (sum of spatial loads) x (sum of temporal profiles) --
        _forceSpreadBufferTime.setZero();
        if (!_sustainedForces) {
            _forceSpreadBufferSpace.setZero();
            for (auto it = _activeForces.begin(); it != _activeForces.end();) {
                assert(it->force && "obsolete forces should be removed");
                if (!it->force->Add(_forceSpreadBufferTime)) it = _activeForces.erase(it);
                else { _forceSpreadBufferSpace += it->data; ++it; }
            }
        } else {
            assert(_activeForces.size() == 1 && "Should only have 1 concurrent sustained force");
            auto it = _activeForces.begin();
            if (it->forceType == ForceType::AutoregressiveForce) {
                AutoregressiveForceParam<T> arprm;
                if (dequeueArprmMessage(arprm))
                    static_cast<AutoregressiveForce<T, BUF_SIZE>*>(it->force.get())->SetParam(arprm);
            }
            it->force->Add(_forceSpreadBufferTime);
            _forceSpreadBufferSpace = it->data;
        }
        // -- transfer: swap at buffer boundaries, no cross-fade --
        bool useTransfer = _useTransferCache;
        if (_useTransferMutex.try_lock()) { useTransfer = _useTransfer; _useTransferMutex.unlock(); }
        if (useTransfer) {
            TransMessage<T> trans;
            if (dequeueTransMessage(trans)) { _latest_transfer = trans; _transferDirty = true; _latestIsUnit = false; }
        } else if (!_latestIsUnit) {
            // the reference re-fills the unit vector every buffer (:254); its values never change, so the device
            // table is refreshed only on the switch from a real transfer to the unit one
            _latest_transfer.setToUnit(); _transferDirty = true; _latestIsUnit = true;
        }
        _useTransferCache = useTransfer;
        assert(_forceSpreadBufferSpace.size() == _N_modes && "dimension of force message incorrect");

        if (_N_modes == 0) {                                   // empty solver: the reference's loop yields zeros
            _mess_sound.data.setZero();
            _mess_qnorm.data.resize(0);
            _queue_qnorm.try_enqueue(_mess_qnorm);
            enqueueSoundMessageNoFail(_mess_sound, -1);
            return;
        }
        // -- hot loop on the B200: BUF_SIZE x (Step + dot + q^2), then sqrt --
        pbso_integrator* h = _integrator->handle();
        if (_transferDirty) {
            const int nt = (int)_latest_transfer.data.size();
            _stage_trans.resize(nt);
            for (int i = 0; i < nt; ++i) _stage_trans[i] = (double)_latest_transfer.data(i);
            pbso_mirror::check(pbso_integrator_set_transfer(h, _stage_trans.data(), nt, 1), "ModalSolver::step");
            _transferDirty = false;
        }
        for (int i = 0; i < _N_modes; ++i) _stage_space[i] = (double)_forceSpreadBufferSpace(i);
        for (int i = 0; i < BUF_SIZE; ++i) _stage_time[i] = (double)_forceSpreadBufferTime(i);
        pbso_mirror::check(pbso_render_buffer(h, _stage_space.data(), _stage_time.data(), BUF_SIZE, _stage_y.data(),
                                              _stage_qnorm.data()), "ModalSolver::step");
        for (int i = 0; i < BUF_SIZE; ++i) _mess_sound.data(i) = (T)_stage_y[i];
        _mess_qnorm.data.resize(_N_modes);
        for (int i = 0; i < _N_modes; ++i) _mess_qnorm.data(i) = (T)_stage_qnorm[i];
        _queue_qnorm.try_enqueue(_mess_qnorm);            // lossy by design
        enqueueSoundMessageNoFail(_mess_sound, -1);       // back-pressure: waits for the audio thread
    }

    void readFFATMaps(const std::string& mapFolderPath) {
        pbso_ffat* h = nullptr;
        const int rc = pbso_ffat_load_dir(mapFolderPath.c_str(), &h);   // LoadAll: empty map on a bad directory
        _ffat_maps = std::shared_ptr<pbso_ffat>(h, [](pbso_ffat* p) { pbso_ffat_destroy(p); });
        if (rc != PBSO_OK && rc != PBSO_ERR_IO) pbso_mirror::check(rc, "ModalSolver::readFFATMaps");
        pbso_mirror::check(pbso_ffat_num_maps(h, &_N_ffat_maps), "ModalSolver::readFFATMaps");
    }
    // trans[m] = |map_m(pos)| for the solver's modes, enqueued for the sim thread; false if no maps were
    // read or the previous message has not been consumed yet (queue capacity 1).
    bool computeTransfer(const Eigen::Matrix<T, 3, 1>& pos) {
        if (!_ffat_maps) return false;
        const int N = _mess_trans.data.size();
        std::vector<double> out(N);
        eval(pos, N, out.data());
        for (int ii = 0; ii < N; ++ii) _mess_trans.data(ii) = (T)out[ii];
        return enqueueTransMessage(_mess_trans);
    }
    // Same evaluation for every map that was read, written to caller memory (batched listener call site).
    bool computeTransfer(const Eigen::Matrix<T, 3, 1>& pos, T* trans) {
        if (!_ffat_maps) return false;
        const int N = _N_ffat_maps;
        std::vector<double> out(N);
        eval(pos, N, out.data());
        for (int ii = 0; ii < N; ++ii) trans[ii] = (T)out[ii];
        return true;
    }
    // L listeners at once (the reference loops computeTransfer over 10 242 directions, tools/...cpp:921-927):
    // trans is column-major N x L.
    bool computeTransferBatch(const T* pos_Lx3, int L, T* trans) {
        if (!_ffat_maps) return false;
        std::vector<double> p(pos_Lx3, pos_Lx3 + 3 * (size_t)L), out((size_t)L * _N_ffat_maps);
        pbso_mirror::check(pbso_ffat_eval(_ffat_maps.get(), _N_ffat_maps, p.data(), L, _mess_trans.useCompressed, out.data()),
                           "ModalSolver::computeTransferBatch");
        for (size_t i = 0; i < out.size(); ++i) trans[i] = (T)out[i];
        return true;
    }

    bool enqueueForceMessageNoFail(const ForceMessage<T, BUF_SIZE>& mess, const int maxIte = -1) { return spin(_queue_force, mess, maxIte); }
    bool enqueueForceMessage(const ForceMessage<T, BUF_SIZE>& mess) { return _queue_force.try_enqueue(mess); }
    bool dequeueForceMessage(ForceMessage<T, BUF_SIZE>& mess) { return _queue_force.try_dequeue(mess); }
    bool enqueueSoundMessage(const SoundMessage<T, BUF_SIZE>& mess) { return _queue_sound.try_enqueue(mess); }
    bool enqueueSoundMessageNoFail(const SoundMessage<T, BUF_SIZE>& mess, const int maxIte = -1) { return spin(_queue_sound, mess, maxIte); }
    bool dequeueSoundMessage(SoundMessage<T, BUF_SIZE>& mess) { return _queue_sound.try_dequeue(mess); }
    bool enqueueTransMessage(const TransMessage<T>& mess) { return _queue_trans.try_enqueue(mess); }
    bool dequeueTransMessage(TransMessage<T>& mess) { return _queue_trans.try_dequeue(mess); }
    bool enqueueArprmMessage(const AutoregressiveForceParam<T>& mess) { return _queue_arprm.try_enqueue(mess); }
    bool enqueueArprmMessageNoFail(const AutoregressiveForceParam<T>& mess, const int maxIte = -1) { return spin(_queue_arprm, mess, maxIte); }
    bool dequeueArprmMessage(AutoregressiveForceParam<T>& mess) { return _queue_arprm.try_dequeue(mess); }

private:
    template <typename Q, typename M>
    static bool spin(Q& q, const M& mess, const int maxIte) {
        int ite = 0;
        while (maxIte < 0 || ite++ < maxIte)
            if (q.try_enqueue(mess)) return true;
        return false;
    }
    void eval(const Eigen::Matrix<T, 3, 1>& pos, int N, double* out) {
        const double p[3] = {(double)pos(0), (double)pos(1), (double)pos(2)};
        pbso_mirror::check(pbso_ffat_eval(_ffat_maps.get(), N, p, 1, _mess_trans.useCompressed, out), "ModalSolver::computeTransfer");
    }
};
#endif
