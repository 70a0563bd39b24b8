// Offline mode of pbso_render (-batch): the whole script is first PLANNED on the host -- the message bookkeeping of
// ModalSolver::step (reference modal_solver.h:181-259: one force message per buffer, the active-force list and its
// (sum of loads) x (sum of profiles) rank-1 force, sustained forces, clearAllForces, the transfer swap) restated without
// the hot loop -- and then rendered range by range on the batch path of libpbso_b200:
//   * stretches of buffers whose force is an impulse on sample 0 (PointForce, forces.h:81-90) or nothing at all are ONE
//     stateful batch render each (pbso_batch_set_state / set_impulses / render_mix / get_end_state);
//   * a TransMessage that is dequeued in mid-script is a range boundary (pbso_batch_set_transfer in between);
//   * buffers in which a Gaussian or autoregressive force is alive go through the per-buffer path (pbso_render_buffer) from
//     the state the batch range ended in, and hand their state back.
// -gpus R: the object's modes are split into R blocks, one rank (thread + device) each; a rank reshapes its block into
// pseudo-objects of -block modes, renders its partial track and ONE sum-reduce (pbso_comm_reduce_audio_host, NCCL) lands the
// track on rank 0 -- SURVEY 8(e)'s mode-block sharding, through the same C ABI calls bench.py uses.
#ifndef PBSO_OFFLINE_RENDER_H
#define PBSO_OFFLINE_RENDER_H
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <deque>
#include <list>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "modal_solver.h"

namespace pbso_offline {

enum BufKind { Silent = 0, Impulse = 1, General = 2 };

template <int BUF>
struct Plan {
    struct Buffer {
        BufKind kind = Silent;
        int trans = 0;                    // index into transfers
        std::vector<double> space;        // rank-1 force: modal load (empty when silent) ...
        std::vector<double> time;         // ... and temporal profile (General only; Impulse: space already carries time(0))
    };
    std::vector<Buffer> buffers;                      // produced buffers, in order
    std::vector<std::vector<double>> transfers;       // [0] = TransMessage::setToUnit (1e7)
};

// ModalSolver<double, BUF>::step()'s message handling (include/openpbso/modal_solver.h:129-169 = reference :183-259),
// recording what the hot loop would have been given instead of running it.
template <int BUF>
class Planner {
    typedef ForceMessage<double, BUF> FMsg;
    const int _N;
    std::deque<FMsg> _queue_force;                                    // capacity 512 (modal_solver.h:113)
    std::deque<std::vector<double>> _queue_trans;                     // capacity 1
    std::deque<AutoregressiveForceParam<double>> _queue_arprm;        // capacity 1
    std::list<FMsg> _activeForces;
    bool _sustainedForces = false, _useTransfer = true, _latestIsUnit = true;
    int _latest = 0;
public:
    Plan<BUF> plan;
    long stepped = 0;
    explicit Planner(int N) : _N(N) { plan.transfers.push_back(std::vector<double>((size_t)N, 1E7)); }
    bool enqueueForceMessage(const FMsg& m) { if (_queue_force.size() >= 512) return false; _queue_force.push_back(m); return true; }
    bool transQueueFull() const { return !_queue_trans.empty(); }
    bool enqueueTransMessage(const std::vector<double>& t) { if (transQueueFull()) return false; _queue_trans.push_back(t); return true; }
    void enqueueArprmMessage(const AutoregressiveForceParam<double>& p) { if (_queue_arprm.empty()) _queue_arprm.push_back(p); else _queue_arprm.back() = p; }
    void setUseTransfer(bool s) { _useTransfer = s; }
    long produced() const { return (long)plan.buffers.size(); }

    void step() {
        ++stepped;
        if (!_queue_force.empty()) {
            FMsg mess = _queue_force.front(); _queue_force.pop_front();
            if (mess.clearAllForces) { _activeForces.clear(); return; }                                   // no buffer is produced
            if (mess.sustainedForceStart) { _activeForces.clear(); _sustainedForces = true; _activeForces.push_back(mess); }
            if (!_sustainedForces) _activeForces.push_back(mess);
            else _activeForces.begin()->data = mess.data;
            if (mess.sustainedForceEnd) { _activeForces.clear(); _sustainedForces = false; }
        }
        Eigen::Matrix<double, BUF, 1> time; time.setZero();
        Eigen::Matrix<double, -1, 1> space; space.setZero(_N);
        if (!_sustainedForces) {
            for (auto it = _activeForces.begin(); it != _activeForces.end();) {
                if (!it->force->Add(time)) it = _activeForces.erase(it);
                else { space += it->data; ++it; }
            }
        } else {
            auto it = _activeForces.begin();
            if (it->forceType == ForceType::AutoregressiveForce && !_queue_arprm.empty()) {
                static_cast<AutoregressiveForce<double, BUF>*>(it->force.get())->SetParam(_queue_arprm.front());
                _queue_arprm.pop_front();
            }
            it->force->Add(time);
            space = it->data;
        }
        if (_useTransfer) {
            if (!_queue_trans.empty()) {
                plan.transfers.push_back(_queue_trans.front()); _queue_trans.pop_front();
                _latest = (int)plan.transfers.size() - 1; _latestIsUnit = false;
            }
        } else if (!_latestIsUnit) { _latest = 0; _latestIsUnit = true; }
        typename Plan<BUF>::Buffer b;
        b.trans = _latest;
        bool tail = false, any_space = false;
        for (int i = 1; i < BUF; ++i) tail = tail || time(i) != 0.0;
        for (int i = 0; i < _N; ++i) any_space = any_space || space(i) != 0.0;
        if (any_space && (tail || time(0) != 0.0)) {
            b.kind = tail ? General : Impulse;
            b.space.resize((size_t)_N);
            for (int i = 0; i < _N; ++i) b.space[(size_t)i] = space(i) * (tail ? 1.0 : time(0));
            if (tail) { b.time.resize(BUF); for (int i = 0; i < BUF; ++i) b.time[(size_t)i] = time(i); }
        }
        plan.buffers.push_back(std::move(b));
    }
};

inline void check(int rc, const char* what) {
    if (rc != PBSO_OK) throw std::runtime_error(std::string(what) + ": " + pbso_last_error());
}

struct RenderOptions {
    int gpus = 1;
    int block = 256;                  // modes per pseudo-object
    int precision = PBSO_PREC_TC3X;
};

// One rank's share of the track: modes [lo, hi) of the object.
template <int BUF>
static void render_rank(const Plan<BUF>& plan, const std::vector<double>& a, const std::vector<double>& b, double h,
                        int lo, int hi, const RenderOptions& ro, std::vector<double>& part, long* ranges_out) {
    const int nb = (int)plan.buffers.size();
    part.assign((size_t)nb * BUF, 0.0);
    const int n_blk = hi - lo;
    if (n_blk <= 0 || nb == 0) return;
    const int P = std::min(std::max(16, ro.block), (n_blk + 15) / 16 * 16), n_obj = (n_blk + P - 1) / P, np = n_obj * P;
    // pseudo-objects: block o holds modes lo + o P ...; the last one is padded with copies of the last mode that nothing excites
    std::vector<double> pa((size_t)np), pb((size_t)np);
    for (int i = 0; i < np; ++i) { const int m = lo + std::min(i, n_blk - 1); pa[(size_t)i] = a[(size_t)m]; pb[(size_t)i] = b[(size_t)m]; }
    pbso_batch* bt = nullptr; pbso_integrator* it = nullptr;
    check(pbso_batch_create(n_obj, P, h, pa.data(), pb.data(), &bt), "pbso_batch_create");
    auto pad = [&](const std::vector<double>& full) { std::vector<double> v((size_t)np, 0.0); std::copy(full.begin() + lo, full.begin() + hi, v.begin()); return v; };
    std::vector<double> q1((size_t)np, 0.0), q2((size_t)np, 0.0), y((size_t)BUF);
    bool zero_state = true;
    long ranges = 0;
    try {
        int i = 0;
        while (i < nb) {
            if (plan.buffers[(size_t)i].kind != General) {
                // ---- a batch range: up to the next general buffer or transfer change ----
                int j = i;
                while (j < nb && plan.buffers[(size_t)j].kind != General && plan.buffers[(size_t)j].trans == plan.buffers[(size_t)i].trans) ++j;
                std::vector<int> obj, buf; std::vector<double> space;
                for (int k = i; k < j; ++k) {
                    if (plan.buffers[(size_t)k].kind != Impulse) continue;
                    const std::vector<double> sp = pad(plan.buffers[(size_t)k].space);
                    for (int o = 0; o < n_obj; ++o) { obj.push_back(o); buf.push_back(k - i); }
                    space.insert(space.end(), sp.begin(), sp.end());
                }
                if (zero_state && obj.empty()) { i = j; continue; }                            // silence from rest
                const std::vector<double> tr = pad(plan.transfers[(size_t)plan.buffers[(size_t)i].trans]);
                check(pbso_batch_set_transfer(bt, tr.data()), "pbso_batch_set_transfer");
                check(pbso_batch_set_state(bt, zero_state ? nullptr : q1.data(), zero_state ? nullptr : q2.data()), "pbso_batch_set_state");
                check(pbso_batch_set_impulses(bt, (int)obj.size(), obj.data(), buf.data(), space.data()), "pbso_batch_set_impulses");
                check(pbso_batch_render_mix(bt, BUF, j - i, ro.precision, 0, part.data() + (size_t)i * BUF), "pbso_batch_render_mix");
                check(pbso_batch_get_end_state(bt, BUF, j - i, q1.data(), q2.data()), "pbso_batch_get_end_state");
                zero_state = false; ++ranges;
                i = j;
            } else {
                // ---- general forces: the per-buffer path from the same state ----
                if (!it) check(pbso_integrator_create(n_blk, h, pa.data(), pb.data(), &it), "pbso_integrator_create");   // the first n_blk entries are the block
                check(pbso_integrator_set_state(it, q1.data(), q2.data()), "pbso_integrator_set_state");                  // (the padding sits behind them)
                int tr_now = -1;
                while (i < nb && plan.buffers[(size_t)i].kind == General) {
                    const typename Plan<BUF>::Buffer& pbuf = plan.buffers[(size_t)i];
                    if (pbuf.trans != tr_now) {
                        check(pbso_integrator_set_transfer(it, plan.transfers[(size_t)pbuf.trans].data() + lo, n_blk, 1), "pbso_integrator_set_transfer");
                        tr_now = pbuf.trans;
                    }
                    check(pbso_render_buffer(it, pbuf.space.data() + lo, pbuf.time.data(), BUF, y.data(), nullptr), "pbso_render_buffer");
                    std::copy(y.begin(), y.end(), part.begin() + (size_t)i * BUF);
                    ++i;
                }
                std::fill(q1.begin(), q1.end(), 0.0); std::fill(q2.begin(), q2.end(), 0.0);
                check(pbso_integrator_get_state(it, q1.data(), q2.data()), "pbso_integrator_get_state");
                zero_state = false; ++ranges;
            }
        }
    } catch (...) { pbso_batch_destroy(bt); pbso_integrator_destroy(it); throw; }
    pbso_batch_destroy(bt); pbso_integrator_destroy(it);
    if (ranges_out) *ranges_out = ranges;
}

// All ranks; returns the track (produced buffers x BUF) on the caller's side.
template <int BUF>
static std::vector<double> render_plan(const Plan<BUF>& plan, const std::vector<double>& a, const std::vector<double>& b, double h,
                                       const RenderOptions& ro, long* ranges_out) {
    const int N = (int)a.size(), R = std::max(1, ro.gpus);
    const size_t ns = plan.buffers.size() * (size_t)BUF;
    if (R == 1) {
        std::vector<double> track;
        render_rank<BUF>(plan, a, b, h, 0, N, ro, track, ranges_out);
        return track;
    }
    int n_dev = 0; check(pbso_device_count(&n_dev), "pbso_device_count");
    if (n_dev < R) throw std::runtime_error("-gpus " + std::to_string(R) + ": only " + std::to_string(n_dev) + " CUDA device(s) visible");
    unsigned char uid[128]; check(pbso_comm_unique_id(uid), "pbso_comm_unique_id");
    std::vector<std::vector<double>> parts((size_t)R);
    std::vector<std::string> errors((size_t)R);
    std::vector<long> ranges((size_t)R, 0);
    std::vector<std::thread> threads;
    for (int r = 0; r < R; ++r)
        threads.emplace_back([&, r]() {
            pbso_comm* comm = nullptr;
            try {
                check(pbso_set_device(r), "pbso_set_device");
                check(pbso_comm_init(R, r, uid, &comm), "pbso_comm_init");
            } catch (const std::exception& e) { errors[(size_t)r] = e.what(); return; }
            try {
                long long lo = 0, hi = 0;
                check(pbso_comm_shard(comm, ((long long)N + 15) / 16, &lo, &hi), "pbso_comm_shard");    // whole 16-mode K chunks per rank
                render_rank<BUF>(plan, a, b, h, (int)std::min<long long>(16 * lo, N), (int)std::min<long long>(16 * hi, N), ro, parts[(size_t)r], &ranges[(size_t)r]);
            } catch (const std::exception& e) { errors[(size_t)r] = e.what(); }
            // a rank whose render failed still joins the reduce (with what it has): the others are waiting in it
            parts[(size_t)r].resize(ns, 0.0);
            if (pbso_comm_reduce_audio_host(comm, parts[(size_t)r].data(), ns, 0) != PBSO_OK && errors[(size_t)r].empty()) errors[(size_t)r] = pbso_last_error();
            pbso_comm_destroy(comm);
        });
    for (auto& t : threads) t.join();
    for (int r = 0; r < R; ++r) if (!errors[(size_t)r].empty()) throw std::runtime_error("rank " + std::to_string(r) + ": " + errors[(size_t)r]);
    if (ranges_out) *ranges_out = ranges[0];
    return parts[0];
}

}  // namespace pbso_offline
#endif
