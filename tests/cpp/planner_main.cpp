// Host-only harness of the offline planner (tools/offline_render.h): plays a message script given as explicit modal vectors --
// no mesh, no device -- and dumps the plan, so that tests can replay the same messages on the oracle's ModalSolver::step and
// compare what each buffer would have been given.
//   planner_main <N> <script> <out.bin>
// script lines:  point v0..vN-1 | gauss WIDTH_US v.. | ar_start v.. | ar_data v.. | ar_end | arprm a0 a1 sigma mu | clear |
//                trans t0..tN-1 | unit_transfer | use_transfer | run K
// out.bin: int32 n_buffers, int32 n_transfers, then per buffer {int32 kind, int32 trans, double space[N], double time[BUF]}
// (space / time zero-filled where the plan holds none; an impulse buffer has time = delta), then the transfers [n][N].
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "offline_render.h"

static const int BUF = 64;

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int N = atoi(argv[1]);
    pbso_offline::Planner<BUF> planner(N);
    std::ifstream script(argv[2]);
    if (!script) return 3;
    std::string line;
    auto vec = [&](std::istringstream& in, ForceMessage<double, BUF>& m) { m.data.resize(N); for (int i = 0; i < N; ++i) in >> m.data(i); };
    while (std::getline(script, line)) {
        std::istringstream in(line);
        std::string kind;
        if (!(in >> kind)) continue;
        ForceMessage<double, BUF> m;
        bool send = true;
        if (kind == "point") vec(in, m);
        else if (kind == "gauss") { double w; in >> w; vec(in, m); m.forceType = ForceType::GaussianForce; m.force.reset(new GaussianForce<double, BUF>(w)); }
        else if (kind == "ar_start" || kind == "ar_data") {
            vec(in, m); m.forceType = ForceType::AutoregressiveForce; m.force.reset(new AutoregressiveForce<double, BUF>());
            m.sustainedForceStart = kind == "ar_start";
        } else if (kind == "ar_end") {
            m.data.setZero(N); m.forceType = ForceType::AutoregressiveForce; m.force.reset(new AutoregressiveForce<double, BUF>()); m.sustainedForceEnd = true;
        } else if (kind == "clear") { m.data.setZero(N); m.clearAllForces = true; }
        else {
            send = false;
            if (kind == "arprm") { AutoregressiveForceParam<double> p; in >> p.a[0] >> p.a[1] >> p.sigma >> p.mu; planner.enqueueArprmMessage(p); }
            else if (kind == "trans") { std::vector<double> t((size_t)N); for (auto& x : t) in >> x; if (!planner.enqueueTransMessage(t)) fprintf(stderr, "trans dropped\n"); }
            else if (kind == "unit_transfer") planner.setUseTransfer(false);
            else if (kind == "use_transfer") planner.setUseTransfer(true);
            else if (kind == "run") { long k = 0; in >> k; for (long i = 0; i < k; ++i) planner.step(); }
            else return 4;
        }
        if (send && !planner.enqueueForceMessage(m)) return 5;
    }
    FILE* f = fopen(argv[3], "wb");
    if (!f) return 6;
    const auto& plan = planner.plan;
    const int nb = (int)plan.buffers.size(), nt = (int)plan.transfers.size();
    fwrite(&nb, 4, 1, f); fwrite(&nt, 4, 1, f);
    for (const auto& b : plan.buffers) {
        const int kind = (int)b.kind;
        fwrite(&kind, 4, 1, f); fwrite(&b.trans, 4, 1, f);
        std::vector<double> sp((size_t)N, 0.0), tm((size_t)BUF, 0.0);
        if (!b.space.empty()) sp = b.space;
        if (!b.time.empty()) tm = b.time; else if (kind == 1) tm[0] = 1.0;
        fwrite(sp.data(), 8, sp.size(), f); fwrite(tm.data(), 8, tm.size(), f);
    }
    for (const auto& t : plan.transfers) fwrite(t.data(), 8, t.size(), f);
    fclose(f);
    printf("%ld stepped, %d produced\n", planner.stepped, nb);
    return 0;
}
