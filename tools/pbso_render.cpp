// pbso_render -- headless replacement for the reference's GUI tool (tools/real_time_modal_sound.cpp): the same
// inputs and path conventions (:478-501, .meta files :389-397), the same BuildSolver (:309-345), the same
// GetModalForceVertex/Face + ForceMessage flow (:236-295, :607-611, :756-775) and the PortAudio callback's output
// convention (:207-210), driven by an impulse script instead of the mouse and writing a WAV file instead of
// the sound card.  All synthesis runs on the B200 through include/openpbso/ -> libpbso_b200.so.
//
//   pbso_render (-d DIR [-name N] | -meta FILE.meta | -m MESH.obj -s MODES -t MATERIAL -p FFAT_DIR)
//               -script SCRIPT [-buf 64|128|256|512|513] [-o OUT.wav] [-raw OUT.f64] [-volume V] [-stats]
//               [-batch [-prec tc3x|f32|f64] [-gpus N] [-block MODES]]
//
// -batch renders the same script OFFLINE on the batch path (tools/offline_render.h): the script is planned on the host
// with ModalSolver::step's message semantics, impulse stretches become stateful batch renders (tensor cores by default),
// transfer swaps are range boundaries and buffers with a live Gaussian / autoregressive force use the per-buffer path in
// between.  -gpus N splits the object's modes over N devices and sum-reduces the track with NCCL.
//
// Script: one command per line, `#` starts a comment.  Commands that send a message only enqueue it; `run`
// steps the solver (which consumes at most one force message per buffer, modal_solver.h:183).
//   run N                         step N buffers
//   until SECONDS                 step until SECONDS of audio have been produced
//   listener X Y Z                computeTransfer(pos) -> transfer message (swapped in at the next buffer)
//   hit VID                       impulse at vertex VID along its area-weighted normal (the tool's shift-click, :607-609)
//   point VID NX NY NZ            PointForce at vertex VID along (NX,NY,NZ)
//   gauss WIDTH_US VID NX NY NZ   GaussianForce of WIDTH_US microseconds
//   face V0 V1 V2 B0 B1 B2 NX NY NZ   PointForce at barycentric (B0,B1,B2) of face (V0,V1,V2)
//   clear                         clearAllForces (the buffer that consumes it produces no audio, :186-189)
//   ar_start VID NX NY NZ | ar_data VID NX NY NZ | ar_end    sustained autoregressive force
//   arprm A0 A1 SIGMA MU          AutoregressiveForceParam
//   unit_transfer | use_transfer  setUseTransfer(false|true)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "ModalMaterial.h"
#include "ModeData.h"
#include "io.h"
#include "mesh_io.h"
#include "modal_force.h"
#include "modal_solver.h"
#include "wav_writer.h"
#include "offline_render.h"

struct Paths { std::string obj, modes, material, ffat; };

// tools/real_time_modal_sound.cpp:480-501 (+ :389-397 for .meta)
static bool resolve_paths(const std::map<std::string, std::string>& opt, Paths& p) {
    auto get = [&](const char* k) { auto it = opt.find(k); return it == opt.end() ? std::string() : it->second; };
    if (!get("d").empty()) {
        const std::string d = get("d");
        std::string n = get("name");
        if (n.empty()) {
            std::vector<std::string> filenames;
            Gpu_Wavesolver::ListDirFiles(d.c_str(), filenames, ".tet.obj");
            if (filenames.empty()) { fprintf(stderr, "pbso_render: no *.tet.obj in %s\n", d.c_str()); return false; }
            n = Gpu_Wavesolver::Basename(filenames.at(0));
            n = n.substr(0, n.find_first_of("."));
        }
        std::cout << "object name: " << n << std::endl;
        p.obj = d + "/" + n + ".tet.obj";
        p.modes = d + "/" + n + "_surf.modes";
        p.material = d + "/" + n + "_material.txt";
        p.ffat = d + "/" + n + "_ffat_maps";
        return true;
    }
    if (!get("meta").empty()) {
        std::ifstream stream(get("meta").c_str());
        if (!stream) { fprintf(stderr, "pbso_render: cannot open %s\n", get("meta").c_str()); return false; }
        std::getline(stream, p.obj); std::getline(stream, p.modes);
        std::getline(stream, p.material); std::getline(stream, p.ffat);
        return true;
    }
    p.obj = get("m"); p.modes = get("s"); p.material = get("t"); p.ffat = get("p");
    return !p.modes.empty() && !p.material.empty();
}

// tools/real_time_modal_sound.cpp:309-345
template <int BUF>
static ModalSolver<double, BUF>* BuildSolver(const std::unique_ptr<ModalMaterial<double>>& material,
                           const std::unique_ptr<ModeData<double>>& modes, const std::string& ffatMapFolder,
                           int& N_modesAudible) {
    std::ifstream stream((ffatMapFolder + "/freq_threshold.txt").c_str());
    if (stream) {
        std::string line; std::getline(stream, line);
        std::istringstream iss(line); double maxFreq = 0; iss >> maxFreq;
        N_modesAudible = modes->numModesAudible(material->density, maxFreq);
    } else {
        N_modesAudible = modes->numModesAudible(material->density, 20000.);
    }
    ModalSolver<double, BUF>* solver = new ModalSolver<double, BUF>(N_modesAudible);
    std::shared_ptr<ModalIntegrator<double>> integrator(ModalIntegrator<double>::Build(
        material->density, modes->_omegaSquared, material->alpha, material->beta, 1. / (double)SAMPLE_RATE, N_modesAudible));
    solver->setIntegrator(integrator);
    if (!ffatMapFolder.empty()) solver->readFFATMaps(ffatMapFolder);
    return solver;
}

template <int BUF>
static void set_profile(ForceMessage<double, BUF>& m, ForceType t, double width_us) {
    m.forceType = t;
    if (t == ForceType::PointForce) m.force.reset(new PointForce<double, BUF>());
    else if (t == ForceType::GaussianForce) m.force.reset(new GaussianForce<double, BUF>(width_us));
    else m.force.reset(new AutoregressiveForce<double, BUF>());
}

template <int BUF>
static int render(std::map<std::string, std::string>& opt, const Paths& paths);

int main(int argc, char** argv) {
    std::map<std::string, std::string> opt;
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i];
        if (k.size() < 2 || k[0] != '-') { fprintf(stderr, "pbso_render: unexpected argument %s\n", argv[i]); return 2; }
        k = k.substr(k[1] == '-' ? 2 : 1);
        if (k == "stats" || k == "batch") { opt[k] = "1"; continue; }
        if (i + 1 >= argc) { fprintf(stderr, "pbso_render: -%s needs a value\n", k.c_str()); return 2; }
        opt[k] = argv[++i];
    }
    Paths paths;
    if (opt.count("script") == 0 || !resolve_paths(opt, paths)) {
        fprintf(stderr, "usage: pbso_render (-d DIR [-name N] | -meta FILE | -m OBJ -s MODES -t MATERIAL -p FFAT_DIR) "
                        "-script FILE [-buf N] [-o OUT.wav] [-raw OUT.f64] [-volume V] [-stats] [-batch [-prec tc3x|f32|f64] [-gpus N] [-block MODES]]\n");
        return 2;
    }
    // BUF_SIZE is a template parameter of the reference's solver (modal_solver.h:100); FRAMES_PER_BUFFER = 513 is its default.
    const int buf = opt.count("buf") ? std::atoi(opt["buf"].c_str()) : FRAMES_PER_BUFFER;
    switch (buf) {
        case 64: return render<64>(opt, paths);
        case 128: return render<128>(opt, paths);
        case 256: return render<256>(opt, paths);
        case 512: return render<512>(opt, paths);
        case FRAMES_PER_BUFFER: return render<FRAMES_PER_BUFFER>(opt, paths);
        default: fprintf(stderr, "pbso_render: -buf must be one of 64 128 256 512 %d\n", FRAMES_PER_BUFFER); return 2;
    }
}

// What a script command acts on: the live solver (real-time path, one buffer per step) or the offline planner.
template <int BUF>
struct LiveDriver {
    ModalSolver<double, BUF>* solver;
    pbso_wav::StereoFloatWriter* wav; FILE* raw; double volume;
    std::vector<double> lat_us; long n_produced = 0, n_stepped = 0;
    SoundMessage<double, BUF> sound;
    void step() {
        const auto t0 = std::chrono::steady_clock::now();
        solver->step();
        const auto t1 = std::chrono::steady_clock::now();
        lat_us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
        ++n_stepped;
        if (solver->dequeueSoundMessage(sound)) {
            ++n_produced;
            if (wav) wav->write(sound.data.data(), BUF, volume);
            if (raw) fwrite(sound.data.data(), sizeof(double), BUF, raw);
        }
        (void)solver->getQBufferNorm();              // the GUI's consumer; keeps the lossy queue drained
    }
    long produced() const { return n_produced; }
    bool listener(const Eigen::Vector3d& pos) { return solver->computeTransfer(pos); }
    bool send(const ForceMessage<double, BUF>& m) { return solver->enqueueForceMessage(m); }
    void arprm(const AutoregressiveForceParam<double>& p) { solver->enqueueArprmMessageNoFail(p, 1000); }
    void useTransfer(bool s) { solver->setUseTransfer(s); }
};

template <int BUF>
struct PlanDriver {
    ModalSolver<double, BUF>* solver;                // evaluates the FFAT maps (kernel K3) for `listener`
    pbso_offline::Planner<BUF> planner;
    PlanDriver(ModalSolver<double, BUF>* s, int N) : solver(s), planner(N) {}
    void step() { planner.step(); }
    long produced() const { return planner.produced(); }
    bool listener(const Eigen::Vector3d& pos) {
        if (planner.transQueueFull() || !solver->computeTransfer(pos)) return false;     // queue capacity 1 (modal_solver.h:113)
        TransMessage<double> t;
        if (!solver->dequeueTransMessage(t)) return false;
        return planner.enqueueTransMessage(std::vector<double>(t.data.data(), t.data.data() + t.data.size()));
    }
    bool send(const ForceMessage<double, BUF>& m) { return planner.enqueueForceMessage(m); }
    void arprm(const AutoregressiveForceParam<double>& p) { planner.enqueueArprmMessage(p); }
    void useTransfer(bool s) { planner.setUseTransfer(s); }
};

template <int BUF, typename Driver>
static int play_script(const std::string& script_path, int N, const ModeData<double>& modes, const pbso_mesh::TriMesh& mesh,
                       const std::vector<double>& VN, Driver& drv);

template <int BUF>
static int render(std::map<std::string, std::string>& opt, const Paths& paths) {
    typedef ModalSolver<double, BUF> Solver;
    typedef ForceMessage<double, BUF> FMsg;
    const double volume = opt.count("volume") ? std::atof(opt["volume"].c_str()) : 1.0;
    try {
        std::unique_ptr<ModalMaterial<double>> material(ModalMaterial<double>::Read(paths.material.c_str()));
        if (!material) { fprintf(stderr, "pbso_render: cannot read material %s\n", paths.material.c_str()); return 3; }
        if (!Gpu_Wavesolver::IsFile(paths.modes.c_str())) { fprintf(stderr, "pbso_render: cannot read modes %s\n", paths.modes.c_str()); return 3; }
        std::unique_ptr<ModeData<double>> modes(new ModeData<double>());
        modes->read(paths.modes.c_str());
        pbso_mesh::TriMesh mesh;
        std::vector<double> VN;
        if (!paths.obj.empty()) {
            if (!pbso_mesh::read_obj(paths.obj, mesh)) { fprintf(stderr, "pbso_render: cannot read mesh %s\n", paths.obj.c_str()); return 3; }
            if (modes->numDOF() != mesh.numVertices() * 3) { fprintf(stderr, "pbso_render: DOFs mismatch\n"); return 3; }   // :515
            VN = pbso_mesh::per_vertex_normals(mesh);
        }
        int N = 0;
        std::unique_ptr<Solver> solver(BuildSolver<BUF>(material, modes, paths.ffat, N));
        std::cout << "modes audible: " << N << " of " << modes->numModes() << std::endl;

        std::unique_ptr<pbso_wav::StereoFloatWriter> wav;
        if (opt.count("o")) {
            wav.reset(new pbso_wav::StereoFloatWriter(opt["o"]));
            if (!wav->ok()) { fprintf(stderr, "pbso_render: cannot write %s\n", opt["o"].c_str()); return 3; }
        }
        FILE* raw = opt.count("raw") ? fopen(opt["raw"].c_str(), "wb") : nullptr;
        std::vector<double> lat_us;
        long produced = 0, stepped = 0;
        if (opt.count("batch")) {
            // ---- offline: plan the script, then render it on the batch path ----
            PlanDriver<BUF> drv(solver.get(), N);
            if (int rc = play_script<BUF>(opt["script"], N, *modes, mesh, VN, drv)) return rc;
            pbso_offline::RenderOptions ro;
            if (opt.count("gpus")) ro.gpus = std::max(1, std::atoi(opt["gpus"].c_str()));
            if (opt.count("block")) ro.block = std::atoi(opt["block"].c_str());
            if (opt.count("prec")) {
                const std::string& pr = opt["prec"];
                if (pr == "tc3x") ro.precision = PBSO_PREC_TC3X; else if (pr == "f32") ro.precision = PBSO_PREC_F32_TILED; else if (pr == "f64") ro.precision = PBSO_PREC_F64;
                else { fprintf(stderr, "pbso_render: -prec must be tc3x, f32 or f64\n"); return 2; }
            }
            // (a, b) of the audible modes as ModalIntegrator::Build derives them (modal_integrator.h:47-70)
            std::vector<double> a((size_t)N), b((size_t)N);
            for (int ii = 0; ii < N; ++ii) {
                const double omega = sqrt(modes->_omegaSquared.at(ii) / material->density);
                const double xi = 0.5 * (material->alpha / omega + material->beta * omega);
                a[(size_t)ii] = 2 * xi * omega; b[(size_t)ii] = omega * omega;
            }
            const auto t0 = std::chrono::steady_clock::now();
            long ranges = 0;
            const std::vector<double> track = pbso_offline::render_plan<BUF>(drv.planner.plan, a, b, 1. / (double)SAMPLE_RATE, ro, &ranges);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            produced = drv.planner.produced(); stepped = drv.planner.stepped;
            for (long bi = 0; bi < produced; ++bi) {
                if (wav) wav->write(track.data() + (size_t)bi * BUF, BUF, volume);
                if (raw) fwrite(track.data() + (size_t)bi * BUF, sizeof(double), BUF, raw);
            }
            std::cout << "offline: " << ranges << " range(s) on " << ro.gpus << " device(s), " << ms << " ms" << std::endl;
        } else {
            LiveDriver<BUF> drv{solver.get(), wav.get(), raw, volume};
            if (int rc = play_script<BUF>(opt["script"], N, *modes, mesh, VN, drv)) return rc;
            produced = drv.n_produced; stepped = drv.n_stepped; lat_us.swap(drv.lat_us);
        }
        if (wav) wav->close();
        if (raw) fclose(raw);
        std::cout << "buffers: " << produced << " produced / " << stepped << " stepped, "
                  << (double)produced * BUF / SAMPLE_RATE << " s of audio" << std::endl;
        if (opt.count("stats") && !lat_us.empty()) {
            std::sort(lat_us.begin(), lat_us.end());
            auto q = [&](double f) { return lat_us[std::min(lat_us.size() - 1, (size_t)(f * lat_us.size()))]; };
            printf("step latency us: p50 %.1f p99 %.1f max %.1f (budget %.1f)\n", q(0.5), q(0.99), lat_us.back(),
                   1e6 * BUF / SAMPLE_RATE);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "pbso_render: %s\n", e.what());
        return 1;
    }
    return 0;
}

template <int BUF, typename Driver>
static int play_script(const std::string& script_path, int N, const ModeData<double>& modes, const pbso_mesh::TriMesh& mesh,
                       const std::vector<double>& VN, Driver& drv) {
    typedef ForceMessage<double, BUF> FMsg;
        std::ifstream script(script_path.c_str());
        if (!script) { fprintf(stderr, "pbso_render: cannot open script %s\n", script_path.c_str()); return 3; }
        std::string line;
        int lineno = 0;
        while (std::getline(script, line)) {
            ++lineno;
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.resize(hash);
            std::istringstream in(line);
            std::string kind;
            if (!(in >> kind)) continue;
            FMsg msg;
            bool send = false;
            auto vertex = [&](ForceType t, double width_us) {
                int vid = 0; Eigen::Vector3d vn; in >> vid >> vn[0] >> vn[1] >> vn[2];
                GetModalForceVertex<double, BUF>(N, modes, vid, vn, msg);
                set_profile(msg, t, width_us); send = true;
            };
            if (kind == "run") { long n = 0; in >> n; for (long i = 0; i < n; ++i) drv.step(); }
            else if (kind == "until") {
                double sec = 0; in >> sec;
                while ((double)drv.produced() * BUF < sec * SAMPLE_RATE) drv.step();
            }
            else if (kind == "listener") {
                Eigen::Vector3d pos; in >> pos[0] >> pos[1] >> pos[2];
                if (!drv.listener(pos)) fprintf(stderr, "pbso_render:%d: transfer message dropped\n", lineno);
            }
            else if (kind == "hit") {
                int vid = 0; in >> vid;
                if (VN.empty() || vid < 0 || vid >= mesh.numVertices()) { fprintf(stderr, "pbso_render:%d: hit needs a mesh and a valid vertex\n", lineno); return 4; }
                Eigen::Vector3d vn; vn << VN[3 * (size_t)vid], VN[3 * (size_t)vid + 1], VN[3 * (size_t)vid + 2];
                GetModalForceVertex<double, BUF>(N, modes, vid, vn, msg);
                send = true;
            }
            else if (kind == "point") vertex(ForceType::PointForce, 0);
            else if (kind == "gauss") { double w = 0; in >> w; vertex(ForceType::GaussianForce, w); }
            else if (kind == "face") {
                Eigen::Vector3i v; Eigen::Vector3d bc, vn;
                in >> v[0] >> v[1] >> v[2] >> bc[0] >> bc[1] >> bc[2] >> vn[0] >> vn[1] >> vn[2];
                GetModalForceFace<double, BUF>(N, modes, v, bc, vn, msg);
                send = true;
            }
            else if (kind == "clear") { msg.data.setZero(N); msg.clearAllForces = true; send = true; }
            else if (kind == "ar_start") { vertex(ForceType::AutoregressiveForce, 0); msg.sustainedForceStart = true; }
            else if (kind == "ar_data") vertex(ForceType::AutoregressiveForce, 0);
            else if (kind == "ar_end") {                 // the tool's dummy end signal (:764-772)
                GetModalForceVertex<double, BUF>(N, modes, 0, Eigen::Vector3d::Zero(), msg);
                set_profile(msg, ForceType::AutoregressiveForce, 0); msg.sustainedForceEnd = true; send = true;
            }
            else if (kind == "arprm") {
                AutoregressiveForceParam<double> p; in >> p.a[0] >> p.a[1] >> p.sigma >> p.mu;
                drv.arprm(p);
            }
            else if (kind == "unit_transfer") drv.useTransfer(false);
            else if (kind == "use_transfer") drv.useTransfer(true);
            else { fprintf(stderr, "pbso_render:%d: unknown command '%s'\n", lineno, kind.c_str()); return 4; }
            if (in.fail()) { fprintf(stderr, "pbso_render:%d: malformed '%s'\n", lineno, kind.c_str()); return 4; }
            if (send && !drv.send(msg)) { fprintf(stderr, "pbso_render:%d: force queue full\n", lineno); return 4; }
        }
    return 0;
}
