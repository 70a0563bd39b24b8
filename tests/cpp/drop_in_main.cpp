// A headless caller written against the reference's header API (the names in tools/real_time_modal_sound.cpp:
// ReadMaterial / ReadModes / BuildSolver / GetModalForceVertex / computeTransfer / step / dequeueSoundMessage),
// compiled against include/openpbso/ and linked to libpbso_b200.so.  tests/test_gpu_dropin.py feeds it a
// script and compares every buffer with the CPU oracle.
//
//   drop_in_main <dir> <out.bin>
//     <dir>/material.txt  <dir>/object.modes  <dir>/ffat/*.fatcube [+ freq_threshold.txt]  <dir>/script.txt
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include "ModalMaterial.h"
#include "ModeData.h"
#include "modal_solver.h"
#include "modal_force.h"

constexpr int BUF = 256;
typedef ModalSolver<double, BUF> Solver;
typedef ForceMessage<double, BUF> FMsg;

// tools/real_time_modal_sound.cpp:309-345
static Solver* BuildSolver(const std::unique_ptr<ModalMaterial<double>>& material,
                           const std::unique_ptr<ModeData<double>>& modes, const std::string& ffatMapFolder,
                           int& N_modesAudible) {
    std::ifstream stream((ffatMapFolder + "/freq_threshold.txt").c_str());
    N_modesAudible = modes->numModes();
    if (stream) {
        std::string line; std::getline(stream, line);
        std::istringstream iss(line); double maxFreq; iss >> maxFreq;
        N_modesAudible = modes->numModesAudible(material->density, maxFreq);
    } else {
        N_modesAudible = modes->numModesAudible(material->density, 20000.);
    }
    Solver* solver = new Solver(N_modesAudible);
    std::shared_ptr<ModalIntegrator<double>> integrator(ModalIntegrator<double>::Build(
        material->density, modes->_omegaSquared, material->alpha, material->beta, 1. / (double)SAMPLE_RATE, N_modesAudible));
    solver->setIntegrator(integrator);
    solver->readFFATMaps(ffatMapFolder);
    return solver;
}

static void set_force(FMsg& m, ForceType t, double width) {
    m.forceType = t;
    if (t == ForceType::PointForce) m.force.reset(new PointForce<double, BUF>());
    else if (t == ForceType::GaussianForce) m.force.reset(new GaussianForce<double, BUF>(width));
    else m.force.reset(new AutoregressiveForce<double, BUF>());
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <dir> <out.bin>\n", argv[0]); return 2; }
    const std::string d = argv[1];
    try {
        std::unique_ptr<ModalMaterial<double>> material(ModalMaterial<double>::Read((d + "/material.txt").c_str()));
        if (!material) { fprintf(stderr, "no material\n"); return 3; }
        std::unique_ptr<ModeData<double>> modes(new ModeData<double>());
        modes->read((d + "/object.modes").c_str());
        int N = 0;
        std::unique_ptr<Solver> solver(BuildSolver(material, modes, d + "/ffat", N));
        FILE* out = fopen(argv[2], "wb");
        fwrite(&N, sizeof(int), 1, out);
        std::ifstream script((d + "/script.txt").c_str());
        std::string line;
        SoundMessage<double, BUF> sound;
        while (std::getline(script, line)) {
            std::istringstream in(line);
            std::string kind; in >> kind;
            FMsg msg; bool send = false;
            auto vertex = [&](ForceType t, double width) {
                int vid; Eigen::Vector3d vn; in >> vid >> vn[0] >> vn[1] >> vn[2];
                GetModalForceVertex<double, BUF>(N, *modes, vid, vn, msg);
                set_force(msg, t, width); send = true;
            };
            if (kind == "point") vertex(ForceType::PointForce, 0);
            else if (kind == "gauss") { double w; in >> w; vertex(ForceType::GaussianForce, w); }
            else if (kind == "face") {
                Eigen::Vector3i v; Eigen::Vector3d bc, vn;
                in >> v[0] >> v[1] >> v[2] >> bc[0] >> bc[1] >> bc[2] >> vn[0] >> vn[1] >> vn[2];
                GetModalForceFace<double, BUF>(N, *modes, v, bc, vn, msg);
                set_force(msg, ForceType::PointForce, 0); send = true;
            }
            else if (kind == "clear") { msg.data.setZero(N); msg.clearAllForces = true; send = true; }
            else if (kind == "ar_start") { vertex(ForceType::AutoregressiveForce, 0); msg.sustainedForceStart = true; }
            else if (kind == "ar_data") { vertex(ForceType::AutoregressiveForce, 0); }
            else if (kind == "ar_end") { vertex(ForceType::AutoregressiveForce, 0); msg.sustainedForceEnd = true; }
            else if (kind == "arprm") {
                AutoregressiveForceParam<double> p; in >> p.a[0] >> p.a[1] >> p.sigma >> p.mu;
                solver->enqueueArprmMessage(p);
            }
            else if (kind == "listener") {
                Eigen::Vector3d pos; in >> pos[0] >> pos[1] >> pos[2];
                const bool ok = solver->computeTransfer(pos);
                if (!ok) fprintf(stderr, "computeTransfer dropped\n");
            }
            else if (kind == "unit_transfer") solver->setUseTransfer(false);
            else if (kind == "use_transfer") solver->setUseTransfer(true);
            if (send && !solver->enqueueForceMessage(msg)) { fprintf(stderr, "force queue full\n"); return 4; }
            solver->step();
            int produced = solver->dequeueSoundMessage(sound) ? 1 : 0;
            fwrite(&produced, sizeof(int), 1, out);
            if (produced) {
                fwrite(sound.data.data(), sizeof(double), BUF, out);
                Eigen::Matrix<double, -1, 1> qn = solver->getQBufferNorm();
                fwrite(qn.data(), sizeof(double), N, out);
            }
        }
        // batched listener call site (tools/...cpp:921-927): all maps, caller memory
        {
            Eigen::Vector3d pos; pos << 2.0, -3.0, 4.0;
            std::vector<double> tr(4096, 0.0);
            solver->computeTransfer(pos, tr.data());
            fwrite(tr.data(), sizeof(double), N, out);
        }
        fclose(out);
    } catch (const std::exception& e) {
        fprintf(stderr, "drop_in_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
