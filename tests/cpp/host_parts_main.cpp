// Host-only pieces of the synthesis API -- temporal force profiles (forces.h), ModeData read / write / numModesAudible,
// ModalMaterial::Read / xi / omega_di -- run from one source compiled twice by tests/test_host_logic.py: against the header
// mirror (include/openpbso/) and against the REFERENCE's own headers read in place, to compare them number for number.
//   host_parts_main <modes file> <material file> <out dump.f64> <rewritten modes file>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "Eigen/Dense"
#include "config.h"
#include "ModalMaterial.h"
#include "ModeData.h"
#include "forces.h"

template <int BUF>
static void profiles(std::vector<double>& out) {
    std::unique_ptr<Force<double, BUF>> fs[4];
    fs[0].reset(new PointForce<double, BUF>());
    fs[1].reset(new GaussianForce<double, BUF>(700.0));
    fs[2].reset(new GaussianForce<double, BUF>(0.0));
    auto* ar = new AutoregressiveForce<double, BUF>();
    fs[3].reset(ar);
    for (int b = 0; b < 5; ++b) {
        if (b == 3) { AutoregressiveForceParam<double> prm; prm.a = {0.5, 0.2}; prm.sigma = 0.01; prm.mu = 0.3; ar->SetParam(prm); }
        for (auto& f : fs) {
            Eigen::Matrix<double, BUF, 1> spread; spread.setZero();
            const bool alive = f->Add(spread);
            out.push_back(alive ? 1.0 : 0.0);
            for (int i = 0; i < BUF; ++i) out.push_back(spread(i));
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    std::vector<double> out;
    profiles<64>(out); profiles<256>(out); profiles<513>(out);
    ModeData<double> modes;
    modes.read(argv[1]);
    out.push_back(modes.numModes()); out.push_back(modes.numDOF());
    for (int m = 0; m < modes.numModes(); ++m) { out.push_back(modes.omegaSquared(m)); for (int d = 0; d < modes.numDOF(); ++d) out.push_back(modes.mode(m)[d]); }
    for (double f : {50.0, 500.0, 5000.0, 22100.0, 1e9}) out.push_back(modes.numModesAudible(2600.0, f));
    out.push_back(modes.numModesAudible(2600.0, 777.0)); out.push_back(modes.numModesAudible(2600.0, 777.0));   // cached path
    modes.write(argv[4]);
    std::unique_ptr<ModalMaterial<double>> mat(ModalMaterial<double>::Read(argv[2]));
    out.push_back(mat ? 1.0 : 0.0);
    if (mat) {
        for (double v : {mat->density, mat->youngsModulus, mat->poissonRatio, mat->alpha, mat->beta}) out.push_back(v);
        for (double w : {300.0, 4000.0, 90000.0}) { out.push_back(mat->xi(w)); out.push_back(mat->omega_di(w)); }
    }
    std::unique_ptr<ModalMaterial<double>> none(ModalMaterial<double>::Read("/nonexistent/material.txt"));
    out.push_back(none ? 1.0 : 0.0);
    FILE* f = fopen(argv[3], "wb"); fwrite(out.data(), sizeof(double), out.size(), f); fclose(f);
    return 0;
}
