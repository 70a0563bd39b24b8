// Harness for the reference's OWN GetModalForceFace / GetModalForceVertex.  They live inside the GUI program
// tools/real_time_modal_sound.cpp, which cannot be compiled here; tests/test_oracle_vs_ref.py cuts exactly those two function
// templates out of the reference file AT TEST TIME into REF_MODAL_FORCE_EXTRACT (a temporary file: nothing of the reference is
// stored in this repository) and compiles this harness around them with the reference's headers read in place.
//   ref_modal_force_main <modes file> <forceDim> <out.f64>  < "vertex vid nx ny nz" | "face v0 v1 v2 b0 b1 b2 nx ny nz" lines
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
#include "ModeData.h"
#include "forces.h"
#include "modal_solver.h"

// the two members of the tool's settings object the functions read
struct { ForceType forceType = ForceType::PointForce; struct { double timeScale = 300.0; } gaussianForceParameters; } VIEWER_SETTINGS;

#include REF_MODAL_FORCE_EXTRACT

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    ModeData<double> modes; modes.read(argv[1]);
    const int forceDim = atoi(argv[2]);
    std::vector<double> out;
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream iss(line);
        std::string kind; iss >> kind;
        ForceMessage<double> msg;
        if (kind == "vertex") {
            int vid; Eigen::Vector3d vn; iss >> vid >> vn[0] >> vn[1] >> vn[2];
            GetModalForceVertex<double>(forceDim, modes, vid, vn, msg);
        } else if (kind == "face") {
            Eigen::Vector3i v; Eigen::Vector3d b, vn;
            iss >> v[0] >> v[1] >> v[2] >> b[0] >> b[1] >> b[2] >> vn[0] >> vn[1] >> vn[2];
            GetModalForceFace<double>(forceDim, modes, v, b, vn, msg);
        } else continue;
        assert((int)msg.data.size() == forceDim && msg.forceType == ForceType::PointForce && msg.force);
        for (int m = 0; m < forceDim; ++m) out.push_back(msg.data(m));
    }
    FILE* f = fopen(argv[3], "wb"); fwrite(out.data(), sizeof(double), out.size(), f); fclose(f);
    return 0;
}
