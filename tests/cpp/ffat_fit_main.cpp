// Headless caller of the FFAT map construction API, written against the header mirror the way a reference-side
// preprocessing tool would be (the reference ships no caller of FFAT_Map<T,3>::Solve):
//   ffat_fit_main <n_elements.txt> <vertices.f64> <cellSize> <modeId> <k> <pressure file> <binary 0|1> <scaling 0|1>
//                 <out.fatcube> <probe x y z>
// vertices.f64: rows x 3 doubles (row-major).  Prints GetMapVal(probe) with 17 significant digits.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include "ffat_map_serialize.h"
#include "ffat_solver.h"
#include "io.h"

using namespace Gpu_Wavesolver;

int main(int argc, char** argv) {
    if (argc < 13) return 2;
    std::vector<std::vector<std::pair<int, int>>> N_elements;
    FFAT_Map<double, 3>::ReadNElementsFile(argv[1], N_elements);
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<double> raw((size_t)bytes / sizeof(double));
    if (fread(raw.data(), sizeof(double), raw.size(), f) != raw.size()) return 3;
    fclose(f);
    Eigen::Matrix<double, Eigen::Dynamic, 3> V; V.resize((int)(raw.size() / 3), 3);
    for (int i = 0; i < (int)V.rows(); ++i) for (int j = 0; j < 3; ++j) V(i, j) = raw[(size_t)i * 3 + j];
    const double cellSize = atof(argv[3]);
    const int modeId = atoi(argv[4]);
    const double k = atof(argv[5]);
    Eigen::Matrix<std::complex<double>, Eigen::Dynamic, 1> p;
    ReadComplexVector<double, double>(argv[6], p, atoi(argv[7]) != 0);
    FFAT_Map<double, 3> map(modeId, cellSize, V, N_elements);
    map.Solve(k, p, atoi(argv[8]) != 0);
    map.Solve(k, p, atoi(argv[8]) == 0);          // same k: returns at once, Psi unchanged (reference :1009-1010)
    FFAT_Map_Serialize::Save(argv[9], map);
    FFAT_Map<double, 3> back;
    FFAT_Map_Serialize::Load(argv[9], back);
    if (!FFAT_Map_Serialize::Check(map, back)) { fprintf(stderr, "Check failed after Save/Load\n"); return 4; }
    Eigen::Matrix<double, 3, 1> probe; probe << atof(argv[10]), atof(argv[11]), atof(argv[12]);
    printf("%d %d %.17g %.17g\n", (int)map.GetData().rows(), (int)map.GetData().cols(), map.GetMapVal(probe), back.GetMapVal(probe));
    // FFAT_Map<T,3>::Compress: 8-bit view next to _Psi; Save then keeps _compressed_Psi and the loaded map answers
    // GetMapVal(p, true) only
    const double amp = map.Compress();
    const std::string cfile = std::string(argv[9]) + ".compressed";
    FFAT_Map_Serialize::Save(cfile.c_str(), map);
    FFAT_Map<double, 3> cback;
    FFAT_Map_Serialize::Load(cfile.c_str(), cback);
    printf("%.17g %.17g %.17g %.17g\n", amp, map.GetMapVal(probe, true), cback.GetMapVal(probe, true), map.GetMapVal(probe));
    // the legacy file form (FFAT_Map<T,3>::Save / Load / LoadAll, reference :1066-1085) next to the protobuf one: the plain map
    // through LoadAll, the compressed one through Load
    const std::string ldir = std::string(argv[9]) + ".legacy";
    if (system(("mkdir -p '" + ldir + "'").c_str()) != 0) return 5;
    FFAT_Map<double, 3>::Save((ldir + "/m.fatcube").c_str(), back);
    std::map<int, FFAT_Map<double, 3>>* all = FFAT_Map<double, 3>::LoadAll(ldir.c_str());
    const std::string lcfile = std::string(argv[9]) + ".legacy-compressed";
    FFAT_Map<double, 3>::Save(lcfile.c_str(), map);
    FFAT_Map<double, 3> lback;
    FFAT_Map<double, 3>::Load(lcfile.c_str(), lback);
    printf("%d %.17g %.17g\n", (int)all->size(), lback.GetMapVal(probe, true), all->at(modeId).GetMapVal(probe));
    delete all;
    return 0;
}
