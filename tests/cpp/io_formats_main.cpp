// Reads / writes the FFAT-fit input formats and dumps what was parsed.  Compiled twice by tests/test_host_logic.py: against the
// header mirror (include/openpbso/) and against the REFERENCE's own io.h / ffat_solver.h (read in place, with the Eigen shim and
// the stubs of oracle/ref_stubs/), so that the two parsers can be compared on the same files.
//   io_formats_main read <complex file> <binary 0|1> <n_elements file> <dump.f64>
//   io_formats_main write <in dump.f64> <binary 0|1> <out complex file>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ffat_solver.h"
#include "io.h"

using namespace Gpu_Wavesolver;

int main(int argc, char** argv) {
    if (argc >= 6 && !strcmp(argv[1], "read")) {
        Eigen::Matrix<std::complex<double>, Eigen::Dynamic, 1> p;
        ReadComplexVector<double, double>(argv[2], p, atoi(argv[3]) != 0);
        std::vector<std::vector<std::pair<int, int>>> ne;
        FFAT_Map<double, 3>::ReadNElementsFile(argv[4], ne);
        std::vector<double> out;
        out.push_back((double)p.size());
        for (int i = 0; i < (int)p.size(); ++i) { out.push_back(p(i).real()); out.push_back(p(i).imag()); }
        out.push_back((double)ne.size());
        for (const auto& s : ne) for (const auto& q : s) { out.push_back(q.first); out.push_back(q.second); }
        FILE* f = fopen(argv[5], "wb"); fwrite(out.data(), sizeof(double), out.size(), f); fclose(f);
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "write")) {
        FILE* f = fopen(argv[2], "rb"); if (!f) return 3;
        fseek(f, 0, SEEK_END); const long n = ftell(f) / (long)sizeof(double); fseek(f, 0, SEEK_SET);
        std::vector<double> v((size_t)n);
        if (fread(v.data(), sizeof(double), v.size(), f) != v.size()) return 3;
        fclose(f);
        Eigen::Matrix<std::complex<double>, Eigen::Dynamic, 1> p; p.resize((int)(n / 2));
        for (int i = 0; i < (int)(n / 2); ++i) p(i) = std::complex<double>(v[2 * i], v[2 * i + 1]);
        WriteComplexVector<double>(argv[4], p, atoi(argv[3]) != 0);
        return 0;
    }
    return 2;
}
