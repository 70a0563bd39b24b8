// pbso_fit_ffat -- headless FFAT map construction: the step that produces the `.fatcube` directory the synthesis
// path loads (FFAT_Map_Serialize::LoadAll / pbso_render -p).  The reference ships the fitting code
// (FFAT_Map<T,3> constructor + Solve, ffat_solver.h:944-1069) but no caller; this tool is that caller, with all modes
// of an object fitted in one launch on the B200 (kernel K6) through libpbso_b200.so.
//
//   pbso_fit_ffat -n N_ELEMENTS.txt -v VERTICES.f64 -c CELL_SIZE -k WAVENUMBERS.txt -p PRESSURE_TEMPLATE -o OUT_DIR
//                 [-b] [-s] [-first ID] [-compress] [-legacy]
//
//   -n   one line per shell: "Nx Ny" for the six faces +x,-x,+y,-y,+z,-z   (FFAT_Map<T,3>::ReadNElementsFile, :1100-1118)
//   -v   raw doubles, rows x 3: the cube-map mesh vertices, 4 per quad, shells back to back (CubemapMesh order, :334-397)
//   -k   one wavenumber per line; line i belongs to mode id FIRST + i
//   -p   printf template with one %d (the mode id) naming that mode's Dirichlet pressure file: two complex entries per
//        quad as the reference's triangle mesh orders them (ReadComplexVector, io.h:24-65); -b = binary (int count,
//        then count doubles), default text ("re im" per line)
//   -s   power scaling (Solve's powerScaling, :909-929)
//   -o   output directory: OUT_DIR/<mode id>.fatcube for every mode (FFAT_Map_Serialize::Save, ffat_map_serialize.h:90-164)
//   -compress   FFAT_Map<T,3>::Compress every map before saving (8-bit per-face quantisation, ffat_solver.h:1125-1178, without the
//               JPEG file round trip): the files carry _compressed_Psi and is_compressed
//   -legacy     write the legacy igl::serialize form (FFAT_Map<T,3>::Save, ffat_solver.h:1066-1068) instead of the protobuf one
#include <sys/stat.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>
#include "ffat_map_serialize.h"
#include "ffat_solver.h"
#include "io.h"

using namespace Gpu_Wavesolver;

static int fail(const char* what) {
    fprintf(stderr, "pbso_fit_ffat: %s: %s\n", what, pbso_last_error());
    return 1;
}

int main(int argc, char** argv) {
    std::map<std::string, std::string> opt;
    bool binary = false, scaling = false, compress = false, legacy = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-b") binary = true;
        else if (a == "-s") scaling = true;
        else if (a == "-compress") compress = true;
        else if (a == "-legacy") legacy = true;
        else if (a.size() > 1 && a[0] == '-' && i + 1 < argc) opt[a.substr(1)] = argv[++i];
        else { fprintf(stderr, "pbso_fit_ffat: unexpected argument %s\n", a.c_str()); return 2; }
    }
    for (const char* need : {"n", "v", "c", "k", "p", "o"})
        if (!opt.count(need)) {
            fprintf(stderr, "usage: pbso_fit_ffat -n N_ELEMENTS.txt -v VERTICES.f64 -c CELL_SIZE -k WAVENUMBERS.txt -p PRESSURE_TEMPLATE "
                            "-o OUT_DIR [-b] [-s] [-first ID] [-compress] [-legacy]\n");
            return 2;
        }
    const double cell = atof(opt["c"].c_str());
    const int first_id = opt.count("first") ? atoi(opt["first"].c_str()) : 0;

    if (!IsFile(opt["n"].c_str())) { fprintf(stderr, "pbso_fit_ffat: cannot open %s\n", opt["n"].c_str()); return 3; }
    std::vector<std::vector<std::pair<int, int>>> N_elements;
    FFAT_Map<double, 3>::ReadNElementsFile(opt["n"].c_str(), N_elements);
    std::vector<int> ne;
    for (const auto& shell : N_elements) for (const auto& p : shell) { ne.push_back(p.first); ne.push_back(p.second); }

    FILE* fv = fopen(opt["v"].c_str(), "rb");
    if (!fv) { fprintf(stderr, "pbso_fit_ffat: cannot open %s\n", opt["v"].c_str()); return 3; }
    fseek(fv, 0, SEEK_END); const long vbytes = ftell(fv); fseek(fv, 0, SEEK_SET);
    std::vector<double> V((size_t)vbytes / sizeof(double));
    if (fread(V.data(), sizeof(double), V.size(), fv) != V.size()) { fclose(fv); return 3; }
    fclose(fv);

    std::ifstream fk(opt["k"]);
    if (!fk) { fprintf(stderr, "pbso_fit_ffat: cannot open %s\n", opt["k"].c_str()); return 3; }
    std::vector<double> k;
    for (double x; fk >> x;) k.push_back(x);
    const int n_maps = (int)k.size();
    if (n_maps == 0) { fprintf(stderr, "pbso_fit_ffat: no wavenumbers in %s\n", opt["k"].c_str()); return 3; }

    pbso_ffat_fitter* fit = nullptr;
    if (pbso_ffat_fitter_create(cell, V.data(), (int)(V.size() / 3), ne.data(), (int)N_elements.size(), &fit)) return fail("fitter");
    int n_total = 0, n_dir = 0;
    pbso_ffat_fitter_info(fit, nullptr, &n_total, &n_dir, nullptr);

    // every mode's pressure, back to back, in the layout Solve indexes (2 * N_elements_total complex entries per mode)
    std::vector<double> P((size_t)n_maps * 4 * n_total);
    for (int m = 0; m < n_maps; ++m) {
        char name[4096];
        snprintf(name, sizeof(name), opt["p"].c_str(), first_id + m);
        if (!IsFile(name)) { fprintf(stderr, "pbso_fit_ffat: cannot open %s\n", name); return 3; }
        Eigen::Matrix<std::complex<double>, Eigen::Dynamic, 1> p;
        ReadComplexVector<double, double>(name, p, binary);
        if ((int)p.size() != 2 * n_total) {                  // Solve's assert (:1013)
            fprintf(stderr, "pbso_fit_ffat: %s holds %d entries, the shells need %d (Dirichlet pressure wrong size)\n", name, (int)p.size(), 2 * n_total);
            return 4;
        }
        double* dst = P.data() + (size_t)m * 4 * n_total;
        for (int i = 0; i < 2 * n_total; ++i) { dst[2 * i] = p(i).real(); dst[2 * i + 1] = p(i).imag(); }
    }

    std::vector<double> psi((size_t)n_maps * n_dir), scale(n_maps);
    if (pbso_ffat_fitter_solve(fit, n_maps, k.data(), P.data(), scaling ? 1 : 0, psi.data(), scale.data())) return fail("solve");
    float kernel_ms = 0.f;
    pbso_ffat_fitter_last_kernel_ms(fit, &kernel_ms);

    // shell 2 + Psi + k is the run-time map (what FFAT_Map_Serialize::Save keeps)
    double g2[32]; int ig2[18];
    pbso_ffat_fitter_shell(fit, 2, g2, ig2);
    std::vector<double> geom((size_t)n_maps * 32); std::vector<int> igeom((size_t)n_maps * 18), ids(n_maps);
    for (int m = 0; m < n_maps; ++m) {
        std::memcpy(&geom[(size_t)m * 32], g2, sizeof(g2)); geom[(size_t)m * 32 + 31] = k[m];
        std::memcpy(&igeom[(size_t)m * 18], ig2, sizeof(ig2));
        ids[m] = first_id + m;
    }
    pbso_ffat* maps = nullptr;
    if (pbso_ffat_create(n_maps, ids.data(), geom.data(), igeom.data(), psi.data(), n_dir, nullptr, &maps)) return fail("maps");
    if (compress && pbso_ffat_compress(maps, -1, nullptr)) return fail("compress");
    mkdir(opt["o"].c_str(), 0777);
    for (int m = 0; m < n_maps; ++m) {
        const std::string out = opt["o"] + "/" + std::to_string(ids[m]) + ".fatcube";
        if (legacy ? pbso_ffat_save_legacy_file(maps, ids[m], out.c_str()) : pbso_ffat_save_file(maps, ids[m], out.c_str())) return fail("save");
    }
    printf("pbso_fit_ffat: %d modes, %d shells, %d quads, %d directions, fit kernels %.3f ms -> %s/*.fatcube\n", n_maps,
           (int)N_elements.size(), n_total, n_dir, kernel_ms, opt["o"].c_str());
    pbso_ffat_destroy(maps);
    pbso_ffat_fitter_destroy(fit);
    return 0;
}
