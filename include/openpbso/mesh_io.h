// openpbso drop-in, headless side: the two libigl calls the reference's tool makes before it can place an
// impulse (tools/real_time_modal_sound.cpp:508-509): igl::read_triangle_mesh on a plain OBJ and
// igl::per_vertex_normals with libigl's default area weighting
// (external/libigl/include/igl/per_vertex_normals.cpp:61-68, 75-82, 104).  libigl is an empty submodule in the
// reference tree, so these are restated from its published behaviour; host-only, one-time work.
#ifndef PBSO_MESH_IO_H
#define PBSO_MESH_IO_H
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pbso_mesh {

struct TriMesh {
    std::vector<double> V;   // [n_vertices][3]
    std::vector<int> F;      // [n_faces][3], 0-based
    int numVertices() const { return (int)(V.size() / 3); }
    int numFaces() const { return (int)(F.size() / 3); }
};

// `v x y z` and `f a b c ...` records (indices 1-based, negative = relative to the end, `a/b/c` forms keep the
// first field); polygons are fan-triangulated.  Returns false if the file cannot be opened or holds no vertex.
inline bool read_obj(const std::string& path, TriMesh& mesh) {
    std::ifstream in(path.c_str());
    if (!in) return false;
    mesh.V.clear(); mesh.F.clear();
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tag; ls >> tag;
        if (tag == "v") {
            double x = 0, y = 0, z = 0; ls >> x >> y >> z;
            mesh.V.push_back(x); mesh.V.push_back(y); mesh.V.push_back(z);
        } else if (tag == "f") {
            std::vector<int> poly;
            std::string tok;
            while (ls >> tok) {
                const int i = std::atoi(tok.substr(0, tok.find('/')).c_str());
                poly.push_back(i > 0 ? i - 1 : mesh.numVertices() + i);
            }
            for (size_t k = 1; k + 1 < poly.size(); ++k) {
                mesh.F.push_back(poly[0]); mesh.F.push_back(poly[k]); mesh.F.push_back(poly[k + 1]);
            }
        }
    }
    return mesh.numVertices() > 0;
}

// N[v] = normalise( sum over incident faces of doublearea(f) * unit_normal(f) ) = normalise( sum of e1 x e2 ).
inline std::vector<double> per_vertex_normals(const TriMesh& mesh) {
    std::vector<double> N(mesh.V.size(), 0.0);
    for (int f = 0; f < mesh.numFaces(); ++f) {
        const int* t = &mesh.F[3 * (size_t)f];
        const double* a = &mesh.V[3 * (size_t)t[0]];
        const double* b = &mesh.V[3 * (size_t)t[1]];
        const double* c = &mesh.V[3 * (size_t)t[2]];
        const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        for (int j = 0; j < 3; ++j)
            for (int d = 0; d < 3; ++d) N[3 * (size_t)t[j] + d] += cr[d];
    }
    for (int v = 0; v < mesh.numVertices(); ++v) {
        double* n = &N[3 * (size_t)v];
        const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        n[0] /= len; n[1] /= len; n[2] /= len;
    }
    return N;
}

}  // namespace pbso_mesh
#endif
