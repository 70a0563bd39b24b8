// Host-only pieces of the headless driver, exercised without a GPU:
//   headless_units normals <mesh.obj> <out.bin>     int32 n_vertices, int32 n_faces, then n_vertices x 3 doubles
//   headless_units wav <in.f64> <n_per_buffer> <volume> <out.wav>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "mesh_io.h"
#include "wav_writer.h"

int main(int argc, char** argv) {
    if (argc >= 4 && !strcmp(argv[1], "normals")) {
        pbso_mesh::TriMesh mesh;
        if (!pbso_mesh::read_obj(argv[2], mesh)) return 3;
        const std::vector<double> N = pbso_mesh::per_vertex_normals(mesh);
        FILE* f = fopen(argv[3], "wb");
        const int nv = mesh.numVertices(), nf = mesh.numFaces();
        fwrite(&nv, 4, 1, f); fwrite(&nf, 4, 1, f);
        fwrite(N.data(), sizeof(double), N.size(), f);
        fclose(f);
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "wav")) {
        FILE* f = fopen(argv[2], "rb");
        if (!f) return 3;
        const int n = atoi(argv[3]);
        const double volume = atof(argv[4]);
        pbso_wav::StereoFloatWriter w(argv[5]);
        if (!w.ok()) return 3;
        std::vector<double> buf(n);
        size_t got;
        while ((got = fread(buf.data(), sizeof(double), n, f)) > 0) w.write(buf.data(), (int)got, volume);
        fclose(f);
        return 0;
    }
    return 2;
}
