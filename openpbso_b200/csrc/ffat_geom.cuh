// Cube-map geometry shared by the FFAT kernels (ffat.cu: run-time evaluation K3; ffat_fit.cu: map construction K6).
// Restates FFAT_Map<T,1>::Intersect (reference ffat_solver.h:676-712) and ::Interpolate (:736-803) in FP64 with the
// reference's IEEE behaviour (explicit ternaries instead of fmin/fmax so NaN/inf propagate the same way).
#pragma once
#include <cuda_runtime.h>

namespace pbso {
struct Geo {           // one map's geometry in registers / local
    double cell, low[6][3], c1[3], blo[3], bhi[3], c3[3], k;
    int ne[6][2], st[6];
};
__host__ __device__ __forceinline__ void load_geo(Geo& g, const double* __restrict__ geom, const int* __restrict__ igeom) {
    g.cell = geom[0];
#pragma unroll
    for (int f = 0; f < 6; ++f)
#pragma unroll
        for (int d = 0; d < 3; ++d) g.low[f][d] = geom[1 + f * 3 + d];
#pragma unroll
    for (int d = 0; d < 3; ++d) { g.c1[d] = geom[19 + d]; g.blo[d] = geom[22 + d]; g.bhi[d] = geom[25 + d]; g.c3[d] = geom[28 + d]; }
    g.k = geom[31];
#pragma unroll
    for (int f = 0; f < 6; ++f) { g.ne[f][0] = igeom[2 * f]; g.ne[f][1] = igeom[2 * f + 1]; g.st[f] = igeom[12 + f]; }
}

// Intersect + Interpolate: returns the four Psi indices, the bilinear weights and the surface point.
// fxy (optional): {face, x, y, xp - x, yp - y} of the stencil's low corner.
__device__ __forceinline__ void ffat_locate_surf(const Geo& g, const double p[3], int idx[4], double w[4], double surf[3], int* fxy = nullptr) {
    double d[3], t_en = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d[i] = g.c1[i] - p[i];                                          // ffat_solver.h:681
        const double tmin = (g.blo[i] - p[i]) / d[i];                   // :682
        const double tmax = (g.bhi[i] - p[i]) / d[i];                   // :683
        const double te = (tmax < tmin) ? tmax : tmin;                  // :684 std::min semantics
        if (i == 0) t_en = te; else if (te > t_en) t_en = te;           // :685 maxCoeff
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) surf[i] = p[i] + t_en * d[i];           // :686
    double minDist = 1.7976931348623157e308;                            // :688
    int face = 0;
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {                                    // :689-698 strict '<', low first
        const double a = fabs(g.blo[dd] - surf[dd]);
        if (a < minDist) { minDist = a; face = dd * 2 + 1; }
        const double b = fabs(g.bhi[dd] - surf[dd]);
        if (b < minDist) { minDist = b; face = dd * 2; }
    }
    const int dk = face >> 1, di = (dk + 1) % 3, dj = (dk + 2) % 3;     // :699-702
    // (Intersect's own texel index, :706-711, is not used by GetMapVal beyond the face id.)
    const int Nx = g.ne[face][0], Ny = g.ne[face][1];
    const double h = g.cell;
    const double lowi = g.low[face][di], lowj = g.low[face][dj];
    const double xf = (surf[di] - (lowi + 0.5 * h)) / h;                // :757
    const double yf = (surf[dj] - (lowj + 0.5 * h)) / h;                // :758
    int x = (int)floor(xf), y = (int)floor(yf), xp, yp;
    double tx, ty;
    if (x < 0) { x = 0; xp = 0; tx = 0; }                               // :763-776
    else if (x < Nx - 1) { xp = x + 1; tx = xf - (double)x; }
    else { x = Nx - 1; xp = Nx - 1; tx = 0; }
    if (y < 0) { y = 0; yp = 0; ty = 0; }                               // :777-790
    else if (y < Ny - 1) { yp = y + 1; ty = yf - (double)y; }
    else { y = Ny - 1; yp = Ny - 1; ty = 0; }
    { double t = (tx < 0.0) ? 0.0 : tx; tx = (1.0 < t) ? 1.0 : t; }     // :791 min(max(tx,0),1)
    { double t = (ty < 0.0) ? 0.0 : ty; ty = (1.0 < t) ? 1.0 : t; }     // :792
    const int base = g.st[face];
    idx[0] = base + x * Ny + y;   idx[1] = base + xp * Ny + y;          // :141-144, :795-798
    idx[2] = base + x * Ny + yp;  idx[3] = base + xp * Ny + yp;
    w[0] = (1.0 - tx) * (1.0 - ty); w[1] = tx * (1.0 - ty);             // :799-802
    w[2] = (1.0 - tx) * ty;         w[3] = tx * ty;
    if (fxy) { fxy[0] = face; fxy[1] = x; fxy[2] = y; fxy[3] = xp - x; fxy[4] = yp - y; }
}
// GetMapVal's use (ffat_solver.h:1180-1206): stencil at the listener's direction, r = |p - centre|.
__device__ __forceinline__ void ffat_locate(const Geo& g, const double p[3], int idx[4], double w[4], double& r, int* fxy = nullptr) {
    double surf[3];
    ffat_locate_surf(g, p, idx, w, surf, fxy);
    const double dx = p[0] - g.c3[0], dy = p[1] - g.c3[1], dz = p[2] - g.c3[2];
    r = sqrt(dx * dx + dy * dy + dz * dz);                              // :1205 (p-_center).norm()
}

}  // namespace pbso
