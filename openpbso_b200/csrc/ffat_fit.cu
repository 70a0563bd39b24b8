// =============================================================================
// ffat_fit.cu -- FFAT cube-map CONSTRUCTION on the device (kernel K6): the step
// before the synthesis path, which turns the Dirichlet pressure an acoustic solver sampled on three
// nested cube shells into the run-time map Psi that K3 (ffat.cu) evaluates.
//
// Restates, for a batch of modes that share one shell geometry:
//   FFAT_Map<T,3>::FFAT_Map(modeId, cellSize, V, N_elements)   ffat_solver.h:944-989  (host, fitter_create)
//     -> FFAT_Map<T,1>::FFAT_Map(...)                          :399-428   low corners, strides, centre, bbox
//   FFAT_Map<T,3>::Solve(k, dirichletPressure, powerScaling)   :1007-1069
//     per direction (texel centre of shell 2) and shell: Intersect :676-712, r = |surf - centre| :1046,
//     Interpolate :736-803, p = sum_4 w * pressure(2*stride_shell + 2*quad) :1052-1057
//     -> FFAT_Solver<T,3>::Solve :872-897   least squares of [1/(k r_s)] psi = |p_s| over the shells
//     -> FFAT_Solver<T,3>::Scaling :909-929 psi *= sqrt(sum |p_0|^2 / sum (psi/(k r_0))^2)
//
// Split.  Everything that depends on geometry only -- 4 indices, 4 weights and the radius per (direction, shell)
// -- is the same for every mode: k_fit_stencil computes it once per fitter (one thread per direction) into a
// structure-of-arrays table.  k_fit_solve then streams the pressure: thread = direction (consecutive threads
// read consecutive quads), the stencil sits in registers and is reused for MPB modes, 12 x 16-byte gathers and
// one 8-byte store per (direction, mode).  HBM-bound: the reference's pressure vector carries two entries per
// quad (triangle pairs) of which one is read, so every 32-byte sector is half used; algorithmic bytes per mode
// = 16 * N_elements_total + 8 * N_directions, DRAM traffic ~ 32 * N_elements_total + 8 * N_directions.
// The power-scaling sums are reduced per block in a fixed order and finished by k_fit_scale (deterministic).
// All arithmetic is FP64.  The one-column least squares is folded into the stencil: with b_s = 1/(k r_s),
//   psi = (b . |p|) / (b . b) = k * sum_s c_s |p_s|,   c_s = (1/r_s) / sum_t (1/r_t)^2   (geometry only),
// so the per-mode work has no division (the reference's form costs S + 2 of them per direction, and FP64 division
// plus hypot made the first version of this kernel issue-bound at 0.44 of HBM peak); results differ from the
// reference's evaluation order by a few ulp (tests hold 1e-12).
// =============================================================================
#include "common.cuh"
#include "ffat_geom.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace pbso;

namespace {
constexpr int FIT_THREADS = 128;
constexpr int FIT_MAX_SHELLS = 8;
#ifndef FIT_MIN_BLOCKS
#define FIT_MIN_BLOCKS 4
#endif
struct FitGeo {                      // per shell, host + device copy
    double geom[32];
    int igeom[18];
};
}  // namespace

struct pbso_ffat_fitter {
    int device = 0;
    int S = 0, n_total = 0, n_dir = 0;
    double cell = 0;
    std::vector<FitGeo> shells;
    std::vector<int> shell_strides;
    cudaStream_t stream = nullptr;
    // stencil table, SoA over directions
    int* d_idx = nullptr;            // [S][4][n_dir]  element index into one mode's complex pressure vector
    double* d_w = nullptr;           // [S][4][n_dir]
    double* d_c = nullptr;           // [S][n_dir]  least-squares weight of shell s
    double* d_inv_r0 = nullptr;      // [n_dir]     1 / r on shell 0 (power scaling)
    // staging for the host-pointer entry
    double *d_k = nullptr, *d_p = nullptr, *d_psi = nullptr, *d_scale = nullptr, *d_partial = nullptr;
    size_t cap_maps = 0, cap_partial = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
};

// One thread per direction: the texel centre of shell 2 (ffat_solver.h:1021-1036) projected on every shell.
__global__ void __launch_bounds__(FIT_THREADS)
k_fit_stencil(int S, int n_dir, const double* __restrict__ geom, const int* __restrict__ igeom,
              const int* __restrict__ shell_strides, int* __restrict__ st_idx, double* __restrict__ st_w,
              double* __restrict__ st_c, double* __restrict__ st_inv_r0) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dir) return;
    Geo outer; load_geo(outer, geom + 2 * 32, igeom + 2 * 18);
    // direction d -> (face, ii, jj) in the enumeration order of Solve (:1021-1063): faces in order, ii outer, jj inner
    int face = 0, base = 0;
    for (int f = 0; f < 5; ++f) {
        const int n = outer.ne[f][0] * outer.ne[f][1];
        if (d < base + n) break;
        base += n; face = f + 1;
    }
    const int loc = d - base, dim2 = outer.ne[face][1];
    const int ii = loc / dim2, jj = loc - ii * dim2;
    const int dk = face >> 1, di = (dk + 1) % 3, dj = (dk + 2) % 3;
    double ijk[3]; ijk[dk] = 0.0; ijk[di] = 0.5 + (double)ii; ijk[dj] = 0.5 + (double)jj;     // :1031-1034
    double pos0[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) pos0[a] = outer.low[face][a] + ijk[a] * outer.cell;         // :1035-1036
    double inv_r[FIT_MAX_SHELLS], sum2 = 0.0;
    for (int s = 0; s < S; ++s) {
        Geo g; load_geo(g, geom + (size_t)s * 32, igeom + (size_t)s * 18);
        int idx[4]; double w[4], surf[3];
        ffat_locate_surf(g, pos0, idx, w, surf);                                             // :1043, :1051
        const double dx = surf[0] - outer.c1[0], dy = surf[1] - outer.c1[1], dz = surf[2] - outer.c1[2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);                                  // :1046, _center = shell 2's (:982)
        inv_r[s] = 1.0 / r; sum2 += inv_r[s] * inv_r[s];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            st_idx[((size_t)s * 4 + kk) * n_dir + d] = 2 * shell_strides[s] + 2 * idx[kk];   // :1054-1056
            st_w[((size_t)s * 4 + kk) * n_dir + d] = w[kk];
        }
    }
    for (int s = 0; s < S; ++s) st_c[(size_t)s * n_dir + d] = inv_r[s] / sum2;               // least-squares weights (:881-895)
    st_inv_r0[d] = inv_r[0];                                                                 // Scaling reads shell 0 (:918-921)
}

// |z| for the interpolated pressure: sqrt(re^2 + im^2) when that cannot over/underflow, hypot() (what std::abs of a
// complex uses, :885) otherwise.
__device__ __forceinline__ double cabs_fast(double re, double im) {
    const double s = re * re + im * im;
    if (s > 1e-280 && s < 1e280) return sqrt(s);
    return hypot(re, im);
}

// Per (direction, mode): interpolated pressure on each shell, folded least squares, optional scaling sums.
// S_T > 0: stencil held in registers across the block's modes.  S_T == 0: any shell count, stencil re-read (L2).
template <int S_T>
__global__ void __launch_bounds__(FIT_THREADS, S_T ? FIT_MIN_BLOCKS : 1)
k_fit_solve(int S_rt, int n_dir, int n_total, int n_maps, int maps_per_block, const int* __restrict__ st_idx,
            const double* __restrict__ st_w, const double* __restrict__ st_c, const double* __restrict__ st_inv_r0,
            const double* __restrict__ kvec, const double2* __restrict__ pressure, int packed, double* __restrict__ psi_out,
            int power_scaling, double2* __restrict__ partial) {
    const int S = S_T ? S_T : S_rt;
    const int d = blockIdx.x * FIT_THREADS + threadIdx.x;
    const bool live = d < n_dir;
    const int dc = live ? d : n_dir - 1;
    constexpr int SR = S_T ? S_T : 1;
    int idx[SR][4]; double w[SR][4], c[SR];
    if (S_T) {
#pragma unroll
        for (int s = 0; s < SR; ++s) {
            c[s] = st_c[(size_t)s * n_dir + dc];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                idx[s][kk] = st_idx[((size_t)s * 4 + kk) * n_dir + dc] >> packed;     // packed: one complex per quad
                w[s][kk] = st_w[((size_t)s * 4 + kk) * n_dir + dc];
            }
        }
    }
    const double inv_r0 = st_inv_r0[dc];
    const int m0 = blockIdx.y * maps_per_block, m1 = min(n_maps, m0 + maps_per_block);
    for (int m = m0; m < m1; ++m) {
        const double k = kvec[m];
        const double2* P = pressure + (size_t)m * (packed ? n_total : 2 * n_total);
        double acc = 0.0, pa0 = 0.0;
        if (S_T) {
            double2 v[SR][4];
#pragma unroll
            for (int s = 0; s < SR; ++s)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) v[s][kk] = __ldg(P + idx[s][kk]);       // all gathers in flight first
            // ... and the next mode's lines are requested now, so that its gathers find them on the way (ncu: half of all stall
            // samples sit on the first FMA behind the gathers -- the kernel waits for DRAM once per mode and warp)
            // (measured: 115.7 -> 103.4 us on the same box; two modes ahead or prefetching into L1: no further gain)
            if (m + 1 < m1) {
                const double2* Pn = P + (packed ? (size_t)n_total : (size_t)2 * n_total);
#pragma unroll
                for (int s = 0; s < SR; ++s)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) asm volatile("prefetch.global.L2 [%0];" ::"l"(Pn + idx[s][kk]));
            }
#pragma unroll
            for (int s = 0; s < SR; ++s) {
                double pre = 0.0, pim = 0.0;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { pre += w[s][kk] * v[s][kk].x; pim += w[s][kk] * v[s][kk].y; }   // ffat_solver.h:1052-1057
                const double p2 = cabs_fast(pre, pim);                                 // :885
                acc += c[s] * p2;
                if (s == 0) pa0 = p2;
            }
        } else {
            for (int s = 0; s < S; ++s) {
                double pre = 0.0, pim = 0.0;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double2 v = __ldg(P + (st_idx[((size_t)s * 4 + kk) * n_dir + dc] >> packed));
                    const double ww = st_w[((size_t)s * 4 + kk) * n_dir + dc];
                    pre += ww * v.x; pim += ww * v.y;
                }
                const double p2 = cabs_fast(pre, pim);
                acc += st_c[(size_t)s * n_dir + dc] * p2;
                if (s == 0) pa0 = p2;
            }
        }
        const double psi = k * acc;                                                    // :888-895, folded
        if (live) psi_out[(size_t)m * n_dir + d] = psi;
        if (power_scaling) {                                                           // :918-923, shell 0
            const double q = acc * inv_r0;                                             // psi / (k r_0)
            double2 t2 = live ? make_double2(pa0 * pa0, q * q) : make_double2(0.0, 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t2.x += __shfl_down_sync(0xffffffffu, t2.x, o);
                t2.y += __shfl_down_sync(0xffffffffu, t2.y, o);
            }
            // one partial per WARP, straight to global memory: no block barrier inside the mode loop (two __syncthreads per
            // mode cost 25 % of the kernel); k_fit_scale adds them in a fixed order
            if ((threadIdx.x & 31) == 0) partial[((size_t)m * gridDim.x + blockIdx.x) * (FIT_THREADS / 32) + (threadIdx.x >> 5)] = t2;
        }
    }
}

// One block per mode: finish the two sums in block order, scale = sqrt(numer/denom) (:924), Psi *= scale (:925-927).
__global__ void __launch_bounds__(256)
k_fit_scale(int n_dir, int n_blocks, const double2* __restrict__ partial, double* __restrict__ psi,
            double* __restrict__ scale_out) {
    __shared__ double s_scale;
    const int m = blockIdx.x;
    if (threadIdx.x < 32) {
        // lane j takes partials j, j + 32, ... in order, then a fixed shuffle tree: the same summation order every run
        double numer = 0.0, denom = 0.0;
        for (int b = threadIdx.x; b < n_blocks; b += 32) { const double2 t = partial[(size_t)m * n_blocks + b]; numer += t.x; denom += t.y; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            numer += __shfl_down_sync(0xffffffffu, numer, o);
            denom += __shfl_down_sync(0xffffffffu, denom, o);
        }
        if (threadIdx.x == 0) {
            s_scale = sqrt(numer / denom);
            if (scale_out) scale_out[m] = s_scale;
        }
    }
    __syncthreads();
    if (!psi) return;                                                 // deferred: the caller folds the scale in
    const double sc = s_scale;
    double* row = psi + (size_t)m * n_dir;
    int d = threadIdx.x;
    for (; d + 3 * 256 < n_dir; d += 4 * 256) {                       // four independent read-modify-writes in flight
        const double a0 = row[d], a1 = row[d + 256], a2 = row[d + 512], a3 = row[d + 768];
        row[d] = a0 * sc; row[d + 256] = a1 * sc; row[d + 512] = a2 * sc; row[d + 768] = a3 * sc;
    }
    for (; d < n_dir; d += 256) row[d] *= sc;
}

__global__ void k_fit_fill(int n, double v, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

static int launch_solve(pbso_ffat_fitter* f, int n_maps, const double* d_k, const double* d_p, int flags,
                        double* d_psi, double* d_scale, cudaStream_t s) {
    const int power_scaling = flags & PBSO_FIT_POWER_SCALING, packed = (flags & PBSO_FIT_PACKED) ? 1 : 0;
    const bool defer = (flags & PBSO_FIT_DEFER_SCALE) != 0;
    const int gx = div_up(f->n_dir, FIT_THREADS);
    if (power_scaling) {
        const size_t need = (size_t)n_maps * gx * (FIT_THREADS / 32);
        if (need > f->cap_partial) {
            cudaFree(f->d_partial);
            f->d_partial = nullptr; f->cap_partial = 0;
            PBSO_CUDA(cudaMalloc(&f->d_partial, need * sizeof(double2)));
            f->cap_partial = need;
        }
    }
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, f->device);
    const double2* P = reinterpret_cast<const double2*>(d_p);
    double2* part = reinterpret_cast<double2*>(f->d_partial);
    // (measured and removed, profiles/r1_ffat_fit.md and r2_ffat_fit.md: a TMA-staged variant -- 16-byte rows at a 32-byte pitch
    // through a shared-memory ring; one thread per (direction, shell) -- 56 registers, 36 resident warps instead of 16, 10 %
    // slower; issuing the y + 1 taps after the y taps so that they hit L1 -- no change.  ncu: L2 sector traffic is 3 x the
    // algorithmic bytes (the four taps and three shells of neighbouring directions re-read sectors that L1 does not hold
    // long enough); staging each block's three contiguous runs of quads through shared memory with cp.async, double-buffered
    // over the modes -- bit-identical, 8 % slower)
    {
        // direct gathers: enough blocks for ~16 resident CTAs on every SM before modes are folded into a block
        int mpb = 1;
        while (mpb < 8 && (long long)gx * div_up(n_maps, mpb * 2) >= (long long)sm * 16) mpb *= 2;
        const dim3 grid(gx, div_up(n_maps, mpb));
        if (f->S == 3)
            k_fit_solve<3><<<grid, FIT_THREADS, 0, s>>>(3, f->n_dir, f->n_total, n_maps, mpb, f->d_idx, f->d_w, f->d_c, f->d_inv_r0, d_k, P, packed,
                                                        d_psi, power_scaling, part);
        else
            k_fit_solve<0><<<grid, FIT_THREADS, 0, s>>>(f->S, f->n_dir, f->n_total, n_maps, mpb, f->d_idx, f->d_w, f->d_c, f->d_inv_r0, d_k, P, packed,
                                                        d_psi, power_scaling, part);
    }
    PBSO_CUDA(cudaGetLastError());
    if (power_scaling) {
        if (defer) PBSO_REQUIRE(d_scale, PBSO_ERR_INVALID, "PBSO_FIT_DEFER_SCALE needs the scale output");
        k_fit_scale<<<n_maps, defer ? 32 : 256, 0, s>>>(f->n_dir, gx * (FIT_THREADS / 32), part, defer ? nullptr : d_psi, d_scale);
        PBSO_CUDA(cudaGetLastError());
    } else if (d_scale) {
        k_fit_fill<<<div_up(n_maps, 256), 256, 0, s>>>(n_maps, 1.0, d_scale);
        PBSO_CUDA(cudaGetLastError());
    }
    return PBSO_OK;
}

extern "C" {

int pbso_ffat_fitter_create(double cell_size, const double* V, int n_rows, const int* n_elements, int n_shells,
                            pbso_ffat_fitter** out) {
    PBSO_REQUIRE(out, PBSO_ERR_INVALID, "null output handle");
    *out = nullptr;
    PBSO_REQUIRE(V && n_elements, PBSO_ERR_INVALID, "null argument");
    // the reference asserts N_shells > 1 (:954) and then reads _shells.at(2) (:982): three shells are the minimum
    PBSO_REQUIRE(n_shells >= 3 && n_shells <= FIT_MAX_SHELLS, PBSO_ERR_INVALID, "need 3..8 shells (shell 2 is the run-time map)");
    PBSO_REQUIRE(cell_size > 0.0, PBSO_ERR_INVALID, "cell size must be positive");
    long long quads = 0;
    for (int i = 0; i < n_shells * 12; ++i) PBSO_REQUIRE(n_elements[i] > 0, PBSO_ERR_INVALID, "N_elements must be positive");
    for (int s = 0; s < n_shells; ++s)
        for (int f = 0; f < 6; ++f) quads += (long long)n_elements[(s * 6 + f) * 2] * n_elements[(s * 6 + f) * 2 + 1];
    PBSO_REQUIRE(quads * 4 <= (long long)n_rows, PBSO_ERR_INVALID, "V has fewer than 4 rows per quad (V.block, ffat_solver.h:971)");
    PBSO_REQUIRE(quads < (1ll << 28), PBSO_ERR_INVALID, "too many quads");
    if (int rc = check_device()) return rc;

    pbso_ffat_fitter* f = new pbso_ffat_fitter();
    f->S = n_shells; f->cell = cell_size;
    f->shells.resize(n_shells); f->shell_strides.resize(n_shells);
    int total = 0; size_t row = 0;
    for (int s = 0; s < n_shells; ++s) {
        FitGeo& g = f->shells[s];
        double* low = g.geom + 1;
        int sum = 0;
        for (int fc = 0; fc < 6; ++fc) {                                        // ffat_solver.h:407-416
            const int nx = n_elements[(s * 6 + fc) * 2], ny = n_elements[(s * 6 + fc) * 2 + 1];
            for (int d = 0; d < 3; ++d) low[fc * 3 + d] = V[(row + (size_t)sum * 4) * 3 + d];
            g.igeom[2 * fc] = nx; g.igeom[2 * fc + 1] = ny; g.igeom[12 + fc] = sum;
            sum += nx * ny;
        }
        g.geom[0] = cell_size;
        g.geom[19] = (low[0 * 3 + 0] + low[1 * 3 + 0]) / 2.0;                   // :419-422
        g.geom[20] = (low[2 * 3 + 1] + low[3 * 3 + 1]) / 2.0;
        g.geom[21] = (low[4 * 3 + 2] + low[5 * 3 + 2]) / 2.0;
        // :423-428 take min/max of the six low corners against UNINITIALISED members (undefined behaviour in the
        // reference).  Here the bounds start from the first corner: the evident intent, and identical to a
        // zero-filled start whenever the box straddles the origin.
        for (int j = 0; j < 3; ++j) {
            double lo = low[j], hi = low[j];
            for (int fc = 1; fc < 6; ++fc) { lo = std::min(lo, low[fc * 3 + j]); hi = std::max(hi, low[fc * 3 + j]); }
            g.geom[22 + j] = lo; g.geom[25 + j] = hi;
        }
        g.geom[31] = -1.0;                                                      // _k before Solve (:268)
        f->shell_strides[s] = total;                                            // :963
        total += sum; row += (size_t)sum * 4;                                   // :964, :976
    }
    for (int s = 0; s < n_shells; ++s) std::memcpy(f->shells[s].geom + 28, f->shells[2].geom + 19, 3 * sizeof(double));   // :982
    f->n_total = total;
    f->n_dir = 0;
    for (int fc = 0; fc < 6; ++fc) f->n_dir += f->shells[2].igeom[2 * fc] * f->shells[2].igeom[2 * fc + 1];            // :983-986

    auto fail = [&](int rc) { pbso_ffat_fitter_destroy(f); return rc; };
    if (cudaGetDevice(&f->device) != cudaSuccess) return fail(set_error(PBSO_ERR_CUDA, "cudaGetDevice failed"));
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_error(PBSO_ERR_CUDA, "stream creation failed"));
    cudaEventCreate(&f->ev0); cudaEventCreate(&f->ev1);
    std::vector<double> geom((size_t)n_shells * 32); std::vector<int> igeom((size_t)n_shells * 18);
    for (int s = 0; s < n_shells; ++s) { std::memcpy(&geom[(size_t)s * 32], f->shells[s].geom, sizeof(double) * 32); std::memcpy(&igeom[(size_t)s * 18], f->shells[s].igeom, sizeof(int) * 18); }
    double* d_geom = nullptr; int *d_igeom = nullptr, *d_ss = nullptr;
    const size_t nd = (size_t)f->n_dir;
    cudaError_t e = cudaMalloc(&d_geom, geom.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_igeom, igeom.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_ss, n_shells * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_idx, nd * 4 * n_shells * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_w, nd * 4 * n_shells * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_c, nd * n_shells * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_inv_r0, nd * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_geom, geom.data(), geom.size() * sizeof(double), cudaMemcpyHostToDevice, f->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_igeom, igeom.data(), igeom.size() * sizeof(int), cudaMemcpyHostToDevice, f->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ss, f->shell_strides.data(), n_shells * sizeof(int), cudaMemcpyHostToDevice, f->stream);
    if (e == cudaSuccess) {
        k_fit_stencil<<<div_up(f->n_dir, FIT_THREADS), FIT_THREADS, 0, f->stream>>>(n_shells, f->n_dir, d_geom, d_igeom, d_ss, f->d_idx, f->d_w, f->d_c, f->d_inv_r0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
    cudaFree(d_geom); cudaFree(d_igeom); cudaFree(d_ss);
    if (e != cudaSuccess) return fail(set_error(PBSO_ERR_CUDA, "FFAT fitter setup failed: %s", cudaGetErrorString(e)));
    *out = f;
    return PBSO_OK;
}

int pbso_ffat_fitter_destroy(pbso_ffat_fitter* f) {
    if (!f) return PBSO_OK;
    if (f->stream) {
        DeviceGuard g(f->device);
        cudaStreamSynchronize(f->stream);
        cudaFree(f->d_idx); cudaFree(f->d_w); cudaFree(f->d_c); cudaFree(f->d_inv_r0);
        cudaFree(f->d_k); cudaFree(f->d_p); cudaFree(f->d_psi); cudaFree(f->d_scale); cudaFree(f->d_partial);
        if (f->ev0) cudaEventDestroy(f->ev0);
        if (f->ev1) cudaEventDestroy(f->ev1);
        cudaStreamDestroy(f->stream);
    }
    delete f;
    return PBSO_OK;
}

int pbso_ffat_fitter_info(const pbso_ffat_fitter* f, int* n_shells, int* n_elements_total, int* n_directions,
                          int* shell_strides) {
    PBSO_REQUIRE(f, PBSO_ERR_INVALID, "null handle");
    if (n_shells) *n_shells = f->S;
    if (n_elements_total) *n_elements_total = f->n_total;
    if (n_directions) *n_directions = f->n_dir;
    if (shell_strides) std::memcpy(shell_strides, f->shell_strides.data(), sizeof(int) * f->S);
    return PBSO_OK;
}

int pbso_ffat_fitter_shell(const pbso_ffat_fitter* f, int shell, double* geom32, int* igeom18) {
    PBSO_REQUIRE(f, PBSO_ERR_INVALID, "null handle");
    if (shell < 0 || shell >= f->S) return set_error(PBSO_ERR_RANGE, "shell %d out of range (_shells.at, ffat_solver.h:1043)", shell);
    if (geom32) std::memcpy(geom32, f->shells[shell].geom, sizeof(double) * 32);
    if (igeom18) std::memcpy(igeom18, f->shells[shell].igeom, sizeof(int) * 18);
    return PBSO_OK;
}

int pbso_ffat_fitter_solve_device(pbso_ffat_fitter* f, int n_maps, const double* d_k, const double* d_pressure,
                                  int flags, double* d_psi, double* d_scale, void* cuda_stream) {
    PBSO_REQUIRE(f && n_maps >= 0, PBSO_ERR_INVALID, "bad argument");
    if (n_maps == 0) return PBSO_OK;
    PBSO_REQUIRE(d_k && d_pressure && d_psi, PBSO_ERR_INVALID, "null argument");
    PBSO_REQUIRE((reinterpret_cast<uintptr_t>(d_pressure) & 15) == 0, PBSO_ERR_INVALID, "pressure must be 16-byte aligned");
    DeviceGuard g(f->device);
    return launch_solve(f, n_maps, d_k, d_pressure, flags, d_psi, d_scale, cuda_stream ? (cudaStream_t)cuda_stream : f->stream);
}

int pbso_ffat_fitter_solve(pbso_ffat_fitter* f, int n_maps, const double* k, const double* pressure, int flags,
                           double* psi, double* scale) {
    PBSO_REQUIRE(f && n_maps >= 0, PBSO_ERR_INVALID, "bad argument");
    if (n_maps == 0) return PBSO_OK;
    PBSO_REQUIRE(k && pressure && psi, PBSO_ERR_INVALID, "null argument");
    // Solve asserts _N_directions > 0 (:1012); k == 0 would divide by zero exactly as the reference does (inf/nan out)
    DeviceGuard g(f->device);
    const size_t per_map = (size_t)((flags & PBSO_FIT_PACKED) ? 2 : 4) * f->n_total;   // doubles: 2 * n_total complex entries (:1013), half of that packed
    const size_t chunk = std::max<size_t>(1, std::min<size_t>((size_t)n_maps, ((size_t)256 << 20) / (per_map * sizeof(double))));
    if (chunk > f->cap_maps) {
        cudaFree(f->d_k); cudaFree(f->d_p); cudaFree(f->d_psi); cudaFree(f->d_scale);
        f->d_k = f->d_p = f->d_psi = f->d_scale = nullptr; f->cap_maps = 0;
        PBSO_CUDA(cudaMalloc(&f->d_k, chunk * sizeof(double)));
        PBSO_CUDA(cudaMalloc(&f->d_p, chunk * per_map * sizeof(double)));
        PBSO_CUDA(cudaMalloc(&f->d_psi, chunk * f->n_dir * sizeof(double)));
        PBSO_CUDA(cudaMalloc(&f->d_scale, chunk * sizeof(double)));
        f->cap_maps = chunk;
    }
    f->last_ms = 0.f;
    for (size_t m0 = 0; m0 < (size_t)n_maps; m0 += chunk) {
        const int n = (int)std::min(chunk, (size_t)n_maps - m0);
        PBSO_CUDA(cudaMemcpyAsync(f->d_k, k + m0, n * sizeof(double), cudaMemcpyHostToDevice, f->stream));
        PBSO_CUDA(cudaMemcpyAsync(f->d_p, pressure + m0 * per_map, n * per_map * sizeof(double), cudaMemcpyHostToDevice, f->stream));
        PBSO_CUDA(cudaEventRecord(f->ev0, f->stream));
        if (int rc = launch_solve(f, n, f->d_k, f->d_p, flags, f->d_psi, f->d_scale, f->stream)) return rc;
        PBSO_CUDA(cudaEventRecord(f->ev1, f->stream));
        PBSO_CUDA(cudaMemcpyAsync(psi + m0 * f->n_dir, f->d_psi, (size_t)n * f->n_dir * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
        if (scale) PBSO_CUDA(cudaMemcpyAsync(scale + m0, f->d_scale, n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
        PBSO_CUDA(cudaStreamSynchronize(f->stream));
        float ms = 0.f; cudaEventElapsedTime(&ms, f->ev0, f->ev1); f->last_ms += ms;
    }
    return PBSO_OK;
}

int pbso_ffat_fitter_last_kernel_ms(const pbso_ffat_fitter* f, float* ms) {
    PBSO_REQUIRE(f && ms, PBSO_ERR_INVALID, "null argument");
    *ms = f->last_ms;
    return PBSO_OK;
}

}  // extern "C"
