// =============================================================================
// ffat.cu -- FFAT cube-map set on the device + kernel K3 (transfer evaluation).
//
// Restates, per (mode m, listener l):   ModalSolver::computeTransfer   modal_solver.h:286-315
//   FFAT_Map<T,3>::GetMapVal   ffat_solver.h:1180-1206
//     -> FFAT_Map<T,1>::Intersect     :676-712   ray from listener toward centre vs bbox slabs
//     -> FFAT_Map<T,1>::Interpolate   :736-803   texel-centre bilinear indices + weights
//     -> GetDataQuadStride            :141-144   strides[f] + x*Ny + y
//     -> FFAT_Solver<T,3>::Reconstruct:899-906   |psi / (k r)|
// All geometry runs in FP64 and reproduces the reference's IEEE behaviour (explicit ternaries
// instead of fmin/fmax so NaN/inf propagate the same way).
//
// HBM layout.  geom[n][32] / igeom[n][18] records (see pbso_b200.h), Psi column 0 stored twice:
//   psi_mm [n][D]   mode-major  -- general path, one map per thread block column
//   psi_tm [D][n]   texel-major -- used when every map shares bit-identical geometry (checked at
//                    build time like FFAT_Map_Serialize_Double::Check, ffat_map_serialize.h:281-306):
//                    a listener then hits the SAME four texels in every map, so the gather becomes
//                    four fully coalesced row reads of n doubles and the geometry is solved once per
//                    listener instead of once per (listener, mode).
// =============================================================================
#include "common.cuh"
#include "fatcube_codec.h"
#include "ffat_geom.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <vector>

using namespace pbso;

namespace {
struct HostMap {
    double geom[32];
    int igeom[18];
    bool is_compressed = false;
    std::vector<std::vector<double>> psi;     // columns
    int modeid = 0;
    // FFAT_Map<T,3>::Compress done in memory (ffat_solver.h:1125-1178): _Psi stays, _compressed_Psi is added.  The
    // compressed view is kept as what the reference derives it from -- one byte per texel and maxAmp/255 per face
    // (cpsi[i] == (double)q8[i] * q8_scale[face], bit for bit) -- and the byte table is what the device reads.
    std::vector<uint8_t> q8;
    double q8_scale[6] = {0, 0, 0, 0, 0, 0}, q8_amp[6] = {0, 0, 0, 0, 0, 0};   // maxAmp/255. and maxAmp per face
    bool has_plain() const { return !is_compressed || !q8.empty(); }
    std::vector<double> compressed_column() const {
        if (q8.empty()) return psi[0];
        std::vector<double> c(q8.size(), 0.0);
        size_t off = 0;
        for (int fc = 0; fc < 6; ++fc) {
            const size_t n = (size_t)igeom[2 * fc] * igeom[2 * fc + 1];
            for (size_t i = 0; i < n && off + i < c.size(); ++i) c[off + i] = (double)q8[off + i] * q8_scale[fc];   // data_s -> CV_64F, *= maxAmp/255.
            off += n;
        }
        return c;
    }
};
}  // namespace

struct pbso_ffat {
    int device = 0;
    std::map<int, HostMap> maps;               // keyed by modeId like LoadAll (:267-279)
    // device mirror of ids [0, n_dense): built lazily by ensure_device()
    bool dirty = true;
    int n_dense = 0;                           // largest n such that ids 0..n-1 all exist
    int D = 0;                                 // psi length when uniform, else 0
    bool shared_geom = false;
    double* d_geom = nullptr; int* d_igeom = nullptr;
    double* d_psi_mm = nullptr; double* d_psi_tm = nullptr;
    size_t* d_psi_off = nullptr;               // per-map offset into d_psi_mm
    // 8-bit view (present when every map of [0, n_dense) went through Compress in memory): one byte per texel in the
    // same two layouts + maxAmp/255 per (map, face)
    uint8_t* d_q8_mm = nullptr; uint8_t* d_q8_tm = nullptr; double* d_q8_scale = nullptr;   // scale: [6][n]
    double* d_q8_sk = nullptr;                 // [6][q8_stride]: (maxAmp/255) / k per (face, map), for k_ffat_gather_q8x4
    int q8_stride = 0;                         // maps per row of d_q8_tm / d_q8_scale / d_q8_sk: n_dense rounded up to 4
    bool q8_ready = false;
    int q8x4_ctas = 0;                         // resident CTAs per SM of k_ffat_gather_q8x4 (occupancy query, once)
    cudaStream_t stream = nullptr;
    double* d_pos = nullptr; double* d_out = nullptr; size_t pos_cap = 0, out_cap = 0;
    void* d_loc = nullptr; size_t loc_cap = 0;          // per-listener stencils (shared-geometry path)
    void* d_tile_rec = nullptr; size_t tile_rec_cap = 0;    // [n_tiles][L] stencil records binned by texel tile (texel-tile path)
    int* d_tile_order = nullptr;                            // tiles by decreasing solid angle (listeners per tile, for directions uniform on the sphere)
    double* d_psi_tiles = nullptr;                          // [slab][tile][FT_H*FT_H + 1][FT_MS]: each work item's texels contiguous (built on first use)
    bool tiles_attr_set = false, staged_attr_set = false;   // per handle = per device: function attributes are per device
    int* d_tile_cnt = nullptr; int cnt_parity = 0;          // [2][n_tiles] ping-pong counters: a call fills one, zeroes the other
    int tiles_y[6] = {0}, tile_base[7] = {0};           // texel tiles of FT_T x FT_T per face (shared geometry)
    int n_uncompressed = 0, n_compressed = 0;
    int sm_count = 148;           // leading maps with is_compressed == false / true
};

// One texel of a map: a stored double, or (8-bit view) the byte Compress kept times the face's maxAmp/255 -- rounded on
// its own (__dmul_rn), so that the value is the double the reference would have stored in _compressed_Psi.
__device__ __forceinline__ double psi_val(const double* t, size_t i, double) { return t[i]; }
__device__ __forceinline__ double psi_val(const uint8_t* t, size_t i, double scale) { return __dmul_rn((double)t[i], scale); }
// the six face scales of one map ride in registers; the face is uniform over a warp (one listener at a time)
struct FaceScale {
    double s0, s1, s2, s3, s4, s5;
    // table is face-major, [6][n]: a warp's 32 modes read 32 consecutive doubles per face
    __device__ __forceinline__ void load(const double* p, int n) { s0 = p[0]; s1 = p[n]; s2 = p[2 * n]; s3 = p[3 * n]; s4 = p[4 * n]; s5 = p[5 * n]; }
    __device__ __forceinline__ double at(int f) const { return f == 0 ? s0 : f == 1 ? s1 : f == 2 ? s2 : f == 3 ? s3 : f == 4 ? s4 : s5; }
};

// General path: one thread per (mode, listener); per-map geometry.
template <typename PT>
__global__ void __launch_bounds__(128)
k_ffat_eval_general(int n_modes, int L, const double* __restrict__ geom, const int* __restrict__ igeom,
                    const PT* __restrict__ psi, const size_t* __restrict__ psi_off, const double* __restrict__ q8_scale, int n_scale,
                    const double* __restrict__ pos, double* __restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_modes) return;
    Geo g; load_geo(g, geom + (size_t)m * 32, igeom + (size_t)m * 18);
    const PT* P = psi + psi_off[m];
    FaceScale fs = {};
    if (sizeof(PT) == 1) fs.load(q8_scale + m, n_scale);
    for (int l = blockIdx.y; l < L; l += gridDim.y) {                       // gridDim.y is capped at 65535
        const double p[3] = {pos[3 * l], pos[3 * l + 1], pos[3 * l + 2]};
        int idx[4], fxy[5]; double w[4], r;
        ffat_locate(g, p, idx, w, r, fxy);
        const double sc = sizeof(PT) == 1 ? fs.at(fxy[0]) : 0.0;
        double psi0 = 0.0;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) psi0 += w[kk] * psi_val(P, idx[kk], sc);   // :1198-1204
        out[(size_t)l * n_modes + m] = fabs(psi0 / (g.k * r));              // :904-905 + :295 std::abs
    }
}

// Shared-geometry path, two launches.
//  k_ffat_locate: one thread per listener solves the ray/box intersection and the bilinear stencil ONCE
//                 ("interpolates once per buffer per listener") -> loc[l] = {idx[4], w[4], r}
//  k_ffat_gather: block = 256 modes x FG_LPB listeners; the stencil sits in shared memory, every thread gathers its
//                 mode's four texels from the texel-major table -- a warp reads 32 consecutive doubles per texel
//                 row, fully coalesced -- and writes out[l][m] coalesced.  k differs per mode (geom[m][31]).
constexpr int FL_THREADS = 64;          // listeners per CTA of k_ffat_locate (latency-bound; 32 measured the same)
struct __align__(16) FfatLoc { int idx[4]; double w[4]; double r; int tile; int lxy; };   // 64 B: int4 + 3 x double2 loads
// lxy: position of the stencil's low corner inside its texel tile and the clamp flags: lx | ly << 4 | (xp - x) << 8 | (yp - y) << 9,
// and the cube face the stencil lies on in bits 12-14 (the 8-bit view scales per face)
constexpr int FT_T = 8;                 // texel tile edge
struct TileTable { int tiles_y[6]; int tile_base[7]; };
// What k_ffat_tiles needs of one listener, stored in its tile's bin: bilinear weights, 1/r, listener id, lxy.
struct __align__(16) TileRec { double w[4]; double inv_r; int l; int lxy; };   // 48 B

__global__ void __launch_bounds__(FL_THREADS)
k_ffat_locate(int L, const __grid_constant__ Geo g,        // the shared geometry rides in the parameter bank: no global round trip
              const double* __restrict__ pos, FfatLoc* __restrict__ loc, TileTable tt, TileRec* __restrict__ tile_rec,
              int* __restrict__ cnt_cur, int* __restrict__ cnt_next, bool inv_r) {
    // programmatic dependent launch: let k_ffat_tiles start staging its texel tiles now; it waits (griddepcontrol.wait)
    // for this grid to finish before it reads the bins
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile_rec && l <= tt.tile_base[6]) cnt_next[l] = 0;                // the NEXT call's counters (ping-pong, no extra launch)
    if (l >= L) return;
    const double p[3] = {pos[3 * l], pos[3 * l + 1], pos[3 * l + 2]};
    FfatLoc o; int fxy[5];
    ffat_locate(g, p, o.idx, o.w, o.r, fxy);
    o.tile = tt.tile_base[fxy[0]] + (fxy[1] / FT_T) * tt.tiles_y[fxy[0]] + fxy[2] / FT_T;
    o.lxy = (fxy[1] % FT_T) | (fxy[2] % FT_T) << 4 | fxy[3] << 8 | fxy[4] << 9 | fxy[0] << 12;
    if (tile_rec) {                                                      // bin by tile; order within a bin is irrelevant
        TileRec t; t.w[0] = o.w[0]; t.w[1] = o.w[1]; t.w[2] = o.w[2]; t.w[3] = o.w[3]; t.inv_r = 1.0 / o.r; t.l = l; t.lxy = o.lxy;
        // warp-aggregated: neighbouring listeners mostly fall in the same tile, and ~100 listeners per tile returning atomics to
        // one address were a third of this kernel's time (ncu source view: 37 % of the stall samples behind the atomic) -- the
        // lanes of a warp that share a tile send ONE atomic and take consecutive slots
        const unsigned live = __activemask();
        const unsigned peers = __match_any_sync(live, o.tile);
        const int leader = __ffs(peers) - 1, rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
        int base = 0;
        if ((threadIdx.x & 31) == leader) base = atomicAdd(&cnt_cur[o.tile], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        tile_rec[(size_t)o.tile * L + base + rank] = t;
    } else {
        if (inv_r) o.r = 1.0 / o.r;                                      // k_ffat_gather_q8x4 multiplies
        loc[l] = o;
    }
}

constexpr int FG_LPB = 8;
constexpr int FG_LPB_Q8 = 16;           // byte view: the six face scales a thread preloads are amortised over more listeners
template <typename PT, int LPB>
__global__ void __launch_bounds__(256)
k_ffat_gather(int n_modes, int L, const double* __restrict__ geom, const PT* __restrict__ psi_tm, int n_stride,
              const double* __restrict__ q8_scale, const FfatLoc* __restrict__ loc, double* __restrict__ out) {
    __shared__ FfatLoc s_loc[LPB];
    const int l0 = blockIdx.y * LPB;
    if (threadIdx.x < LPB && l0 + threadIdx.x < L) s_loc[threadIdx.x] = loc[l0 + threadIdx.x];
    __syncthreads();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_modes) return;
    const double k = geom[(size_t)m * 32 + 31];
    FaceScale fs = {};
    if (sizeof(PT) == 1) fs.load(q8_scale + m, n_stride);
#pragma unroll
    for (int i = 0; i < LPB; ++i) {
        if (l0 + i >= L) break;
        const FfatLoc& q = s_loc[i];
        const double sc = sizeof(PT) == 1 ? fs.at((q.lxy >> 12) & 7) : 0.0;
        double psi0 = 0.0;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) psi0 += q.w[kk] * psi_val(psi_tm, (size_t)q.idx[kk] * n_stride + m, sc);    // ffat_solver.h:1198-1204
        out[(size_t)(l0 + i) * n_modes + m] = fabs(psi0 / (k * q.r));                                  // :904-905
    }
}

// Byte view, many listeners: FOUR maps per thread.  With the table at one byte per texel the per-listener gather is no longer
// bound by L2 bytes (6 MB table, L2-resident) but by instructions issued (ncu, profiles/r2_ffat_u8.md: 86 per warp and
// listener in the one-map-per-thread kernel, a third of them the FP64 division), so here a thread loads its four maps' texels
// of one tap as ONE 32-bit word, turns bytes into doubles with a DADD (2^52 + q, minus 2^52 -- exact, and on the FP64 pipe
// instead of the quarter-rate conversion unit), and scales by (maxAmp/255)/k and 1/r -- both rounded once, on the host and
// by the staging thread -- instead of dividing: out = |sum_k w_k q_k| * (s/k) * (1/r), within 2 ulp of the division form.
// Stores are two 16-byte runs per thread, 1 KB contiguous per warp.  Starts under k_ffat_locate (programmatic dependent
// launch) and waits for its stencils.
__device__ __forceinline__ double u8_to_f64(unsigned word, int j) {     // byte j of word (one PRMT) -> double
    return __hiloint2double(0x43300000, (int)__byte_perm(word, 0u, 0x4440u + j)) - 4503599627370496.0;
}
__device__ __forceinline__ double abs_bits(double x) { return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x)); }
template <int U, bool SPLIT>
__global__ void __launch_bounds__(256)
k_ffat_gather_q8x4(int n_modes, int L, const uint8_t* __restrict__ q8_tm, int n_stride, const double* __restrict__ sk,
                   const FfatLoc* __restrict__ loc, double* __restrict__ out) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // SPLIT (n_modes a multiple of 128): a lane owns maps {2 i, 2 i + 1} and {64 + 2 i, 65 + 2 i} of its warp's 128, so
    // that each 16-byte store instruction of the warp covers 512 contiguous bytes (whole 32-byte sectors); otherwise four
    // consecutive maps (each store instruction writes half of every sector it touches)
    const int m4 = SPLIT ? (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 4 + 2 * (threadIdx.x & 31) : (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    constexpr int HI = SPLIT ? 64 : 2;     // second pair of maps, relative to the first
    if (m4 >= n_modes) return;
    const uint8_t* tab = q8_tm + m4;
    const double* skm = sk + m4;
    auto tap = [&](int texel) -> unsigned {
        const uint8_t* p = tab + (size_t)texel * n_stride;
        if (SPLIT) return (unsigned)__ldg(reinterpret_cast<const unsigned short*>(p)) | (unsigned)__ldg(reinterpret_cast<const unsigned short*>(p + 64)) << 16;
        return __ldg(reinterpret_cast<const unsigned*>(p));
    };
    // listeners are dealt round-robin to the rows of the grid (a fixed grid of a few CTAs per SM: every SM gets the same
    // share, whatever L is); a listener's 64-byte stencil is read by all threads of the block from the same address
    // (one broadcast transaction per warp and 16 bytes) -- no staging, no barrier, iterations independent
    // U listeners in flight per thread: all loads of a group are issued before any arithmetic
    for (int l0 = blockIdx.y; l0 < L; l0 += U * gridDim.y) {
        int4 idx[U]; double2 w01[U], w23[U], rt[U]; unsigned t[U][4]; double2 sk01[U], sk23[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const FfatLoc* q = loc + min(l0 + u * (int)gridDim.y, L - 1);                  // past the end: re-read the last one, store nothing
            idx[u] = __ldg(reinterpret_cast<const int4*>(q));
            w01[u] = __ldg(reinterpret_cast<const double2*>(q) + 1); w23[u] = __ldg(reinterpret_cast<const double2*>(q) + 2);
            rt[u] = __ldg(reinterpret_cast<const double2*>(q) + 3);                        // 1/r (k_ffat_locate, inv_r = true) | tile, lxy
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            t[u][0] = tap(idx[u].x); t[u][1] = tap(idx[u].y); t[u][2] = tap(idx[u].z); t[u][3] = tap(idx[u].w);
            const double* skp = skm + (size_t)((__double2hiint(rt[u].y) >> 12) & 7) * n_stride;
            sk01[u] = __ldg(reinterpret_cast<const double2*>(skp)); sk23[u] = __ldg(reinterpret_cast<const double2*>(skp + HI));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int l = l0 + u * (int)gridDim.y;
            double a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                                             // ffat_solver.h:1198-1204
                a[j] = w01[u].x * u8_to_f64(t[u][0], j);
                a[j] += w01[u].y * u8_to_f64(t[u][1], j);
                a[j] += w23[u].x * u8_to_f64(t[u][2], j);
                a[j] += w23[u].y * u8_to_f64(t[u][3], j);
            }
            if (l < L) {
                double* o = out + (size_t)l * n_modes + m4;
                *reinterpret_cast<double2*>(o) = make_double2(abs_bits(a[0] * sk01[u].x * rt[u].x), abs_bits(a[1] * sk01[u].y * rt[u].x));      // :904-905, :295
                *reinterpret_cast<double2*>(o + HI) = make_double2(abs_bits(a[2] * sk23[u].x * rt[u].x), abs_bits(a[3] * sk23[u].y * rt[u].x));
            }
        }
    }
}

// Few listeners (L <= FG_FUSED_MAX: the per-buffer cases cfg2 / cfg4): one launch.  The first FG_LPB threads of every block
// solve their listeners' ray/box + bilinear stencil themselves (geometry from the parameter bank) instead of reading it from
// a k_ffat_locate launch: a few microseconds of redundant FP64 per block buy back a dependent kernel launch, which is what
// a 2.6 MB problem costs most.
constexpr int FG_FUSED_MAX = 256;
template <typename PT>
__global__ void __launch_bounds__(256)
k_ffat_gather_fused(int n_modes, int L, const __grid_constant__ Geo g, const double* __restrict__ geom,
                    const PT* __restrict__ psi_tm, int n_stride, const double* __restrict__ q8_scale,
                    const double* __restrict__ pos, double* __restrict__ out) {
    __shared__ FfatLoc s_loc[FG_LPB];
    const int l0 = blockIdx.y * FG_LPB;
    if (threadIdx.x < FG_LPB && l0 + threadIdx.x < L) {
        const int l = l0 + threadIdx.x;
        const double p[3] = {pos[3 * l], pos[3 * l + 1], pos[3 * l + 2]};
        FfatLoc o; int fxy[5];
        ffat_locate(g, p, o.idx, o.w, o.r, fxy);
        o.tile = 0; o.lxy = fxy[0] << 12;
        s_loc[threadIdx.x] = o;
    }
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const double k = m < n_modes ? geom[(size_t)m * 32 + 31] : 1.0;      // in flight while the stencils are solved
    FaceScale fs = {};
    if (sizeof(PT) == 1 && m < n_modes) fs.load(q8_scale + m, n_stride);
    __syncthreads();
    if (m >= n_modes) return;
#pragma unroll
    for (int i = 0; i < FG_LPB; ++i) {
        if (l0 + i >= L) break;
        const FfatLoc& q = s_loc[i];
        const double sc = sizeof(PT) == 1 ? fs.at((q.lxy >> 12) & 7) : 0.0;
        double psi0 = 0.0;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) psi0 += q.w[kk] * psi_val(psi_tm, (size_t)q.idx[kk] * n_stride + m, sc);    // ffat_solver.h:1198-1204
        out[(size_t)(l0 + i) * n_modes + m] = fabs(psi0 / (k * q.r));                                  // :904-905
    }
}

// Many listeners, texel-stationary (default for L >= FT_MIN_L): when 4 L exceeds the texel count, gathering per
// listener re-reads every texel row from L2 several times (4 L M 8 bytes: 335 MB for 10 242 listeners x 1024 modes, the
// bound of k_ffat_gather).  Here a CTA owns one FT_T x FT_T texel tile (plus the one-texel halo the bilinear stencil
// reaches into) of FT_MS consecutive modes: the item's texels are ONE contiguous block of a tiled copy of Psi and are
// pulled into shared memory with a single bulk async copy (cp.async.bulk + mbarrier, SASS UBLKCP; 41 KB -- issuing the
// tile as 81 row copies from the texel-major table made the producer the bottleneck); k_ffat_locate has already binned the
// listeners' stencil records by the tile their stencil starts in (one global atomic each), so the tile's records are
// one contiguous run that rides in with the same barrier.  Then one warp per listener: lanes
// own four modes each, read the four texels from shared memory conflict-free, and store |psi/(k r)| as two coalesced
// 512-byte runs of out[l][*].  HBM traffic: the table once (x 81/64 for the halo) plus the output.
constexpr int FT_H = FT_T + 1;          // tile edge incl. halo
#ifndef PBSO_FT_MS
#define PBSO_FT_MS 128
#endif
constexpr int FT_MS = PBSO_FT_MS;       // modes per CTA (64 or 128): lanes own FT_MS / 32 modes
constexpr int FT_CW = 16;               // consumer warps (24 measured the same: not bound by consumer latency)
constexpr int FT_CTAS_PER_SM = 1;       // 2 x (81 KB tile + 12 KB records) of shared memory per CTA
constexpr int FT_THREADS = (FT_CW + 1) * 32;   // + one producer warp
constexpr int FT_MIN_L = 2048;
constexpr int FT_REC = 256;             // listener records staged per pass (12 KB)
constexpr int FT_BLK = (FT_H * FT_H + 1) * FT_MS;   // doubles per work item in the tiled table: texel rows + one row of 1/k

// Persistent, warp-specialised: FT_CW consumer warps + one producer warp per CTA, one CTA per SM.  The producer claims
// work items (tile, mode slab) from a global counter, stages the item's texel rows into one of two tile buffers and its
// listener records (FT_REC at a time) into one of two record buffers, all with bulk async copies that complete on
// "full" mbarriers; consumers release buffers through "empty" mbarriers.  Loads of item i+1 overlap the arithmetic and
// the output stores of item i.
__device__ __forceinline__ void ft_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
struct FtItem { int tile, m0, cnt, pad; };

__global__ void __launch_bounds__(FT_THREADS, FT_CTAS_PER_SM)
k_ffat_tiles(int n_modes, int L, int n_items, int n_tiles,
             const double* __restrict__ psi_tiles, const TileRec* __restrict__ tile_rec,
             const int* __restrict__ tile_cnt, const int* __restrict__ tile_order, int n_slabs, int* __restrict__ work_counter,
             double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char ft_smem[];
    constexpr size_t PSI_BYTES = (size_t)FT_BLK * sizeof(double);
    double* s_psi0 = reinterpret_cast<double*>(ft_smem);                                  // [2][FT_H * FT_H + 1][FT_MS]: texels, then 1/k
    TileRec* s_rec0 = reinterpret_cast<TileRec*>(ft_smem + 2 * PSI_BYTES);                // [2][FT_REC]
    __shared__ __align__(8) unsigned long long s_bar[8];    // full_psi[2], empty_psi[2], full_rec[2], empty_rec[2]
    __shared__ FtItem s_item[2];
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
    auto full_psi = [&](int b) { return bar0 + 8u * b; };
    auto empty_psi = [&](int b) { return bar0 + 16u + 8u * b; };
    auto full_rec = [&](int b) { return bar0 + 32u + 8u * b; };
    auto empty_rec = [&](int b) { return bar0 + 48u + 8u * b; };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_psi(b)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_psi(b)), "r"(FT_CW));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_rec(b)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_rec(b)), "r"(FT_CW));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == FT_CW) {
        // ------------------------------------------------------------------ producer warp
        int rec_it = 0;                                                   // record chunks issued so far
        bool dep_waited = false;
        // Work items are claimed from a global counter (listener counts per tile differ by ~5x between face centres
        // and corners, so a static split leaves a long tail).  The claim and the listener count of the NEXT item are
        // requested one iteration ahead, so neither global round trip sits on the producer's critical path.
        int item = blockIdx.x, cnt_next = 0, next_item = 0;
        if (lane == 0) next_item = atomicAdd(work_counter, 1) + gridDim.x;
        for (int it = 0;; ++it) {
            const int pb = it & 1;
            if (it >= 2) ft_wait(empty_psi(pb), (unsigned)(((it >> 1) - 1) & 1));
            if (item >= n_items) {
                if (lane == 0) {
                    s_item[pb].tile = -1;
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_psi(pb)) : "memory");
                }
                break;
            }
            // heaviest tiles first (longest-processing-time order keeps the tail short): item -> (tile rank, mode slab)
            const int slab = item % n_slabs, tile = tile_order[item / n_slabs], m0 = slab * FT_MS;
            // the item's texels (tile + halo, FT_MS modes) are one contiguous block of the tiled table: a single bulk
            // copy.  They do not depend on the listeners: on the first item the copy is requested before
            // k_ffat_locate (the grid this one is launched behind) has finished
            if (lane == 0) {
                asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(full_psi(pb)), "r"((unsigned)PSI_BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((unsigned)__cvta_generic_to_shared(s_psi0 + (size_t)pb * FT_BLK)),
                               "l"(psi_tiles + ((size_t)slab * n_tiles + tile) * FT_BLK), "r"((unsigned)PSI_BYTES), "r"(full_psi(pb)) : "memory");
            }
            if (!dep_waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); dep_waited = true; }
            int cnt = cnt_next;
            if (it == 0) { if (lane == 0) cnt = tile_cnt[tile]; cnt = __shfl_sync(0xffffffffu, cnt, 0); }
            if (lane == 0) {
                s_item[pb].tile = tile; s_item[pb].m0 = m0; s_item[pb].cnt = cnt;
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_psi(pb)) : "memory");   // + tx bytes => phase done
            }
            const TileRec* rec = tile_rec + (size_t)tile * L;
            for (int r0 = 0; r0 < cnt; r0 += FT_REC, ++rec_it) {
                const int rb = rec_it & 1, nrec = min(FT_REC, cnt - r0);
                if (rec_it >= 2) ft_wait(empty_rec(rb), (unsigned)(((rec_it >> 1) - 1) & 1));
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_rec(rb)), "r"((unsigned)(nrec * sizeof(TileRec))) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((unsigned)__cvta_generic_to_shared(s_rec0 + (size_t)rb * FT_REC)), "l"(rec + r0),
                                   "r"((unsigned)(nrec * sizeof(TileRec))), "r"(full_rec(rb)) : "memory");
                }
            }
            item = __shfl_sync(0xffffffffu, next_item, 0);                 // claimed one iteration ago
            cnt_next = 0;
            if (lane == 0) {
                if (item < n_items) cnt_next = tile_cnt[tile_order[item / n_slabs]];   // used next iteration
                next_item = atomicAdd(work_counter, 1) + gridDim.x;
            }
            cnt_next = __shfl_sync(0xffffffffu, cnt_next, 0);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const bool vec = (n_modes & 1) == 0;
    int rec_it = 0;
    for (int it = 0;; ++it) {
        const int pb = it & 1;
        ft_wait(full_psi(pb), (unsigned)((it >> 1) & 1));
        const int tile = s_item[pb].tile;
        if (tile < 0) break;
        const int m0 = s_item[pb].m0, cnt = s_item[pb].cnt;
        const int live = min(FT_MS, n_modes - m0);
        const double* s_psi = s_psi0 + (size_t)pb * FT_BLK;
        // this lane's modes: m0 + 64 h + 2 lane + {0, 1}, h < FT_MS / 64; their 1/k rides in the item's block
        double inv_k[FT_MS / 32];
#pragma unroll
        for (int j = 0; j < FT_MS / 32; ++j) inv_k[j] = s_psi[FT_H * FT_H * FT_MS + (j >> 1) * 64 + 2 * lane + (j & 1)];
        for (int r0 = 0; r0 < cnt; r0 += FT_REC, ++rec_it) {
            const int rb = rec_it & 1, nrec = min(FT_REC, cnt - r0);
            ft_wait(full_rec(rb), (unsigned)((rec_it >> 1) & 1));
            const TileRec* s_rec = s_rec0 + (size_t)rb * FT_REC;
#pragma unroll 2
            for (int i = warp; i < nrec; i += FT_CW) {
                const TileRec q = s_rec[i];
                const int lx = q.lxy & 15, ly = (q.lxy >> 4) & 15, dx = (q.lxy >> 8) & 1, dy = (q.lxy >> 9) & 1;
                const double* t00 = s_psi + (size_t)(lx * FT_H + ly) * FT_MS;
                const double* t10 = t00 + (size_t)dx * FT_H * FT_MS;                       // (xp, y)
                const double* t01 = t00 + (size_t)dy * FT_MS;                              // (x, yp)
                const double* t11 = t10 + (size_t)dy * FT_MS;                              // (xp, yp)
                double* o = out + (size_t)q.l * n_modes + m0;
#pragma unroll
                for (int h = 0; h < FT_MS / 64; ++h) {
                    const int c = h * 64 + 2 * lane;
                    const double2 a = *reinterpret_cast<const double2*>(t00 + c), b = *reinterpret_cast<const double2*>(t10 + c);
                    const double2 cc = *reinterpret_cast<const double2*>(t01 + c), d = *reinterpret_cast<const double2*>(t11 + c);
                    double v0 = 0.0, v1 = 0.0;                                             // ffat_solver.h:1198-1204, same order
                    v0 += q.w[0] * a.x; v0 += q.w[1] * b.x; v0 += q.w[2] * cc.x; v0 += q.w[3] * d.x;
                    v1 += q.w[0] * a.y; v1 += q.w[1] * b.y; v1 += q.w[2] * cc.y; v1 += q.w[3] * d.y;
                    v0 = fabs(v0 * inv_k[2 * h] * q.inv_r); v1 = fabs(v1 * inv_k[2 * h + 1] * q.inv_r);   // |psi / (k r)|, :904-905
                    if (vec && c + 1 < live) *reinterpret_cast<double2*>(o + c) = make_double2(v0, v1);
                    else { if (c < live) o[c] = v0; if (c + 1 < live) o[c + 1] = v1; }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_rec(rb)) : "memory");
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_psi(pb)) : "memory");
    }
}

// Many listeners (the 10 242-direction sphere of tools/real_time_modal_sound.cpp:921-927, or any L >> D/4):
// every texel of every map is needed, so each CTA stages FS_G whole maps (column 0 of Psi, mode-major, 48 KB each
// for 6 x 32 x 32 texels) in shared memory with bulk async copies (cp.async.bulk + mbarrier: the TMA engine, SASS
// UBLKCP) and then streams the listeners' precomputed stencils past them: four shared-memory gathers per
// (listener, map), one 32-byte store of four adjacent modes per listener.
constexpr int FS_G = 4;
constexpr int FS_THREADS = 1024;

__global__ void __launch_bounds__(FS_THREADS, 1)
k_ffat_staged(int n_modes, int L, int D, int l_split, const double* __restrict__ geom, const double* __restrict__ psi_mm,
              const FfatLoc* __restrict__ loc, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char fs_smem[];
    double* s_psi = reinterpret_cast<double*>(fs_smem);                    // [FS_G][D]
    __shared__ __align__(8) unsigned long long s_bar;
    const int m0 = blockIdx.x * FS_G;
    const int g_live = min(FS_G, n_modes - m0);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned bytes = (unsigned)(D * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * g_live) : "memory");
        for (int g = 0; g < g_live; ++g) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_psi + (size_t)g * D);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(psi_mm + (size_t)(m0 + g) * D), "r"(bytes), "r"(bar) : "memory");
        }
    }
    double inv_k[FS_G];
#pragma unroll
    for (int g = 0; g < FS_G; ++g) inv_k[g] = g < g_live ? 1.0 / geom[(size_t)(m0 + g) * 32 + 31] : 0.0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(bar) : "memory");
    const int per = (L + l_split - 1) / l_split;
    const int l_begin = blockIdx.y * per, l_end = min(L, l_begin + per);
    const bool vec4 = (g_live == FS_G) && ((n_modes & 3) == 0);
    for (int l = l_begin + threadIdx.x; l < l_end; l += FS_THREADS) {
        FfatLoc q;
        {
            const int4* p4 = reinterpret_cast<const int4*>(loc + l);
            const int4 i4 = p4[0];
            const double2 w01 = reinterpret_cast<const double2*>(p4)[1], w23 = reinterpret_cast<const double2*>(p4)[2];
            const double2 rr = reinterpret_cast<const double2*>(p4)[3];
            q.idx[0] = i4.x; q.idx[1] = i4.y; q.idx[2] = i4.z; q.idx[3] = i4.w;
            q.w[0] = w01.x; q.w[1] = w01.y; q.w[2] = w23.x; q.w[3] = w23.y; q.r = rr.x;
        }
        const double inv_r = 1.0 / q.r;
        double v[FS_G];
#pragma unroll
        for (int g = 0; g < FS_G; ++g) {
            const double* P = s_psi + (size_t)g * D;
            double psi0 = 0.0;
            if (g < g_live) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) psi0 += q.w[kk] * P[q.idx[kk]];          // ffat_solver.h:1198-1204
            }
            v[g] = fabs(psi0 * inv_k[g] * inv_r);                                        // |psi / (k r)|, :904-905
        }
        double* o = out + (size_t)l * n_modes + m0;
        if (vec4) {
            *reinterpret_cast<double2*>(o) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(o + 2) = make_double2(v[2], v[3]);
        } else {
            for (int g = 0; g < g_live; ++g) o[g] = v[g];
        }
    }
}

// ---------------------------------------------------------------------------------------------
static int to_host_map(const FatcubeMap& fm, HostMap& hm, std::string& err) {
    // sizes FFAT_Map_Serialize_Double::Load relies on (ffat_map_serialize.h:181-222, 243)
    if (fm.center1.size() != 3 || fm.bboxlow.size() != 3 || fm.bboxtop.size() != 3 || fm.center3.size() != 3) {
        err = "center/bboxlow/bboxtop must have exactly 3 items (ffat_map_serialize.h:22-28)"; return PBSO_ERR_FORMAT; }
    if (fm.lowcorners.size() != 6 || fm.n_elements.size() != 6 || fm.strides.size() != 6) {
        err = "a runtime cube map needs 6 lowcorners / n_elements / strides (ffat_solver.h:689-711)"; return PBSO_ERR_FORMAT; }
    for (auto& v : fm.lowcorners) if (v.size() < 3) { err = "lowcorners row shorter than 3"; return PBSO_ERR_FORMAT; }
    for (auto& v : fm.n_elements) if (v.size() < 2) { err = "n_elements row shorter than 2"; return PBSO_ERR_FORMAT; }
    if (fm.psi.empty()) { err = "psi has no column (Load reads psi().item(0), ffat_map_serialize.h:243)"; return PBSO_ERR_FORMAT; }
    hm.geom[0] = fm.cellsize;
    for (int f = 0; f < 6; ++f) for (int d = 0; d < 3; ++d) hm.geom[1 + 3 * f + d] = fm.lowcorners[f][d];
    for (int d = 0; d < 3; ++d) { hm.geom[19 + d] = fm.center1[d]; hm.geom[22 + d] = fm.bboxlow[d]; hm.geom[25 + d] = fm.bboxtop[d]; hm.geom[28 + d] = fm.center3[d]; }
    hm.geom[31] = fm.k;
    for (int f = 0; f < 6; ++f) { hm.igeom[2 * f] = fm.n_elements[f][0]; hm.igeom[2 * f + 1] = fm.n_elements[f][1]; hm.igeom[12 + f] = fm.strides[f]; }
    hm.is_compressed = fm.is_compressed;
    // DESERIALIZE_EMAT (:44-53): rows = psi().item(0).item_size(); every column must have that many
    const size_t rows = fm.psi[0].size();
    for (auto& c : fm.psi) if (c.size() < rows) { err = "psi columns of unequal length"; return PBSO_ERR_FORMAT; }
    hm.psi = fm.psi;
    for (auto& c : hm.psi) c.resize(rows);
    hm.modeid = fm.modeid;
    // every texel index Interpolate can produce must be inside Psi
    for (int f = 0; f < 6; ++f) {
        long long last = (long long)hm.igeom[12 + f] + (long long)hm.igeom[2 * f] * hm.igeom[2 * f + 1];
        if (hm.igeom[2 * f] <= 0 || hm.igeom[2 * f + 1] <= 0 || hm.igeom[12 + f] < 0 || last > (long long)rows) {
            err = "strides/n_elements address texels outside psi"; return PBSO_ERR_FORMAT; }
    }
    return PBSO_OK;
}

static void from_host_map(const HostMap& hm, FatcubeMap& fm) {
    fm.cellsize = hm.geom[0];
    fm.lowcorners.assign(6, std::vector<double>(3));
    fm.n_elements.assign(6, std::vector<int>(2));
    fm.strides.resize(6);
    for (int f = 0; f < 6; ++f) {
        for (int d = 0; d < 3; ++d) fm.lowcorners[f][d] = hm.geom[1 + 3 * f + d];
        fm.n_elements[f][0] = hm.igeom[2 * f]; fm.n_elements[f][1] = hm.igeom[2 * f + 1];
        fm.strides[f] = hm.igeom[12 + f];
    }
    fm.center1.assign(hm.geom + 19, hm.geom + 22);
    fm.bboxlow.assign(hm.geom + 22, hm.geom + 25);
    fm.bboxtop.assign(hm.geom + 25, hm.geom + 28);
    fm.center3.assign(hm.geom + 28, hm.geom + 31);
    fm.k = hm.geom[31];
    fm.is_compressed = hm.is_compressed;
    fm.psi = hm.psi;
    if (hm.is_compressed && !hm.q8.empty()) fm.psi.assign(1, hm.compressed_column());   // Save writes _compressed_Psi (ffat_map_serialize.h:149-153)
    fm.modeid = hm.modeid;
}

static int load_one(const char* filename, HostMap& hm) {
    std::ifstream stream(filename, std::ios::binary);
    if (!stream) return set_error(PBSO_ERR_IO, "cannot open %s", filename);
    std::stringstream ss; ss << stream.rdbuf();
    const std::string buf = ss.str();
    FatcubeMap fm; std::string err;
    // both of the reference's loaders read "*.fatcube": FFAT_Map_Serialize::Load the protobuf form (ffat_map_serialize.h:166-254),
    // FFAT_Map<T,3>::Load the legacy igl::serialize form (ffat_solver.h:1069-1071) -- told apart by the legacy chunk header
    const bool legacy = legacy_fatcube_sniff((const uint8_t*)buf.data(), buf.size());
    if (!(legacy ? legacy_fatcube_decode((const uint8_t*)buf.data(), buf.size(), fm, err)
                 : fatcube_decode((const uint8_t*)buf.data(), buf.size(), fm, err)))
        return set_error(PBSO_ERR_FORMAT, "%s: %s", filename, err.c_str());
    if (int rc = to_host_map(fm, hm, err)) return set_error(rc, "%s: %s", filename, err.c_str());
    return PBSO_OK;
}

static void free_device(pbso_ffat* f) {
    cudaFree(f->d_geom); cudaFree(f->d_igeom); cudaFree(f->d_psi_mm); cudaFree(f->d_psi_tm); cudaFree(f->d_psi_off); cudaFree(f->d_psi_tiles); cudaFree(f->d_tile_order);
    cudaFree(f->d_q8_mm); cudaFree(f->d_q8_tm); cudaFree(f->d_q8_scale); cudaFree(f->d_q8_sk);
    f->d_q8_mm = f->d_q8_tm = nullptr; f->d_q8_scale = f->d_q8_sk = nullptr; f->q8_ready = false;
    cudaFree(f->d_tile_rec); cudaFree(f->d_tile_cnt); f->d_tile_rec = nullptr; f->tile_rec_cap = 0; f->d_tile_cnt = nullptr;   // sized by the tile count
    f->d_psi_tiles = nullptr; f->d_tile_order = nullptr; f->d_geom = nullptr; f->d_igeom = nullptr; f->d_psi_mm = nullptr; f->d_psi_tm = nullptr; f->d_psi_off = nullptr;
}

static int ensure_device(pbso_ffat* f) {
    if (!f->dirty) return PBSO_OK;
    if (int rc = check_device()) return rc;
    free_device(f);
    if (!f->stream) {
        PBSO_CUDA(cudaGetDevice(&f->device));
        PBSO_CUDA(cudaDeviceGetAttribute(&f->sm_count, cudaDevAttrMultiProcessorCount, f->device));
        PBSO_CUDA(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
    }
    int n = 0;
    while (f->maps.count(n)) ++n;               // ids 0..n-1 are what computeTransfer can reach
    f->n_dense = n;
    if (n == 0) { f->dirty = false; return PBSO_OK; }
    std::vector<double> geom((size_t)n * 32); std::vector<int> igeom((size_t)n * 18);
    std::vector<size_t> off(n);
    size_t total = 0; bool uniform = true, same_geo = true;
    const HostMap& first = f->maps.at(0);
    for (int m = 0; m < n; ++m) {
        const HostMap& hm = f->maps.at(m);
        std::memcpy(&geom[(size_t)m * 32], hm.geom, sizeof(hm.geom));
        std::memcpy(&igeom[(size_t)m * 18], hm.igeom, sizeof(hm.igeom));
        off[m] = total; total += hm.psi[0].size();
        if (hm.psi[0].size() != first.psi[0].size()) uniform = false;
        // bitwise comparison of everything except k (geom[31]) -- cf. Check/MatchBits (:256-306)
        if (std::memcmp(hm.geom, first.geom, sizeof(double) * 31) != 0 ||
            std::memcmp(hm.igeom, first.igeom, sizeof(hm.igeom)) != 0) same_geo = false;
    }
    f->D = uniform ? (int)first.psi[0].size() : 0;
    f->shared_geom = uniform && same_geo;
    std::vector<double> psi(total);
    for (int m = 0; m < n; ++m) { const auto& c = f->maps.at(m).psi[0]; std::memcpy(&psi[off[m]], c.data(), c.size() * sizeof(double)); }
    PBSO_CUDA(cudaMalloc(&f->d_geom, geom.size() * sizeof(double)));
    PBSO_CUDA(cudaMalloc(&f->d_igeom, igeom.size() * sizeof(int)));
    PBSO_CUDA(cudaMalloc(&f->d_psi_off, off.size() * sizeof(size_t)));
    PBSO_CUDA(cudaMalloc(&f->d_psi_mm, std::max<size_t>(total, 1) * sizeof(double)));
    PBSO_CUDA(cudaMemcpy(f->d_geom, geom.data(), geom.size() * sizeof(double), cudaMemcpyHostToDevice));
    PBSO_CUDA(cudaMemcpy(f->d_igeom, igeom.data(), igeom.size() * sizeof(int), cudaMemcpyHostToDevice));
    PBSO_CUDA(cudaMemcpy(f->d_psi_off, off.data(), off.size() * sizeof(size_t), cudaMemcpyHostToDevice));
    PBSO_CUDA(cudaMemcpy(f->d_psi_mm, psi.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    if (f->shared_geom) {
        const int D = f->D;
        std::vector<double> tm((size_t)D * n);
        for (int m = 0; m < n; ++m) { const auto& c = f->maps.at(m).psi[0]; for (int t = 0; t < D; ++t) tm[(size_t)t * n + m] = c[t]; }
        PBSO_CUDA(cudaMalloc(&f->d_psi_tm, tm.size() * sizeof(double)));
        PBSO_CUDA(cudaMemcpy(f->d_psi_tm, tm.data(), tm.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (f->shared_geom) {
        int tb = 0;
        for (int fc = 0; fc < 6; ++fc) {
            const int nx = first.igeom[2 * fc], ny = first.igeom[2 * fc + 1];
            f->tiles_y[fc] = (ny + FT_T - 1) / FT_T;
            f->tile_base[fc] = tb;
            tb += ((nx + FT_T - 1) / FT_T) * f->tiles_y[fc];
        }
        f->tile_base[6] = tb;
    }
    // 8-bit view: every map went through Compress in memory -> byte tables in the same two layouts + maxAmp/255 per face
    bool all_q8 = true;
    for (int m = 0; m < n; ++m) all_q8 = all_q8 && f->maps.at(m).q8.size() == f->maps.at(m).psi[0].size();
    if (all_q8) {
        const int ns = (n + 3) / 4 * 4; f->q8_stride = ns;
        std::vector<uint8_t> q(total); std::vector<double> sc((size_t)ns * 6, 0.0), sk((size_t)ns * 6, 0.0);
        for (int m = 0; m < n; ++m) {
            const HostMap& hm = f->maps.at(m);
            std::memcpy(&q[off[m]], hm.q8.data(), hm.q8.size());
            for (int fc = 0; fc < 6; ++fc) { sc[(size_t)fc * ns + m] = hm.q8_scale[fc]; sk[(size_t)fc * ns + m] = hm.q8_scale[fc] / hm.geom[31]; }
        }
        PBSO_CUDA(cudaMalloc(&f->d_q8_mm, std::max<size_t>(total, 1)));
        PBSO_CUDA(cudaMemcpy(f->d_q8_mm, q.data(), total, cudaMemcpyHostToDevice));
        PBSO_CUDA(cudaMalloc(&f->d_q8_scale, sc.size() * sizeof(double)));
        PBSO_CUDA(cudaMemcpy(f->d_q8_scale, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice));
        PBSO_CUDA(cudaMalloc(&f->d_q8_sk, sk.size() * sizeof(double)));
        PBSO_CUDA(cudaMemcpy(f->d_q8_sk, sk.data(), sk.size() * sizeof(double), cudaMemcpyHostToDevice));
        if (f->shared_geom) {
            const int D = f->D;
            std::vector<uint8_t> tm((size_t)D * ns, 0);
            for (int m = 0; m < n; ++m) { const auto& c = f->maps.at(m).q8; for (int t = 0; t < D; ++t) tm[(size_t)t * ns + m] = c[t]; }
            PBSO_CUDA(cudaMalloc(&f->d_q8_tm, tm.size()));
            PBSO_CUDA(cudaMemcpy(f->d_q8_tm, tm.data(), tm.size(), cudaMemcpyHostToDevice));
        }
        f->q8_ready = true;
    }
    // leading maps whose _Psi / _compressed_Psi is there to be read (GetMapVal's two views)
    f->n_uncompressed = 0; while (f->n_uncompressed < n && f->maps.at(f->n_uncompressed).has_plain()) ++f->n_uncompressed;
    f->n_compressed = 0;
    if (f->q8_ready) f->n_compressed = n;
    else while (f->n_compressed < n && f->maps.at(f->n_compressed).is_compressed && f->maps.at(f->n_compressed).q8.empty()) ++f->n_compressed;
    f->dirty = false;
    return PBSO_OK;
}

// q8: read the byte tables of the compressed view (maps compressed in memory); otherwise the stored doubles
static int launch_eval(pbso_ffat* f, int n_modes, const double* d_pos, int L, double* d_out, cudaStream_t s, bool q8 = false) {
    if (f->shared_geom) {
        if ((size_t)L > f->loc_cap) {
            cudaFree(f->d_loc); f->d_loc = nullptr; f->loc_cap = 0;
            PBSO_CUDA(cudaMalloc(&f->d_loc, sizeof(FfatLoc) * (size_t)L));
            f->loc_cap = L;
        }
        TileTable tt;
        std::memcpy(tt.tiles_y, f->tiles_y, sizeof(tt.tiles_y)); std::memcpy(tt.tile_base, f->tile_base, sizeof(tt.tile_base));
        const char* env = getenv("PBSO_FFAT_STAGED");
        const bool staged = env && env[0] == '1' && !q8;
        const char* env_g = getenv("PBSO_FFAT_GATHER");
        const int n_tiles = f->tile_base[6];
        const size_t list_need = (size_t)n_tiles * L;
        // the byte table of 1024 maps is 6 MB and sits in L2: gathering per listener costs 1 sector per warp and tap, the
        // output write is what is left, so the 8-bit view needs no texel-stationary pass
        const bool tiles = !q8 && L >= FT_MIN_L && 4ll * L >= f->D && list_need * sizeof(TileRec) <= ((size_t)512 << 20) && !staged && !(env_g && env_g[0] == '1');
        int *cnt_cur = nullptr, *cnt_next = nullptr;
        if (tiles && !f->d_psi_tiles) {
            // tiled copy of Psi: [mode slab][tile][FT_H x FT_H texels incl. the high-side halo][FT_MS modes], zero-filled
            // outside the face / past the last mode, so that one work item is one contiguous 16-byte-aligned block
            const int n = f->n_dense, n_slabs = div_up(n, FT_MS);
            std::vector<double> tl((size_t)n_slabs * n_tiles * FT_BLK, 0.0);
            const HostMap& first = f->maps.at(0);
            for (int m = 0; m < n; ++m) {
                const std::vector<double>& c = f->maps.at(m).psi[0];
                const int slab = m / FT_MS, mm = m % FT_MS;
                for (int fc = 0; fc < 6; ++fc) {
                    const int nx = first.igeom[2 * fc], ny = first.igeom[2 * fc + 1], st = first.igeom[12 + fc];
                    const int txn = div_up(nx, FT_T);
                    for (int tx = 0; tx < txn; ++tx)
                        for (int ty = 0; ty < f->tiles_y[fc]; ++ty) {
                            const size_t item = (size_t)slab * n_tiles + f->tile_base[fc] + tx * f->tiles_y[fc] + ty;
                            for (int a = 0; a < FT_H && tx * FT_T + a < nx; ++a)
                                for (int b = 0; b < FT_H && ty * FT_T + b < ny; ++b)
                                    tl[item * FT_BLK + (size_t)(a * FT_H + b) * FT_MS + mm] = c[(size_t)st + (tx * FT_T + a) * ny + ty * FT_T + b];
                            tl[item * FT_BLK + (size_t)FT_H * FT_H * FT_MS + mm] = 1.0 / f->maps.at(m).geom[31];
                        }
                }
            }
            // work order: a tile's expected listener count for directions uniform on the sphere is its solid angle seen from
            // the centre, sum over texels of (p . n) / |p|^3 with p = texel centre - centre
            std::vector<std::pair<double, int>> wt(n_tiles);
            for (int fc = 0; fc < 6; ++fc) {
                const int nx = first.igeom[2 * fc], ny = first.igeom[2 * fc + 1];
                const int dk = fc / 2, di = (dk + 1) % 3, dj = (dk + 2) % 3;
                const double h = first.geom[0];
                for (int x = 0; x < nx; ++x)
                    for (int y = 0; y < ny; ++y) {
                        double p[3];
                        p[dk] = first.geom[1 + 3 * fc + dk] - first.geom[28 + dk];
                        p[di] = first.geom[1 + 3 * fc + di] + (x + 0.5) * h - first.geom[28 + di];
                        p[dj] = first.geom[1 + 3 * fc + dj] + (y + 0.5) * h - first.geom[28 + dj];
                        const double r2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
                        const int t = f->tile_base[fc] + (x / FT_T) * f->tiles_y[fc] + y / FT_T;
                        wt[t].first += std::fabs(p[dk]) / (r2 * std::sqrt(r2)); wt[t].second = t;
                    }
            }
            std::sort(wt.begin(), wt.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first || (a.first == b.first && a.second < b.second); });
            std::vector<int> order(n_tiles);
            for (int i = 0; i < n_tiles; ++i) order[i] = wt[i].second;
            PBSO_CUDA(cudaMalloc(&f->d_tile_order, order.size() * sizeof(int)));
            PBSO_CUDA(cudaMemcpy(f->d_tile_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
            PBSO_CUDA(cudaMalloc(&f->d_psi_tiles, tl.size() * sizeof(double)));
            PBSO_CUDA(cudaMemcpy(f->d_psi_tiles, tl.data(), tl.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        if (tiles) {
            if (list_need > f->tile_rec_cap) {
                cudaFree(f->d_tile_rec); f->d_tile_rec = nullptr; f->tile_rec_cap = 0;
                PBSO_CUDA(cudaMalloc(&f->d_tile_rec, sizeof(TileRec) * list_need));
                f->tile_rec_cap = list_need;
            }
            if (!f->d_tile_cnt) {
                PBSO_CUDA(cudaMalloc(&f->d_tile_cnt, sizeof(int) * 2 * (n_tiles + 1)));
                PBSO_CUDA(cudaMemsetAsync(f->d_tile_cnt, 0, sizeof(int) * 2 * (n_tiles + 1), s));
                f->cnt_parity = 0;
            }
            cnt_cur = f->d_tile_cnt + (size_t)f->cnt_parity * (n_tiles + 1);
            cnt_next = f->d_tile_cnt + (size_t)(f->cnt_parity ^ 1) * (n_tiles + 1);
            f->cnt_parity ^= 1;
        }
        Geo g0; load_geo(g0, f->maps.at(0).geom, f->maps.at(0).igeom);
        const bool q8x4 = q8 && L > FG_FUSED_MAX && n_modes % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0;
        if (L <= FG_FUSED_MAX && !staged) {
            const dim3 grid(div_up(n_modes, 256), div_up(L, FG_LPB));
            if (q8) k_ffat_gather_fused<uint8_t><<<grid, 256, 0, s>>>(n_modes, L, g0, f->d_geom, f->d_q8_tm, f->q8_stride, f->d_q8_scale, d_pos, d_out);
            else k_ffat_gather_fused<double><<<grid, 256, 0, s>>>(n_modes, L, g0, f->d_geom, f->d_psi_tm, f->n_dense, nullptr, d_pos, d_out);
            PBSO_CUDA(cudaGetLastError());
            return PBSO_OK;
        }
        k_ffat_locate<<<div_up(std::max(L, tiles ? n_tiles + 1 : 0), FL_THREADS), FL_THREADS, 0, s>>>(L, g0, d_pos, (FfatLoc*)f->d_loc, tt,
                                                                                tiles ? (TileRec*)f->d_tile_rec : nullptr, cnt_cur, cnt_next, q8x4);
        const size_t stage_bytes = (size_t)FS_G * f->D * sizeof(double);
        // Measured on B200 (profiles/r1_ffat.md): for 1024 maps x 10 242 listeners the staged kernel is bound by LSU
        // wavefronts (scattered 8-byte shared-memory gathers + 32-byte output segments) at ~100 us, the coalesced
        // texel-major gather by L2 bandwidth at ~54 us -- so the gather is the default and staging is opt-in.
        if (tiles) {
            const size_t tile_bytes = 2 * ((size_t)FT_BLK * sizeof(double) + (size_t)FT_REC * sizeof(TileRec));
            if (!f->tiles_attr_set) { PBSO_CUDA(cudaFuncSetAttribute(k_ffat_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes)); f->tiles_attr_set = true; }
            const int n_items = n_tiles * div_up(n_modes, FT_MS);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(std::min(n_items, FT_CTAS_PER_SM * f->sm_count)); cfg.blockDim = dim3(FT_THREADS);
            cfg.dynamicSmemBytes = tile_bytes; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;                      // may start while k_ffat_locate runs (griddepcontrol)
            PBSO_CUDA(cudaLaunchKernelEx(&cfg, k_ffat_tiles, n_modes, L, n_items, n_tiles,
                                         (const double*)f->d_psi_tiles, (const TileRec*)f->d_tile_rec, (const int*)cnt_cur, (const int*)f->d_tile_order,
                                         div_up(n_modes, FT_MS), cnt_cur + n_tiles, d_out));
        } else if (staged && L >= 1024 && (f->D % 2 == 0) && stage_bytes <= 200 * 1024) {
            // whole maps in shared memory; listeners split so that the grid is ~7 waves of SMs
            if (!f->staged_attr_set) { PBSO_CUDA(cudaFuncSetAttribute(k_ffat_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); f->staged_attr_set = true; }
            const int groups = div_up(n_modes, FS_G);
            // one pass over the listeners per staged group unless there are too few groups to occupy the SMs
            int l_split = std::max(1, std::min(div_up(f->sm_count, groups), div_up(L, FS_THREADS)));
            k_ffat_staged<<<dim3(groups, l_split), FS_THREADS, stage_bytes, s>>>(n_modes, L, f->D, l_split, f->d_geom, f->d_psi_mm,
                                                                                 (const FfatLoc*)f->d_loc, d_out);
        } else {
            if (q8x4) {
                cudaLaunchConfig_t cfg = {};
                const int gx = div_up(n_modes, 1024);
                // measured on B200 (1024 maps x 10 242 listeners, locate included): whole-sector stores 27.6 us against 35.8 us
                // with half-sector stores; 2 listeners in flight per thread and one resident wave (1 or 4 in flight, 4 waves:
                // within 2 us)
                const bool split = n_modes % 128 == 0;
                auto kern = split ? k_ffat_gather_q8x4<2, true> : k_ffat_gather_q8x4<2, false>;
                PBSO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&f->q8x4_ctas, kern, 256, 0));
                cfg.gridDim = dim3(gx, std::min(L, std::max(1, f->q8x4_ctas * f->sm_count / gx))); cfg.blockDim = dim3(256); cfg.stream = s;   // one resident wave
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                PBSO_CUDA(cudaLaunchKernelEx(&cfg, kern, n_modes, L, (const uint8_t*)f->d_q8_tm, f->q8_stride,
                                             (const double*)f->d_q8_sk, (const FfatLoc*)f->d_loc, d_out));
            } else if (q8) k_ffat_gather<uint8_t, FG_LPB_Q8><<<dim3(div_up(n_modes, 256), div_up(L, FG_LPB_Q8)), 256, 0, s>>>(n_modes, L, f->d_geom, f->d_q8_tm, f->q8_stride, f->d_q8_scale, (const FfatLoc*)f->d_loc, d_out);
            else k_ffat_gather<double, FG_LPB><<<dim3(div_up(n_modes, 256), div_up(L, FG_LPB)), 256, 0, s>>>(n_modes, L, f->d_geom, f->d_psi_tm, f->n_dense, nullptr, (const FfatLoc*)f->d_loc, d_out);
        }
    } else {
        dim3 grid(div_up(n_modes, 128), std::min(L, 65535));
        if (q8) k_ffat_eval_general<uint8_t><<<grid, 128, 0, s>>>(n_modes, L, f->d_geom, f->d_igeom, f->d_q8_mm, f->d_psi_off, f->d_q8_scale, f->q8_stride, d_pos, d_out);
        else k_ffat_eval_general<double><<<grid, 128, 0, s>>>(n_modes, L, f->d_geom, f->d_igeom, f->d_psi_mm, f->d_psi_off, nullptr, 0, d_pos, d_out);
    }
    PBSO_CUDA(cudaGetLastError());
    return PBSO_OK;
}

static int check_eval_args(pbso_ffat* f, int n_modes, int L, int use_compressed) {
    PBSO_REQUIRE(f && n_modes >= 0 && L >= 0, PBSO_ERR_INVALID, "bad argument");
    if (int rc = ensure_device(f)) return rc;
    if (n_modes > f->n_dense)
        return set_error(PBSO_ERR_RANGE, "mode id %d has no FFAT map (_ffat_maps->at(ii), modal_solver.h:296)", f->n_dense);
    // GetMapVal(pos, getCompressed) asserts _is_compressed when asked for compressed values
    // (ffat_solver.h:1183-1186); a compressed file leaves _Psi empty (ffat_map_serialize.h:238-252)
    const int ok_upto = use_compressed ? f->n_compressed : f->n_uncompressed;
    if (n_modes > ok_upto)
        return set_error(PBSO_ERR_UNSUPPORTED,
                         "map %d: is_compressed does not match use_compressed=%d (the reference reads an empty matrix here)",
                         ok_upto, use_compressed);
    return PBSO_OK;
}

extern "C" {

int pbso_ffat_load_dir(const char* dirname, pbso_ffat** out) {
    PBSO_REQUIRE(out && dirname, PBSO_ERR_INVALID, "null argument");
    pbso_ffat* f = new pbso_ffat();
    *out = f;                                    // like LoadAll: always hands back a (possibly empty) map
    std::vector<std::string> names;
    if (!list_dir_files(dirname, names, ".fatcube"))
        return set_error(PBSO_ERR_IO, "cannot open directory %s", dirname);
    for (const auto& name : names) {
        HostMap hm;
        if (int rc = load_one(name.c_str(), hm)) return rc;
        f->maps[hm.modeid] = std::move(hm);      // (*map)[map_.modeId] = map_  (:276)
    }
    return PBSO_OK;
}

int pbso_ffat_load_file(const char* filename, pbso_ffat** out) {
    PBSO_REQUIRE(out && filename, PBSO_ERR_INVALID, "null argument");
    *out = nullptr;
    HostMap hm;
    if (int rc = load_one(filename, hm)) return rc;
    pbso_ffat* f = new pbso_ffat();
    f->maps[hm.modeid] = std::move(hm);
    *out = f;
    return PBSO_OK;
}

int pbso_ffat_create(int n_maps, const int* mode_ids, const double* geom, const int* igeom,
                     const double* psi, int psi_len, const unsigned char* is_compressed, pbso_ffat** out) {
    PBSO_REQUIRE(out, PBSO_ERR_INVALID, "null output handle");
    *out = nullptr;
    PBSO_REQUIRE(n_maps >= 0 && (n_maps == 0 || (geom && igeom && psi && psi_len > 0)), PBSO_ERR_INVALID, "bad argument");
    pbso_ffat* f = new pbso_ffat();
    for (int i = 0; i < n_maps; ++i) {
        HostMap hm;
        std::memcpy(hm.geom, geom + (size_t)i * 32, sizeof(hm.geom));
        std::memcpy(hm.igeom, igeom + (size_t)i * 18, sizeof(hm.igeom));
        hm.psi.assign(1, std::vector<double>(psi + (size_t)i * psi_len, psi + (size_t)(i + 1) * psi_len));
        hm.is_compressed = is_compressed ? is_compressed[i] != 0 : false;
        hm.modeid = mode_ids ? mode_ids[i] : i;
        FatcubeMap fm; from_host_map(hm, fm);
        HostMap checked; std::string err;
        if (int rc = to_host_map(fm, checked, err)) { delete f; return set_error(rc, "map %d: %s", i, err.c_str()); }
        f->maps[hm.modeid] = std::move(hm);
    }
    *out = f;
    return PBSO_OK;
}

int pbso_ffat_destroy(pbso_ffat* f) {
    if (!f) return PBSO_OK;
    if (f->stream) {
        DeviceGuard g(f->device);
        cudaStreamSynchronize(f->stream);
        free_device(f); cudaFree(f->d_pos); cudaFree(f->d_out); cudaFree(f->d_loc); cudaFree(f->d_tile_rec); cudaFree(f->d_tile_cnt);
        cudaStreamDestroy(f->stream);
    }
    delete f;
    return PBSO_OK;
}

int pbso_ffat_num_maps(const pbso_ffat* f, int* n) {
    PBSO_REQUIRE(f && n, PBSO_ERR_INVALID, "null argument");
    *n = (int)f->maps.size();
    return PBSO_OK;
}

int pbso_ffat_mode_ids(const pbso_ffat* f, int* ids) {
    PBSO_REQUIRE(f && ids, PBSO_ERR_INVALID, "null argument");
    int i = 0;
    for (auto& kv : f->maps) ids[i++] = kv.first;
    return PBSO_OK;
}

int pbso_ffat_get_map(const pbso_ffat* f, int mode_id, double* geom32, int* igeom18, int* psi_len,
                      int* psi_cols, int* is_compressed, double* psi) {
    PBSO_REQUIRE(f, PBSO_ERR_INVALID, "null handle");
    auto it = f->maps.find(mode_id);
    if (it == f->maps.end()) return set_error(PBSO_ERR_RANGE, "no map with mode id %d", mode_id);
    const HostMap& hm = it->second;
    if (geom32) std::memcpy(geom32, hm.geom, sizeof(hm.geom));
    if (igeom18) std::memcpy(igeom18, hm.igeom, sizeof(hm.igeom));
    if (psi_len) *psi_len = (int)hm.psi[0].size();
    if (psi_cols) *psi_cols = (int)hm.psi.size();
    if (is_compressed) *is_compressed = hm.is_compressed;
    if (psi) for (size_t c = 0; c < hm.psi.size(); ++c) std::memcpy(psi + c * hm.psi[0].size(), hm.psi[c].data(), hm.psi[0].size() * sizeof(double));
    return PBSO_OK;
}

int pbso_ffat_save_file(const pbso_ffat* f, int mode_id, const char* filename) {
    PBSO_REQUIRE(f && filename, PBSO_ERR_INVALID, "null argument");
    auto it = f->maps.find(mode_id);
    if (it == f->maps.end()) return set_error(PBSO_ERR_RANGE, "no map with mode id %d", mode_id);
    FatcubeMap fm; from_host_map(it->second, fm);
    std::string bytes; fatcube_encode(fm, bytes);
    std::ofstream stream(filename, std::ios::binary);
    if (!stream) return set_error(PBSO_ERR_IO, "cannot open %s for writing", filename);
    stream.write(bytes.data(), (std::streamsize)bytes.size());
    return stream.good() ? PBSO_OK : set_error(PBSO_ERR_IO, "short write to %s", filename);
}

int pbso_ffat_save_legacy_file(const pbso_ffat* f, int mode_id, const char* filename) {
    PBSO_REQUIRE(f && filename, PBSO_ERR_INVALID, "null argument");
    auto it = f->maps.find(mode_id);
    if (it == f->maps.end()) return set_error(PBSO_ERR_RANGE, "no map with mode id %d", mode_id);
    FatcubeMap fm; from_host_map(it->second, fm);
    std::string bytes; legacy_fatcube_encode(fm, bytes);
    std::ofstream stream(filename, std::ios::binary);
    if (!stream) return set_error(PBSO_ERR_IO, "cannot open %s for writing", filename);
    stream.write(bytes.data(), (std::streamsize)bytes.size());
    return stream.good() ? PBSO_OK : set_error(PBSO_ERR_IO, "short write to %s", filename);
}

int pbso_ffat_eval(const pbso_ffat* fc, int n_modes, const double* pos, int L, int use_compressed, double* out) {
    pbso_ffat* f = const_cast<pbso_ffat*>(fc);
    if (int rc = check_eval_args(f, n_modes, L, use_compressed)) return rc;
    PBSO_REQUIRE(pos && out, PBSO_ERR_INVALID, "null argument");
    if (n_modes == 0 || L == 0) return PBSO_OK;
    DeviceGuard g(f->device);
    const size_t np = (size_t)3 * L, no = (size_t)L * n_modes;
    if (np > f->pos_cap) { cudaFree(f->d_pos); PBSO_CUDA(cudaMalloc(&f->d_pos, np * sizeof(double))); f->pos_cap = np; }
    if (no > f->out_cap) { cudaFree(f->d_out); PBSO_CUDA(cudaMalloc(&f->d_out, no * sizeof(double))); f->out_cap = no; }
    PBSO_CUDA(cudaMemcpyAsync(f->d_pos, pos, np * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    if (int rc = launch_eval(f, n_modes, f->d_pos, L, f->d_out, f->stream, use_compressed && f->q8_ready)) return rc;
    PBSO_CUDA(cudaMemcpyAsync(out, f->d_out, no * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    PBSO_CUDA(cudaStreamSynchronize(f->stream));
    return PBSO_OK;
}

int pbso_ffat_eval_device_view(const pbso_ffat* fc, int n_modes, const double* d_pos, int L, int use_compressed, double* d_out, void* cuda_stream) {
    pbso_ffat* f = const_cast<pbso_ffat*>(fc);
    if (int rc = check_eval_args(f, n_modes, L, use_compressed)) return rc;
    PBSO_REQUIRE(d_pos && d_out, PBSO_ERR_INVALID, "null argument");
    if (n_modes == 0 || L == 0) return PBSO_OK;
    DeviceGuard g(f->device);
    return launch_eval(f, n_modes, d_pos, L, d_out, cuda_stream ? (cudaStream_t)cuda_stream : f->stream, use_compressed && f->q8_ready);
}

int pbso_ffat_eval_device(const pbso_ffat* fc, int n_modes, const double* d_pos, int L, double* d_out, void* cuda_stream) {
    return pbso_ffat_eval_device_view(fc, n_modes, d_pos, L, 0, d_out, cuda_stream);
}

// ---- FFAT_Map<T,3>::Compress (ffat_solver.h:1125-1178), split where the reference goes through a JPEG file ----------
// cv::Mat::convertTo(CV_8U) is saturate_cast<uchar>(double): cvRound (round half to even; NaN and anything outside int
// come back as INT_MIN) and then a clamp to [0, 255].
static inline uint8_t cv_saturate_u8(double v) {
    if (!(std::fabs(v) < 2147483648.0)) return 0;              // NaN / out of int range -> INT_MIN -> 0
    const double r = std::nearbyint(v);                          // default rounding mode: half to even, like cvtsd2si
    return r <= 0.0 ? 0 : r >= 255.0 ? 255 : (uint8_t)r;
}

int pbso_ffat_quantise(const pbso_ffat* f, int mode_id, unsigned char* q8, double* max_amp6, double* max_amp_global) {
    PBSO_REQUIRE(f, PBSO_ERR_INVALID, "null handle");
    auto it = f->maps.find(mode_id);
    if (it == f->maps.end()) return set_error(PBSO_ERR_RANGE, "no map with mode id %d", mode_id);
    const HostMap& hm = it->second;
    if (!hm.has_plain()) return set_error(PBSO_ERR_UNSUPPORTED, "map %d was loaded compressed: it has no _Psi to compress", mode_id);
    const std::vector<double>& P = hm.psi[0];
    size_t off = 0; double gmax = -1.0;                           // :1132 maxAmp_global = -1
    for (int fc = 0; fc < 6; ++fc) {
        const size_t n = (size_t)hm.igeom[2 * fc] * hm.igeom[2 * fc + 1];       // ConvertToImages (:1106-1122): running offset
        if (off + n > P.size()) return set_error(PBSO_ERR_FORMAT, "map %d: faces address texels outside psi", mode_id);
        double mx = P[off];
        for (size_t i = 1; i < n; ++i) mx = std::max(mx, P[off + i]);          // A_amp.maxCoeff() (:1139)
        gmax = std::max(gmax, mx);                                             // :1134-1136
        const double up = 255 / mx;                                            // :1144 A_amp *= 255/maxAmp
        if (q8) for (size_t i = 0; i < n; ++i) q8[off + i] = cv_saturate_u8(P[off + i] * up);   // :1145-1147
        if (max_amp6) max_amp6[fc] = mx;
        off += n;
    }
    if (q8) for (size_t i = off; i < P.size(); ++i) q8[i] = 0;
    if (max_amp_global) *max_amp_global = gmax;
    return PBSO_OK;
}

int pbso_ffat_set_compressed_u8(pbso_ffat* f, int mode_id, const unsigned char* q8, int len, const double* max_amp6) {
    PBSO_REQUIRE(f && q8 && max_amp6, PBSO_ERR_INVALID, "null argument");
    auto it = f->maps.find(mode_id);
    if (it == f->maps.end()) return set_error(PBSO_ERR_RANGE, "no map with mode id %d", mode_id);
    HostMap& hm = it->second;
    if (!hm.has_plain()) return set_error(PBSO_ERR_UNSUPPORTED, "map %d was loaded compressed", mode_id);
    PBSO_REQUIRE(len == (int)hm.psi[0].size(), PBSO_ERR_INVALID, "q8 must hold one byte per texel of psi");
    hm.q8.assign(q8, q8 + len);
    for (int fc = 0; fc < 6; ++fc) { hm.q8_amp[fc] = max_amp6[fc]; hm.q8_scale[fc] = max_amp6[fc] / 255.; }   // :1162 A_amp *= maxAmp/255.
    hm.is_compressed = true;                                                   // :1173
    f->dirty = true;
    return PBSO_OK;
}

int pbso_ffat_compress(pbso_ffat* f, int mode_id, double* max_amp_global) {
    PBSO_REQUIRE(f, PBSO_ERR_INVALID, "null handle");
    std::vector<int> ids;
    if (mode_id >= 0) ids.push_back(mode_id); else for (auto& kv : f->maps) ids.push_back(kv.first);
    std::vector<unsigned char> q;
    for (size_t i = 0; i < ids.size(); ++i) {
        auto it = f->maps.find(ids[i]);
        if (it == f->maps.end()) return set_error(PBSO_ERR_RANGE, "no map with mode id %d", ids[i]);
        q.resize(it->second.psi[0].size());
        double amp[6], g = 0.0;
        if (int rc = pbso_ffat_quantise(f, ids[i], q.data(), amp, &g)) return rc;
        if (int rc = pbso_ffat_set_compressed_u8(f, ids[i], q.data(), (int)q.size(), amp)) return rc;
        if (max_amp_global) max_amp_global[i] = g;
    }
    return PBSO_OK;
}

int pbso_ffat_get_compressed(const pbso_ffat* f, int mode_id, unsigned char* q8, double* max_amp6, double* compressed_psi) {
    PBSO_REQUIRE(f, PBSO_ERR_INVALID, "null handle");
    auto it = f->maps.find(mode_id);
    if (it == f->maps.end()) return set_error(PBSO_ERR_RANGE, "no map with mode id %d", mode_id);
    const HostMap& hm = it->second;
    if (!hm.is_compressed) return set_error(PBSO_ERR_UNSUPPORTED, "map %d is not compressed", mode_id);
    if ((q8 || max_amp6) && hm.q8.empty()) return set_error(PBSO_ERR_UNSUPPORTED, "map %d was loaded compressed: only its doubles are known", mode_id);
    if (q8) std::memcpy(q8, hm.q8.data(), hm.q8.size());
    if (max_amp6) std::memcpy(max_amp6, hm.q8_amp, sizeof(hm.q8_amp));
    if (compressed_psi) { const std::vector<double> c = hm.compressed_column(); std::memcpy(compressed_psi, c.data(), c.size() * sizeof(double)); }
    return PBSO_OK;
}

}  // extern "C"
