// =============================================================================
// batch_tc.cu -- offline batch synthesis on the 5th-generation tensor cores (PBSO_PREC_TC3X).
//
// Same job as k_batch_pow (batch.cu): every object is a ModalSolver::step loop (modal_solver.h:181-276) driven
// by PointForce messages, mixed down.  The pole-power form of the recurrence (modal_integrator.h:109-110)
//     q_m[(i L) + j] = Re v_m(i) * Im(w_m^j) + Im v_m(i) * Re(w_m^j),      v_m(i) = state at the start of tile i
// summed over modes with the transfer T_m (modal_solver.h:267-269) and over objects is ONE matrix product
//     Y[i][j] = sum_k A[i][k] B[j][k],    k = (object, mode, component),
//     A[i][(o,m,0..1)] = T_{o,m} (Re, Im) v_{o,m}(i)   -- tile-start states  ("advance each mode with z^k")
//     B[j][(o,m,0..1)] = (Im, Re) w_{o,m}^j            -- pole powers
// with i = tile index in time (L = 128 samples per tile), j = offset inside the tile.  For cfg5 that is a
// [3446 x 128] output contracted over K = 4096*512*2 = 4.2 M, run by tcgen05.mma kind::tf32 with accumulators in
// TMEM.
//
// Round-2 structure (profiles/r2_k_batch_tc.md; round 1 was latency-bound on its own five-role pipeline):
//  * CTA PAIRS (cluster of 2, tcgen05.mma cta_group::2, M = 256): the pair renders two M-tiles (128 time tiles each) of
//    the SAME object.  Operand B depends on the object only, so each CTA generates HALF of it (64 of the 128 rows of
//    every K chunk: pole powers from table factors by FP32 complex products, split into TF32 hi / lo, stored as UMMA
//    tiles -- K-major, 128-byte swizzle) and the tensor cores of both SMs read the two halves: B generation, its
//    shared-memory stores and the UMMA operand fetches per output halve.  The leader CTA issues the MMAs for both;
//    generators of both CTAs arrive on the leader's barriers through plain (CTA-scope) remote arrives -- the
//    .release.cluster / .acquire.cluster forms compile to GPU-scope fences and L1 invalidations and cost 30 %.
//    (Streaming precomputed B tiles from L2 with multicast bulk copies was built and measured first: correct, but
//    no faster than generating -- the per-SM ingest of bulk copies (~35 B/clk) and L2 bandwidth bound it.)
//    Work items are (window of 12 objects for cfg5, M-tile pair), dealt round-robin to the clusters, so the clusters rendering
//    the M-tile pairs of one object window read the same 5.4 KB table blocks at about the same time: HBM sees every
//    block about once per render (1.2 GB for cfg5 in total), L2 serves the rest.
//  * Operand A (tile-start states) is generated on the SM straight into TMEM (tcgen05.st, MMA in TS mode):
//    A[row] = X[blk] * R[j] (row = 16 blk + j), X[blk] = v W^(16 blk) from a seed warp, R[j] = W^j, W = w^L, all
//    FP32 complex products (fma.rn.f32x2) of table factors computed in FP64 and rounded once (k_tc_tabs); v comes
//    from the FP64 carrier pass (k_tc_carrier, per render; impulses injected in FP64, transfer folded in).
//    Everything a chunk needs -- 5.4 KB of table, 128 B of states -- arrives through a ring of bulk copies
//    (cp.async.bulk + mbarrier) requested eight chunks ahead: no role ever waits on a global load.
//  * "3xTF32": every FP32 element x is split x = hi + lo (the tensor core ignores the low 13 mantissa bits itself,
//    so the raw x is the hi operand; lo = x - trunc(x) rounded to nearest TF32); hi*lo + lo*hi + hi*hi.  The tensor
//    core truncates when it adds into its FP32 accumulator, so accumulation is two-level: chains of two K chunks
//    (16 small MMAs first -- their truncation is relative to a 2^-11 smaller sum -- then the 8 hi*hi MMAs on top,
//    ONE accumulator) are promoted into FP32 registers with round-to-nearest packed adds, and the registers are
//    added to the FP64 mix (RED.64) every few units.  The remaining mean truncation of a chain (a coherent gain
//    error of ~2e-7) is measured once per device by tc_calibrate() -- the same kernel on a synthetic batch against
//    the FP64 direct-form kernel -- and divided out at the flush.
//
// CTA = 24 warps, ordered by how much the pipeline waits for them (the scheduler prefers the higher warp id of a
// sub-partition): warps 0-7 generate A and warps 8-15 generate B (two groups of four each, alternate K chunks; they run
// ahead of the MMAs), warps 16-19 drain TMEM (epilogue), warps 20-21 issue the MMAs (even / odd K chunks), warp 22 computes
// the seeds, warp 23 lane 0 is the table loader.  setmaxnreg: 64 / 64 / 64 / 64 / 168 / 56 registers per thread of the six warp groups.
// TMEM (512 columns): accumulators 0..127 | 128..255, four A stages of 64 columns (hi 32 | lo 32) at 256.
// =============================================================================
#include "common.cuh"
#include "umma.cuh"
#include "batch_tc.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace pbso;
using namespace pbso::umma;

namespace {

constexpr int TCB_L = 128;                 // samples per tile = N of the MMA
constexpr int TCB_ROWS = 128;              // tiles per M-tile = M of the MMA
constexpr int TCB_KMODES = 16;             // modes per K chunk (K = 32 fp32 = one 128-byte swizzle row)
constexpr int TCB_TILE_BYTES = 128 * 32 * 4;
constexpr int TCB_BT_BYTES = 2 * TCB_TILE_BYTES;        // one chunk of operand B in shared memory: hi tile | lo tile
constexpr int TCB_THREADS = 768;              // 24 warps, see the role table at k_batch_tc
constexpr int TCB_TMEM_COLS = 512;
constexpr int TCB_NB = 4;                  // B stages
constexpr int TCB_NA = 4;                  // A stages in TMEM
constexpr int TCB_NT = 8;                  // table slots (requested well ahead: the copy latency is what the ring hides)
constexpr int TCB_NS = 4;                  // seed slots
constexpr int TCB_RSTRIDE = 144;                        // bytes between the R rows of a table block (128 + 16: 16 rows spread over all banks)
constexpr int TCB_TABG_BYTES = 1024 + 17 * TCB_RSTRIDE + 2048;   // per chunk in HBM: P[8 blk][16 m] | R[16 j + a zero row][16 m (+pad)] | tabB[16 entries][16 m], float2
constexpr int TCB_TAB_R = 1024, TCB_TAB_B = 1024 + 17 * TCB_RSTRIDE;
constexpr int TCB_TAB_BYTES = TCB_TABG_BYTES + 128;     // + the chunk's 16 tile-start states
constexpr int TCB_SEED_BYTES = 1024;                    // X[8 blk][16 m] float2: v W^(16 blk)
constexpr int TCB_SMEM = TCB_NB * TCB_BT_BYTES + TCB_NT * TCB_TAB_BYTES + TCB_NS * TCB_SEED_BYTES + 512 + 1024;
constexpr int TCB_WINDOW = 8;              // objects per work item, at least (build_units: TCB_FLUSH_CHUNKS / chunks per unit, 12 for 512 modes)
constexpr int TCB_FLUSH_UNITS = 8;         // units between FP64 flushes of the register accumulators, at least (= the window)
constexpr int TCB_FLUSH_CHUNKS = 384;      // ... and never more K chunks than this in one FP32 running sum (12 units of 512 modes)
static_assert(TCB_SMEM <= 232448, "shared memory budget");
static_assert(TCB_NA == TCB_NB, "A and B stages share their full / empty barriers");

// One unit of work: object `obj` over M-tile `it` (128 time tiles) from the state block `src` -- tile-start states of a
// carry unit, or the injected state of an impulse unit landing on row `re` of the M-tile.  CTA pairs (cta_group::2)
// render TWO M-tiles of the object per unit, 2 it + rank, each CTA from its own src[rank] / re[rank] (or the zero block).
struct Unit {
    int obj, it;
    unsigned src[2];       // block index into the state buffer (blocks of mp float2)
    short re[2];           // impulse row inside the M-tile, -1 for a carry (or idle) unit
    short flush, pad;      // add the register accumulators to the mix after this unit
};

struct Cplx { double x, y; };
__device__ __forceinline__ Cplx cmul(const Cplx a, const Cplx b) {
    Cplx r;
    r.x = fma(-a.y, b.y, a.x * b.x);
    r.y = fma(a.x, b.y, a.y * b.x);
    return r;
}
__device__ __forceinline__ Cplx pole_pow64(double le, double th, double k) {
    double s, c;
    sincos(k * th, &s, &c);
    const double e = exp(k * le);
    return Cplx{e * c, e * s};
}

// byte offset of (row r, K-columns 2m..2m+1) inside a K-major SWIZZLE_128B tile with 128-byte rows
__host__ __device__ __forceinline__ uint32_t sw128_pair(int r, int m) {
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((m >> 1) ^ (r & 7)) << 4) + (uint32_t)(m & 1) * 8u;
}
__device__ __forceinline__ float tf32_rn(float x) {                   // round to nearest TF32 (10 explicit mantissa bits)
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// ---- operand tables, once per handle (they depend on the poles only) -------------------------------------------
// tab[(o * cpu + ch)]: 5.4 KB = P[8 blk][16 m] | R[16 j + zero row][16 m + pad] | tabB[16 entries][16 m], float2.
//   operand A:  P[blk] = W^(16 blk), R[j] = W^j with W = w^L (tile to tile)
//   operand B:  rows hold (Im, Re) w^j = i conj(w^j); i conj(z1 z2) = (i conj z1) conj(z2), so entries 0..7 are the
//               16-row block starts i conj(w^(16 a)) and entries 8..10 / 11..13 the conjugated steps conj(w^(4 t)),
//               conj(w^t), t = 1..3: the generator runs plain complex products.
// Modes past n_modes are zero.
__global__ void k_tc_tabs(int n_obj, int n_modes, int cpu, int obj0, const double* __restrict__ lneps, const double* __restrict__ theta,
                          uint8_t* __restrict__ tab) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;       // (object, padded mode)
    const size_t n = (size_t)n_obj * cpu * TCB_KMODES;
    if (i >= n) return;
    const int o = (int)(i / ((size_t)cpu * TCB_KMODES)), mp = (int)(i % ((size_t)cpu * TCB_KMODES));
    const int m_l = mp % TCB_KMODES;
    uint8_t* blk = tab + ((size_t)o * cpu + mp / TCB_KMODES) * TCB_TABG_BYTES;
    float2* P = reinterpret_cast<float2*>(blk) + m_l;                     // entry b at P[16 b]
    uint8_t* Rb = blk + TCB_TAB_R + m_l * 8;                              // entry j at Rb + j * TCB_RSTRIDE
    float2* TB = reinterpret_cast<float2*>(blk + TCB_TAB_B) + m_l;        // entry e at TB[16 e]
    *reinterpret_cast<float2*>(Rb + 16 * TCB_RSTRIDE) = make_float2(0.f, 0.f);   // row 16 = 0: rows before an impulse
    if (m_l == 0)                                                          // the pad column of every R row
        for (int j = 0; j < 17; ++j) { float2* pad = reinterpret_cast<float2*>(blk + TCB_TAB_R + j * TCB_RSTRIDE + 128); pad[0] = pad[1] = make_float2(0.f, 0.f); }
    if (mp >= n_modes) {
#pragma unroll
        for (int b = 0; b < 8; ++b) P[16 * b] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 16; ++j) { *reinterpret_cast<float2*>(Rb + j * TCB_RSTRIDE) = make_float2(0.f, 0.f); TB[16 * j] = make_float2(0.f, 0.f); }
        return;
    }
    const size_t src = (size_t)(obj0 + o) * n_modes + mp;
    const double le = lneps[src], th = theta[src];
#pragma unroll
    for (int b = 0; b < 8; ++b) { const Cplx z = pole_pow64(le, th, 16.0 * b * TCB_L); P[16 * b] = make_float2((float)z.x, (float)z.y); }
#pragma unroll
    for (int j = 0; j < 16; ++j) { const Cplx z = pole_pow64(le, th, (double)j * TCB_L); *reinterpret_cast<float2*>(Rb + j * TCB_RSTRIDE) = make_float2((float)z.x, (float)z.y); }
#pragma unroll
    for (int a = 0; a < 8; ++a) { const Cplx z = pole_pow64(le, th, 16.0 * a); TB[16 * a] = make_float2((float)z.y, (float)z.x); }
#pragma unroll
    for (int t = 1; t < 4; ++t) {
        { const Cplx z = pole_pow64(le, th, 4.0 * t); TB[16 * (7 + t)] = make_float2((float)z.x, (float)-z.y); }
        { const Cplx z = pole_pow64(le, th, (double)t); TB[16 * (10 + t)] = make_float2((float)z.x, (float)-z.y); }
    }
    TB[16 * 14] = TB[16 * 15] = make_float2(0.f, 0.f);
}

// ---- FP64 carrier: T * state of every (object, mode) at the start of every M-tile -----------------------------
// V[it] excludes impulses landing at rows >= it*128 (those are impulse units of M-tile it).  Layout: blocks of mp
// float2 (mp = modes padded to whole chunks; pad entries stay zero), block it * n_obj + o.
__global__ void k_tc_carrier(int n_obj, int n_modes, int n_it, int buf_size, int mp, int obj0,
                             const double* __restrict__ lneps, const double* __restrict__ theta,
                             const double* __restrict__ c3a, const double* __restrict__ cota, const double* __restrict__ trans,
                             const int* __restrict__ ev_off, const int* __restrict__ ev_buf,
                             const double* __restrict__ ev_space, const double* __restrict__ v0r, const double* __restrict__ v0i,
                             float2* __restrict__ V) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_obj * n_modes) return;
    const int o = (int)(idx / n_modes), m = (int)(idx % n_modes);
    const size_t g = (size_t)(obj0 + o) * n_modes + m;
    const double le = lneps[g], th = theta[g], T = trans[g];
    const Cplx Wm = pole_pow64(le, th, (double)(TCB_L * TCB_ROWS));
    const double inji = c3a[g], injr = inji * cota[g];
    Cplx v{0.0, 0.0};
    if (v0r) v = Cplx{v0r[g], v0i[g]};                       // stateful range render: the state the range starts from
    int e = ev_off[obj0 + o];
    const int e_end = ev_off[obj0 + o + 1];
    for (int it = 0; it < n_it; ++it) {
        V[((size_t)it * n_obj + o) * mp + m] = make_float2((float)(T * v.x), (float)(T * v.y));
        v = cmul(v, Wm);
        const long long row_end = (long long)(it + 1) * TCB_ROWS;
        while (e < e_end) {
            const long long t_e = (long long)ev_buf[e] * buf_size;           // the impulse's sample; it joins the tensor-core
            const long long row = (t_e + TCB_L - 1) / TCB_L;                  // path at the next tile boundary (row)
            if (row >= row_end) break;
            const Cplx z = pole_pow64(le, th, (double)(row_end * TCB_L - t_e));   // from the impulse to the next M-tile start
            const double sp = ev_space[(size_t)e * n_modes + m];
            v.x += sp * (injr * z.x - inji * z.y);
            v.y += sp * (injr * z.y + inji * z.x);
            ++e;
        }
    }
}
// state injected by impulse e (forces.h:87, sample 0 of its buffer): T * space * c3 (cot theta + i); block vimp0 + e
// An impulse that does not land on a tile boundary (buf_size not a multiple of 128) is carried to the next boundary by
// the exact pole power w^(boundary - t_e); the samples in between are k_batch_event_heads' (batch.cu).
__global__ void k_tc_impulse(int n_modes, int mp, int e0, int n_ev, int buf_size, const int* __restrict__ ev_obj, const int* __restrict__ ev_buf,
                             const double* __restrict__ lneps, const double* __restrict__ theta,
                             const double* __restrict__ c3a, const double* __restrict__ cota, const double* __restrict__ trans,
                             const double* __restrict__ ev_space, float2* __restrict__ Vimp) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_ev * n_modes) return;
    const int e = (int)(idx / n_modes), m = (int)(idx % n_modes);
    const size_t g = (size_t)ev_obj[e0 + e] * n_modes + m;
    const double sp = ev_space[(size_t)(e0 + e) * n_modes + m] * trans[g], inji = c3a[g];
    Cplx v{sp * inji * cota[g], sp * inji};
    const long long t_e = (long long)ev_buf[e0 + e] * buf_size;
    const int ahead = (int)((TCB_L - t_e % TCB_L) % TCB_L);
    if (ahead) v = cmul(v, pole_pow64(lneps[g], theta[g], (double)ahead));
    Vimp[(size_t)e * mp + m] = make_float2((float)v.x, (float)v.y);
}

// ---- device helpers of the main kernel ----------------------------------------------------------------------
typedef unsigned long long c32;                                          // packed (re, im) or any f32x2
__device__ __forceinline__ c32 pk(float a, float b) { c32 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(c32 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ c32 mul2(c32 a, c32 b) { c32 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c32 fma2(c32 a, c32 b, c32 c) { c32 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ c32 add2(c32 a, c32 b) { c32 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c32 sub2(c32 a, c32 b) { c32 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// complex product p * q in plain registers (seed warps: a few per chunk)
__device__ __forceinline__ float2 cmulf(float2 p, float2 q) {
    return make_float2(fmaf(-p.y, q.y, p.x * q.x), fmaf(p.x, q.y, p.y * q.x));
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, c32& a, c32& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// acc[0..15] (packed column pairs) += 32 consecutive TMEM columns of this thread's lane, round-to-nearest FP32 adds.
// One asm block: the 32 loaded registers are consumed pairwise by add.rn.f32x2 without leaving the register pairs the
// load wrote (no packing moves).
__device__ __forceinline__ void tmem_ld_add_32x32(uint32_t taddr, c32* acc) {
    asm volatile(
        "{\n\t.reg .b32 t<32>;\n\t.reg .b64 p<16>;\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, "
        "t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n\t"
        "mov.b64 p0, {t0, t1};\n\tmov.b64 p1, {t2, t3};\n\tmov.b64 p2, {t4, t5};\n\tmov.b64 p3, {t6, t7};\n\t"
        "mov.b64 p4, {t8, t9};\n\tmov.b64 p5, {t10, t11};\n\tmov.b64 p6, {t12, t13};\n\tmov.b64 p7, {t14, t15};\n\t"
        "mov.b64 p8, {t16, t17};\n\tmov.b64 p9, {t18, t19};\n\tmov.b64 p10, {t20, t21};\n\tmov.b64 p11, {t22, t23};\n\t"
        "mov.b64 p12, {t24, t25};\n\tmov.b64 p13, {t26, t27};\n\tmov.b64 p14, {t28, t29};\n\tmov.b64 p15, {t30, t31};\n\t"
        "add.rn.f32x2 %0, %0, p0;\n\tadd.rn.f32x2 %1, %1, p1;\n\tadd.rn.f32x2 %2, %2, p2;\n\tadd.rn.f32x2 %3, %3, p3;\n\t"
        "add.rn.f32x2 %4, %4, p4;\n\tadd.rn.f32x2 %5, %5, p5;\n\tadd.rn.f32x2 %6, %6, p6;\n\tadd.rn.f32x2 %7, %7, p7;\n\t"
        "add.rn.f32x2 %8, %8, p8;\n\tadd.rn.f32x2 %9, %9, p9;\n\tadd.rn.f32x2 %10, %10, p10;\n\tadd.rn.f32x2 %11, %11, p11;\n\t"
        "add.rn.f32x2 %12, %12, p12;\n\tadd.rn.f32x2 %13, %13, p13;\n\tadd.rn.f32x2 %14, %14, p14;\n\tadd.rn.f32x2 %15, %15, p15;\n\t}"
        : "+l"(acc[0]), "+l"(acc[1]), "+l"(acc[2]), "+l"(acc[3]), "+l"(acc[4]), "+l"(acc[5]), "+l"(acc[6]), "+l"(acc[7]),
          "+l"(acc[8]), "+l"(acc[9]), "+l"(acc[10]), "+l"(acc[11]), "+l"(acc[12]), "+l"(acc[13]), "+l"(acc[14]), "+l"(acc[15])
        : "r"(taddr) : "memory");
}

// =============================================================================================================
// PAIR = 2: clusters of two CTAs, tcgen05.mma cta_group::2 (M = 256): the pair renders two M-tiles of one object and each
// CTA generates only HALF of operand B (64 of its 128 rows) -- the tensor cores fetch the other half from the peer.
template <int PAIR>
__global__ void __launch_bounds__(TCB_THREADS, 1)
k_batch_tc(int n_modes, int n_tiles, long long n_samples, const int* __restrict__ cta_first, const Unit* __restrict__ units,
           const uint8_t* __restrict__ tab, const float2* __restrict__ V, int obj0,
           double* __restrict__ mix, float* __restrict__ stems, double inv_gain, int ablate, unsigned long long* __restrict__ prof, int flush_chunks) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* tabs = smem + TCB_NB * TCB_BT_BYTES;
    uint8_t* seeds = tabs + TCB_NT * TCB_TAB_BYTES;
    uint64_t* bars = (uint64_t*)(seeds + TCB_NS * TCB_SEED_BYTES);
    uint64_t* b_full = bars;                       // [NB] = [NA] operand stage written: A in TMEM + B in shared memory (4 + 4 generator warps)
    uint64_t* b_empty = b_full + TCB_NB;           // [NB] MMAs reading the stage retired (A and B generators both wait on it)
    uint64_t* t_full = b_empty + TCB_NB;           // [NT] table block + states landed (tx bytes)
    uint64_t* t_empty = t_full + TCB_NT;           // [NT] consumed (seed warp + 4 A-generator + 4 B-generator warps)
    uint64_t* s_full = t_empty + TCB_NT;           // [NS] seeds written (seed warp)
    uint64_t* s_empty = s_full + TCB_NS;           // [NS] consumed (4 A-generator warps)
    uint64_t* acc_full = s_empty + TCB_NS;         // [2]  chain's MMAs finished
    uint64_t* acc_empty = acc_full + 2;            // [2]  drained (4 epilogue warps)
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR == 2 ? cluster_ctarank() : 0;
    const int u0 = cta_first[blockIdx.x / PAIR], u1 = cta_first[blockIdx.x / PAIR + 1];
    // optional in-kernel timing (PBSO_TC_PROF=1): cycles every role of CTA 0 spends in each of its barrier waits
    unsigned long long pw[4] = {0ull, 0ull, 0ull, 0ull};
    const bool profiling = prof != nullptr && blockIdx.x == 0;
    // barriers the MMA issuers (rank 0) wait on collect arrivals from both CTAs of a pair
    const uint32_t lead_bfull = PAIR == 2 ? mapa_u32(smem_u32(b_full), 0) : smem_u32(b_full);
    const uint32_t lead_accempty = PAIR == 2 ? mapa_u32(smem_u32(acc_empty), 0) : smem_u32(acc_empty);
#define PBSO_ARRIVE_LEAD(base, idx) do { if (PAIR == 2) mbar_arrive_cluster((base) + 8u * (uint32_t)(idx)); else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((base) + 8u * (uint32_t)(idx)) : "memory"); } while (0)
    const long long t_role0 = clock64();
#define PBSO_TW(slot, stmt) do { if (profiling) { const long long t_ = clock64(); stmt; pw[slot] += (unsigned long long)(clock64() - t_); } else { stmt; } } while (0)
    const int cpu = (n_modes + TCB_KMODES - 1) / TCB_KMODES;          // K chunks per unit
    const int mp = cpu * TCB_KMODES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TCB_NB; ++s) { mbar_init(&b_full[s], 8 * PAIR); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < TCB_NT; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 9); }
        for (int s = 0; s < TCB_NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4 * PAIR); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 20) {
        if (PAIR == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCB_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCB_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tcgen05_fence_before();
    if (PAIR == 2) cluster_sync_all(); else __syncthreads();          // the peer's barriers exist before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ---------------- A generators: thread = row (TMEM lane); A[row][2m..2m+1] = X[blk][m] * R[j][m] ----------------
        // Two groups of four warps (0-3, 4-7) take alternate K chunks: a generator warp is a long dependent chain per
        // chunk (loads, packed complex products, splits, TMEM stores, three barrier round trips), and one group alone
        // cannot turn a chunk around in the 768 cycles its MMAs take.  Per mode two 16-byte loads -- X4 = (xr, xr, xi, xi)
        // from the seed slot (two addresses per warp: broadcast), R4 = (rr, ri, -ri, rr) straight from the table slot -- feed
        // ONE mul.f32x2 + fma.f32x2; an impulse unit's rows before the impulse read the zero row.  A chunk is generated in
        // two halves of 8 modes (32 registers each).
        const int grp = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane, blk = row >> 4, j = row & 15;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        uint32_t q = 0;
#pragma unroll 1
        for (int u = u0; u < u1; ++u) {
            const int re = units[u].re[rank];
            int jj = j;
            if (re >= 0 && blk == (re >> 4)) { jj = j - (re & 15); if (jj < 0) jj = 16; }   // impulse unit: rows of block ae are shifted by be; row 16 = 0
#pragma unroll 1
            for (int ch = 0; ch < cpu; ++ch, ++q) {
                if ((int)(q & 1) != grp) continue;
                const uint32_t ss = q % TCB_NS, sa = q % TCB_NA, ts = q % TCB_NT;
                PBSO_TW(0, mbar_wait(&t_full[ts], (q / TCB_NT) & 1));
                PBSO_TW(1, mbar_wait(&s_full[ss], (q / TCB_NS) & 1));
                const uint32_t sX = smem_u32(seeds) + ss * TCB_SEED_BYTES + blk * 128;
                const uint32_t sR = smem_u32(tabs) + ts * TCB_TAB_BYTES + TCB_TAB_R + jj * TCB_RSTRIDE;
                const uint32_t a_col = tmem_base + lane_off + 256 + sa * 64;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int m2 = 0; m2 < 4; ++m2) {
                        float4 x, r;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(sX + (4 * hf + m2) * 16));
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sR + (4 * hf + m2) * 16));
                        float v[4];
                        v[0] = fmaf(-x.y, r.y, x.x * r.x); v[1] = fmaf(x.x, r.y, x.y * r.x);
                        v[2] = fmaf(-x.w, r.w, x.z * r.z); v[3] = fmaf(x.z, r.w, x.w * r.z);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            // hi = the raw value (the tensor core reads its upper 19 bits); lo = what those bits miss (the tensor
                            // core truncates it to TF32: a mean shift of the gain that tc_calibrate() measures with everything else)
                            const uint32_t bits = __float_as_uint(v[c]);
                            hi[4 * m2 + c] = bits;
                            lo[4 * m2 + c] = __float_as_uint(v[c] - __uint_as_float(bits & 0xFFFFE000u));
                        }
                    }
                    if (hf == 0) {
                        PBSO_TW(2, mbar_wait_relaxed(&b_empty[sa], ((q / TCB_NA) & 1) ^ 1));
                        tcgen05_fence_after();
                    } else {
                        // the slots are released only once every loaded value has been USED: a release issued right behind the
                        // loads can overtake them in the MIO queue, and the seed warp refills its slot within tens of cycles
                        asm volatile("" ::"r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]),
                                     "r"(hi[8]), "r"(hi[9]), "r"(hi[10]), "r"(hi[11]), "r"(hi[12]), "r"(hi[13]), "r"(hi[14]), "r"(hi[15]) : "memory");
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(&s_empty[ss]); mbar_arrive(&t_empty[ts]); }
                    }
                    tmem_st_32x16(a_col + 16 * hf, hi);
                    tmem_st_32x16(a_col + 32 + 16 * hf, lo);
                }
                PBSO_TW(3, tmem_st_wait());
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) PBSO_ARRIVE_LEAD(lead_bfull, sa);
            }
        }
    } else if (warp < 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ---------------- B generators: thread = (mode of the chunk, 16-row block); rows 4t + c = x Rt[t] Rc[c] ----------------
        // two groups of four warps (8-11, 12-15) take alternate K chunks, like the A generators
        // straight into the UMMA K-major 128-byte-swizzle layout: row r = 16 blk + b, K columns (2 m, 2 m + 1) at byte
        // (r / 8) 1024 + (r % 8) 128 + (((m / 2) ^ (r % 8)) 16) + (m % 2) 8; the eight swizzled offsets of a thread never change
        // PAIR = 2: this CTA holds rows [64 rank, 64 rank + 64) of the tile: thread = (mode, 8-row half block), local row 8 hb + i
        const int m_l = lane & 15, bgrp = (warp - 8) >> 2, hb = ((warp - 8) & 3) * 2 + (lane >> 4);
        const int blk = PAIR == 2 ? 4 * (int)rank + (hb >> 1) : hb;               // 16-row block of the tile the thread starts in
        const int t0 = PAIR == 2 ? 2 * (hb & 1) : 0;                               // first 4-row step inside the block
        constexpr int NT4 = PAIR == 2 ? 2 : 4;                                     // 4-row steps per thread
        uint32_t xoff[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xoff[i] = PAIR == 2 ? sw128_pair(8 * hb + i, m_l) : sw128_pair(16 * hb + i, m_l);
        const uint32_t n_chunks = (uint32_t)(u1 - u0) * (uint32_t)cpu;
#pragma unroll 1
        for (uint32_t q = (uint32_t)bgrp; q < n_chunks; q += 2) {
            const uint32_t sb = q % TCB_NB, ts = q % TCB_NT;
            PBSO_TW(0, mbar_wait(&t_full[ts], (q / TCB_NT) & 1));
            const uint32_t tB = smem_u32(tabs) + ts * TCB_TAB_BYTES + TCB_TAB_B + m_l * 8;   // entry e at tB + 128 e
            const float2 ra = lds_f2(tB + 128 * blk);
            float2 rt[3], rc[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) { rt[t] = lds_f2(tB + 128 * (8 + t)); rc[t] = lds_f2(tB + 128 * (11 + t)); }
            // broadcast pairs of the step powers: p * q = (pr, pr) * q + (pi, pi) * (i q)
            c32 rtr[3], rti[3], rcr[3], rci[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) { rtr[t] = pk(rt[t].x, rt[t].x); rti[t] = pk(rt[t].y, rt[t].y); rcr[t] = pk(rc[t].x, rc[t].x); rci[t] = pk(rc[t].y, rc[t].y); }
            const c32 xa = pk(ra.x, ra.y), xar = pk(-ra.y, ra.x);
            // PAIR = 2: the thread's two 4-row steps are t = t0, t0 + 1 with t0 = 0 or 2: multipliers 1, W^4 or W^8, W^12
            const float2 m0 = t0 ? rt[1] : make_float2(1.f, 0.f), m1 = t0 ? rt[2] : rt[0];
            const c32 m0r = pk(m0.x, m0.x), m0i = pk(m0.y, m0.y), m1r = pk(m1.x, m1.x), m1i = pk(m1.y, m1.y);
            PBSO_TW(1, mbar_wait_relaxed(&b_empty[sb], ((q / TCB_NB) & 1) ^ 1));
            const uint32_t st = smem_u32(smem) + sb * TCB_BT_BYTES;
            if (!(ablate & 1)) {
#pragma unroll
                for (int tt = 0; tt < NT4; ++tt) {
                    // step t = t0 + tt of the block: x W^(4 t)
                    c32 pt;
                    if (PAIR == 2) pt = fma2(tt == 0 ? m0i : m1i, xar, mul2(tt == 0 ? m0r : m1r, xa));
                    else pt = tt == 0 ? xa : fma2(rti[tt - 1], xar, mul2(rtr[tt - 1], xa));
                    float pr, pi; upk(pt, pr, pi);
                    const c32 ptr_ = pk(-pi, pr);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const c32 v = c == 0 ? pt : fma2(rci[c - 1], ptr_, mul2(rcr[c - 1], pt));
                        float vr, vi; upk(v, vr, vi);
                        const uint32_t br = __float_as_uint(vr), bi = __float_as_uint(vi);
                        const c32 tr = pk(__uint_as_float(br & 0xFFFFE000u), __uint_as_float(bi & 0xFFFFE000u));
                        float lr, li; upk(sub2(v, tr), lr, li);
                        const int b = 4 * tt + c;                                  // row of the thread's 8- or 16-row run
                        const uint32_t addr = st + xoff[b & 7] + (uint32_t)(b >> 3) * 1024u;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(br), "r"(bi) : "memory");
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + TCB_TILE_BYTES), "f"(lr), "f"(li) : "memory");
                    }
                }
            }
            PBSO_TW(2, fence_proxy_async_smem());
            __syncwarp();
            // the table slot is released only here: its loads are certainly complete once their values have been used
            if (lane == 0) { mbar_arrive(&t_empty[ts]); PBSO_ARRIVE_LEAD(lead_bfull, sb); }
        }
    } else if (warp < 20) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        // ---------------- epilogue: promote finished chains into registers, flush to the FP64 mix ----------------
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        c32 acc[TCB_L / 2];                                                  // packed pairs of columns
#pragma unroll
        for (int j = 0; j < TCB_L / 2; ++j) acc[j] = 0ull;
        uint32_t g = 0;
        int chunks_held = 0;                                                 // K chunks summed in the registers since the last flush
        // registers -> FP64 mix (and / or the object's float stem), then cleared
        auto flush = [&](const Unit& un) {
            const long long tile = ((long long)un.it * PAIR + rank) * TCB_ROWS + row;
            // the last tile of a render whose length is not a multiple of 128 is cut at n_samples
            const int lim = tile < n_tiles ? (int)min((long long)TCB_L, n_samples - tile * TCB_L) : 0;
            if (lim > 0 && mix) {
                double* dst = mix + tile * TCB_L;
#pragma unroll
                for (int j = 0; j < TCB_L / 2; ++j) {
                    float a, b; upk(acc[j], a, b);
                    if (2 * j < lim) atomicAdd(dst + 2 * j, (double)a * inv_gain);
                    if (2 * j + 1 < lim) atomicAdd(dst + 2 * j + 1, (double)b * inv_gain);
                }
            }
            if (lim > 0 && stems) {                                  // per-object stems: the host flushes at every object change
                float* dst = stems + (size_t)un.obj * n_samples + tile * TCB_L;
                const float ig = (float)inv_gain;
#pragma unroll
                for (int j = 0; j < TCB_L / 2; ++j) {
                    float a, b; upk(acc[j], a, b);
                    if (2 * j < lim) atomicAdd(dst + 2 * j, a * ig);
                    if (2 * j + 1 < lim) atomicAdd(dst + 2 * j + 1, b * ig);
                }
            }
#pragma unroll
            for (int j = 0; j < TCB_L / 2; ++j) acc[j] = 0ull;
            chunks_held = 0;
        };
#pragma unroll 1
        for (int u = u0; u < u1; ++u) {
            const Unit un = units[u];
#pragma unroll 1
            for (int ch = 0; ch < cpu; ch += 2, ++g) {
                const int buf = g & 1;
                PBSO_TW(0, mbar_wait(&acc_full[buf], (g >> 1) & 1));
                tcgen05_fence_after();
                if (!(ablate & 4)) {
#pragma unroll
                    for (int qd = 0; qd < TCB_L / 32; ++qd) tmem_ld_add_32x32(tmem_base + lane_off + (uint32_t)(buf * TCB_L + qd * 32), &acc[qd * 16]);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) PBSO_ARRIVE_LEAD(lead_accempty, buf);
                // the FP32 running sum is bounded in CHUNKS, not in units: an object with thousands of modes has hundreds of
                // chunks per unit, and 4096 of them in one FP32 sum cost a factor 2 in max-abs error (1.3e-6 at 8192 modes)
                chunks_held += 2;
                if (chunks_held >= flush_chunks && ch + 2 < cpu) flush(un);
            }
            if (un.flush || chunks_held >= flush_chunks) flush(un);
        }
    } else {
        // warps 20-23: the scheduler prefers the highest warp id of a sub-partition, and the MMA issuers are the warps the
        // whole pipeline waits for
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp < 22) {
          if (rank == 0) {
            // ---------------- MMA issuers: warp 20 takes the even chains (accumulator 0), warp 21 the odd ones ----------------
            // A chain = two K chunks of a unit (one at an odd tail): the 16 small products of both chunks first (hi*lo, lo*hi
            // -- their accumulate-truncation is relative to a 2^-11 smaller sum), then the 8 hi*hi MMAs on top, ONE accumulator,
            // drained once: reading TMEM back costs ~1000 cycles per 128 x 128 accumulator (64 B/clk), more than one chunk's
            // 768 cycles of MMAs, so per-chunk draining would bound the kernel.  Issuing is serial work of one thread; two
            // issuers keep the tensor pipe's queue fed while one of them is between chains.  The whole warp runs the loop
            // so that every operand stays warp-uniform.
            constexpr uint32_t idesc = umma_idesc_tf32(TCB_ROWS * PAIR, TCB_L);
            const uint64_t desc0 = umma_desc_k_sw128(smem_u32(smem));
            const uint32_t cpp = (uint32_t)(cpu + 1) >> 1;                              // chains per unit
            const uint32_t n_chains = (uint32_t)(u1 - u0) * cpp;
            const uint32_t par = (uint32_t)(warp - 20);
            const uint32_t acc = tmem_base + par * TCB_L;
#pragma unroll 1
            for (uint32_t g = par; g < n_chains; g += 2) {
                const uint32_t ui = g / cpp, ci = g - ui * cpp;
                const uint32_t q0 = ui * (uint32_t)cpu + 2 * ci;
                const uint32_t nc = (2 * ci + 1 < (uint32_t)cpu) ? 2u : 1u;
                PBSO_TW(0, mbar_wait(&acc_empty[par], ((g >> 1) & 1) ^ 1));
                PBSO_TW(1, mbar_wait(&b_full[q0 & 3], (q0 >> 2) & 1));
                if (nc == 2) PBSO_TW(1, mbar_wait(&b_full[(q0 + 1) & 3], ((q0 + 1) >> 2) & 1));
                tcgen05_fence_after();
                if (elect_one()) {
                    for (uint32_t i = 0; i < nc; ++i) {
                        const uint32_t st = (q0 + i) & 3;
                        const uint32_t a_hi = tmem_base + 256 + st * 64, a_lo = a_hi + 32;
                        const uint64_t dBh = desc0 + (uint64_t)(st * (TCB_BT_BYTES >> 4)), dBl = dBh + (TCB_TILE_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_ts<0, PAIR>(acc, a_hi + 8 * k, dBl + 2 * k, idesc, (i | (uint32_t)k) ? 1u : 0u);
                            umma_ts<0, PAIR>(acc, a_lo + 8 * k, dBh + 2 * k, idesc, 1u);
                        }
                    }
                    for (uint32_t i = 0; i < nc; ++i) {
                        const uint32_t st = (q0 + i) & 3;
                        const uint32_t a_hi = tmem_base + 256 + st * 64;
                        const uint64_t dBh = desc0 + (uint64_t)(st * (TCB_BT_BYTES >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ts<0, PAIR>(acc, a_hi + 8 * k, dBh + 2 * k, idesc, 1u);
                    }
                    if (PAIR == 2) {
                        for (uint32_t i = 0; i < nc; ++i) umma_commit_2cta(&b_empty[(q0 + i) & 3], 3);
                        umma_commit_2cta(&acc_full[par], 3);
                    } else {
                        for (uint32_t i = 0; i < nc; ++i) umma_commit(&b_empty[(q0 + i) & 3]);
                        umma_commit(&acc_full[par]);
                    }
                }
                __syncwarp();
            }
          }
        } else if (warp == 23) {
            // ---------------- table loader: a chunk's 5 KB table block and its 16 tile-start states, NT chunks ahead ----------------
            if (lane == 0) {
                uint32_t q = 0;
#pragma unroll 1
                for (int u = u0; u < u1; ++u) {
                    const Unit un = units[u];
                    const uint8_t* tsrc = tab + (size_t)(un.obj - obj0) * cpu * TCB_TABG_BYTES;
                    const float2* vsrc = V + (size_t)un.src[rank] * mp;
#pragma unroll 1
                    for (int ch = 0; ch < cpu; ++ch, ++q) {
                        const uint32_t ts = q % TCB_NT;
                        PBSO_TW(0, mbar_wait_relaxed(&t_empty[ts], ((q / TCB_NT) & 1) ^ 1));
                        mbar_expect_tx(&t_full[ts], TCB_TAB_BYTES);
                        const uint32_t tdst = smem_u32(tabs) + ts * TCB_TAB_BYTES;
                        bulk_g2s(tdst, tsrc + (size_t)ch * TCB_TABG_BYTES, TCB_TABG_BYTES, &t_full[ts]);
                        bulk_g2s(tdst + TCB_TABG_BYTES, vsrc + ch * TCB_KMODES, 128, &t_full[ts]);
                    }
                }
            }
        } else if (warp == 22) {
            // ---------------- seed warp: X4[blk][m] = v W^(16 blk), the state at the start of every 16-row block ----------------
            const int m_l = lane & 15, half = lane >> 4;
            uint32_t q = 0;
#pragma unroll 1
            for (int u = u0; u < u1; ++u) {
                const int re = units[u].re[rank];
#pragma unroll 1
                for (int ch = 0; ch < cpu; ++ch, ++q) {
                    const uint32_t ts = q % TCB_NT, ss = q % TCB_NS;
                    PBSO_TW(0, mbar_wait(&t_full[ts], (q / TCB_NT) & 1));
                    const uint32_t tP = smem_u32(tabs) + ts * TCB_TAB_BYTES, tR = tP + TCB_TAB_R, tV = tP + TCB_TABG_BYTES;
                    const float2 v = lds_f2(tV + m_l * 8);
                    float2 x[4];
                    if (re < 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) x[i] = cmulf(v, lds_f2(tP + (4 * half + i) * 128 + m_l * 8));
                    } else {
                        // impulse unit: zero before the impulse row `re`; block ae starts AT the impulse (its rows are
                        // shifted by be in the generators); later blocks start at u W^(16 (blk - ae) - be)
                        const int ae = re >> 4, be = re & 15;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int blk = 4 * half + i;
                            float2 xx = make_float2(0.f, 0.f);
                            if (blk == ae) xx = v;
                            else if (blk > ae) {
                                const int d = 16 * (blk - ae) - be;
                                xx = v;
                                if (d >> 4) xx = cmulf(xx, lds_f2(tP + (d >> 4) * 128 + m_l * 8));
                                if (d & 15) xx = cmulf(xx, lds_f2(tR + (d & 15) * TCB_RSTRIDE + m_l * 8));
                            }
                            x[i] = xx;
                        }
                    }
                    PBSO_TW(1, mbar_wait_relaxed(&s_empty[ss], ((q / TCB_NS) & 1) ^ 1));
                    const uint32_t sX = smem_u32(seeds) + ss * TCB_SEED_BYTES;
#pragma unroll
                    for (int i = 0; i < 4; ++i) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sX + (4 * half + i) * 128 + m_l * 8), "f"(x[i].x), "f"(x[i].y) : "memory");
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&t_empty[ts]); mbar_arrive(&s_full[ss]); }
                }
            }
        }
    }
    if (profiling && lane == 0) {
        unsigned long long* o = prof + warp * 8;
        o[0] = (unsigned long long)(clock64() - t_role0); o[1] = pw[0]; o[2] = pw[1]; o[3] = pw[2]; o[4] = pw[3];
    }
#undef PBSO_TW
#undef PBSO_ARRIVE_LEAD
    tcgen05_fence_before();
    if (PAIR == 2) cluster_sync_all(); else __syncthreads();          // the peer's tensor core may still read this CTA's operands
    if (warp == 20) {
        if (PAIR == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCB_TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCB_TMEM_COLS));
    }
}

}  // namespace

namespace pbso {

struct TcState {
    uint8_t* tab = nullptr;                         // per-handle operand tables of the resident object batch
    float2* V = nullptr; size_t v_cap = 0;          // state blocks: carrier | impulses
    int v_nit = -1, v_no = -1, v_ne = -1;           // layout the pad entries of V were zeroed for
    Unit* units = nullptr; size_t unit_cap = 0;
    int* cta_first = nullptr; int* ev_obj = nullptr; size_t ev_cap = 0;
    int grid = 0, pair = 1;                         // CTAs per cluster: 2 = tcgen05 cta_group::2 pairs
    int tab_obj0 = -1, tab_nobj = 0, batch_obj = 0; // objects [tab_obj0, tab_obj0 + tab_nobj) have tables resident
    size_t npm = 0;
    // cached unit list
    unsigned ev_ver = ~0u; int list_state = -1; int list_tiles = -1, list_buf = -1, list_obj0 = -1, list_nobj = -1, n_units = 0;
    std::vector<Unit> h_units; std::vector<int> h_first;
    std::vector<int> h_ev_obj;
};

void tc_free(TcState* st) {
    if (!st) return;
    cudaFree(st->tab); cudaFree(st->V); cudaFree(st->units); cudaFree(st->cta_first); cudaFree(st->ev_obj);
    delete st;
}

// unit list of the objects [o0, o0 + no): work items = (window of TCB_WINDOW objects, M-tile or pair of M-tiles), dealt
// round-robin to the CTAs (clusters) in (window, M-tile) order -- the CTAs rendering the M-tiles of one window run at about the same time, so a
// window's table blocks are read from HBM once and from L2 by the rest; a CTA keeps one M-tile's partial mix in
// registers across the units of an item.
static void build_units(TcState* st, const TcArgs& a, int o0, int no, int n_tiles, int n_it, int cpu) {
    // an impulse joins the contraction at the first tile boundary at or after its sample (row = ceil(t_e / 128))
    auto ev_row = [&](int e) { return ((long long)a.h_ev_buf[e] * a.buf_size + TCB_L - 1) / TCB_L; };
    const int PAIR = st->pair, ncl = st->grid / PAIR, n_grp = div_up(n_it, PAIR);
    const unsigned imp_blk0 = (unsigned)((size_t)n_it * no);
    const int e_base = a.h_ev_off[o0];
    const unsigned zero_blk = imp_blk0 + (unsigned)(a.h_ev_off[o0 + no] - e_base);          // idle half of a pair
    // objects per work item and units per flush: 8 for objects of 512 modes (32 K chunks each) and more; objects with fewer modes
    // are windowed so that a flush still closes ~256 chunks (a flush stalls the pipeline for ~10 us whatever the unit length)
    static const int window_env = getenv("PBSO_TC_WINDOW") ? std::max(1, atoi(getenv("PBSO_TC_WINDOW"))) : 0;
    static const int flush_env = getenv("PBSO_TC_FLUSH") ? std::max(1, atoi(getenv("PBSO_TC_FLUSH"))) : 0;
    const int window = window_env ? window_env : std::min(64, std::max(TCB_WINDOW, TCB_FLUSH_CHUNKS / std::max(cpu, 1)));
    const int flush_units = flush_env ? flush_env : std::max(TCB_FLUSH_UNITS, window);
    std::vector<std::vector<Unit>> per(ncl);
    std::vector<int> cur(no);
    long long item = 0;
    for (int w0 = 0; w0 < no; w0 += window) {
        const int w1 = std::min(no, w0 + window);
        for (int o = w0; o < w1; ++o) cur[o] = a.h_ev_off[o0 + o];
        for (int t = 0; t < n_grp; ++t) {
            std::vector<Unit> tmp;
            for (int o = w0; o < w1; ++o) {
                const int e_begin = a.h_ev_off[o0 + o], e_end = a.h_ev_off[o0 + o + 1];
                // per M-tile of the group: carry unit (an earlier impulse exists), then the impulses landing inside it
                // (events are sorted by buffer; cur[o] walks them once across the M-tiles)
                std::vector<std::pair<unsigned, int>> u[2];
                for (int r = 0; r < PAIR; ++r) {
                    const int it = t * PAIR + r;
                    if (it >= n_it) continue;
                    const long long row0 = (long long)it * TCB_ROWS, row1 = row0 + TCB_ROWS;
                    if (a.v0r || (e_begin < e_end && ev_row(e_begin) < row0)) u[r].push_back({(unsigned)((size_t)it * no + o), -1});
                    int& e = cur[o];
                    while (e < e_end && ev_row(e) < row1) {
                        const long long row = ev_row(e);
                        if (row >= row0 && row < n_tiles) u[r].push_back({imp_blk0 + (unsigned)(e - e_base), (int)(row - row0)});
                        ++e;
                    }
                }
                const size_t np = std::max(u[0].size(), u[1].size());
                for (size_t i = 0; i < np; ++i) {
                    Unit un{};
                    un.obj = o0 + o; un.it = t;
                    for (int r = 0; r < 2; ++r) {
                        if (i < u[r].size()) { un.src[r] = u[r][i].first; un.re[r] = (short)u[r][i].second; }
                        else { un.src[r] = zero_blk; un.re[r] = -1; }
                    }
                    tmp.push_back(un);
                }
            }
            if (tmp.empty()) continue;
            std::vector<Unit>& out = per[item++ % ncl];
            int since = 0;
            for (size_t i = 0; i < tmp.size(); ++i) {
                Unit& un = tmp[i];
                un.flush = (short)(++since >= flush_units || (a.d_stems && i + 1 < tmp.size() && tmp[i + 1].obj != un.obj));   // stems: at every object change
                if (un.flush) since = 0;
                out.push_back(un);
            }
            out.back().flush = 1;                                    // the M-tiles change with the item
        }
    }
    st->h_units.clear(); st->h_first.assign(ncl + 1, 0);
    for (int c = 0; c < ncl; ++c) {
        st->h_first[c] = (int)st->h_units.size();
        st->h_units.insert(st->h_units.end(), per[c].begin(), per[c].end());
    }
    st->h_first[ncl] = (int)st->h_units.size();
    st->n_units = (int)st->h_units.size();
}

static double g_inv_gain[64];            // per device; 0 = not calibrated yet
static bool g_calibrating = false;

static int tc_calibrate(int device, int sm_count);

int tc_render(TcState** pst, const TcArgs& a, int* launches) {
    if (!*pst) *pst = new TcState();
    TcState* st = *pst;
    *launches = 0;
    int dev = 0; PBSO_CUDA(cudaGetDevice(&dev));
    if (g_inv_gain[dev & 63] == 0.0 && !g_calibrating) { if (int rc = tc_calibrate(dev, a.sm_count)) return rc; }
    const double inv_gain = g_calibrating ? 1.0 : g_inv_gain[dev & 63];
    const size_t npm = (size_t)a.n_obj * a.n_modes;
    const int cpu = div_up(a.n_modes, TCB_KMODES), mp = cpu * TCB_KMODES;
    const long long n_samples = (long long)a.buf_size * a.n_buffers;
    const int n_tiles = (int)((n_samples + TCB_L - 1) / TCB_L);       // 128-sample tiles; the last one may be cut
    const int n_it = div_up(n_tiles, TCB_ROWS);
    static const int force_pair = getenv("PBSO_TC_PAIR") ? atoi(getenv("PBSO_TC_PAIR")) : 0;
    st->pair = (force_pair == 1 || a.sm_count < 2) ? 1 : 2;
    st->grid = std::max(st->pair, a.sm_count / st->pair * st->pair);
    // ---- objects per batch: the operand tables of a batch stay resident (5 KB per object and 16-mode chunk) ----
    if (st->npm != npm || st->batch_obj == 0) {
        cudaFree(st->tab); st->tab = nullptr; st->tab_obj0 = -1; st->npm = 0;
        size_t free_b = 0, total_b = 0; PBSO_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t per_obj = (size_t)cpu * TCB_TABG_BYTES;
        // PBSO_TC_TABLE_BYTES caps the resident tables (tests use it to force the multi-batch path)
        const double frac = getenv("PBSO_TC_TABLE_FRAC") ? atof(getenv("PBSO_TC_TABLE_FRAC")) : 0.4;
        size_t budget = (size_t)(frac * (double)free_b);
        if (getenv("PBSO_TC_TABLE_BYTES")) budget = std::min(budget, (size_t)atoll(getenv("PBSO_TC_TABLE_BYTES")));
        size_t fit = budget / per_obj;
        if (fit < 1) return set_error(PBSO_ERR_CUDA, "PBSO_PREC_TC3X: not enough device memory for one object's operand tables (%zu bytes)", per_obj);
        st->batch_obj = (int)std::min<size_t>(fit, (size_t)a.n_obj);
        PBSO_CUDA(cudaMalloc(&st->tab, (size_t)st->batch_obj * per_obj));
        st->npm = npm; st->ev_ver = ~0u;
    }
    if ((size_t)std::max(a.n_events, 1) > st->ev_cap) {
        cudaFree(st->ev_obj); st->ev_obj = nullptr; st->ev_cap = 0;
        PBSO_CUDA(cudaMalloc(&st->ev_obj, sizeof(int) * std::max(a.n_events, 1))); st->ev_cap = std::max(a.n_events, 1);
    }
    if (!st->cta_first) PBSO_CUDA(cudaMalloc(&st->cta_first, sizeof(int) * (a.sm_count + 2)));
    static const int ablate = getenv("PBSO_TC_ABLATE") ? atoi(getenv("PBSO_TC_ABLATE")) : 0;
    // FP32 running sums of the epilogue: at most 384 K chunks (12 units of 512 modes: cfg5 measured at 256 / 384 / 512 chunks per
    // flush, same box: 14.95 / 14.41 / 14.58 ms, max-abs 4.58e-7 / 4.50e-7 / 5.05e-7) -- 128 for objects with more than 1024 modes,
    // whose chunks all belong to one impulse response and add coherently (measured, 3 objects x 8192 modes, max-abs of full scale:
    // 8.1e-7 at 256, 5.7e-7 at 128, 4.3e-7 at 64).  A flush is 128 FP64 reductions per thread and stalls the pipeline ~10 us.
    static const int flush_env = getenv("PBSO_TC_FLUSH_CHUNKS") ? std::max(2, atoi(getenv("PBSO_TC_FLUSH_CHUNKS"))) : 0;
    const int flush_chunks = flush_env ? flush_env : (cpu > 64 ? 128 : TCB_FLUSH_CHUNKS);
    static bool attr[64] = {};            // per device: function attributes do not carry across devices
    if (!attr[dev & 63]) {
        PBSO_CUDA(cudaFuncSetAttribute(k_batch_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM));
        PBSO_CUDA(cudaFuncSetAttribute(k_batch_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM));
        attr[dev & 63] = true;
    }
    const int n_batches = div_up(a.n_obj, st->batch_obj);
    for (int bi = 0; bi < n_batches; ++bi) {
        const int o0 = bi * st->batch_obj, no = std::min(st->batch_obj, a.n_obj - o0);
        // operand tables (depend on the poles only: built once when every object fits in one batch)
        if (st->tab_obj0 != o0 || st->tab_nobj != no) {
            k_tc_tabs<<<(unsigned)(((size_t)no * mp + 255) / 256), 256, 0, a.stream>>>(no, a.n_modes, cpu, o0, a.lneps, a.theta, st->tab);
            PBSO_CUDA(cudaGetLastError());
            st->tab_obj0 = o0; st->tab_nobj = no; ++*launches;
        }
        // unit list (depends on the impulse script, the render length and the batch)
        const bool relist = st->ev_ver != a.ev_ver || st->list_state != (a.v0r ? 1 : 0) + (a.d_stems ? 2 : 0) || st->list_tiles != n_tiles || st->list_buf != a.buf_size || st->list_obj0 != o0 || st->list_nobj != no;
        if (relist) {
            build_units(st, a, o0, no, n_tiles, n_it, cpu);
            if ((size_t)st->n_units > st->unit_cap) {
                cudaFree(st->units); st->units = nullptr; st->unit_cap = 0;
                PBSO_CUDA(cudaMalloc(&st->units, sizeof(Unit) * st->n_units)); st->unit_cap = st->n_units;
            }
            // pageable sources: cudaMemcpyAsync stages them before it returns, the vectors may be rebuilt right away
            if (st->n_units) PBSO_CUDA(cudaMemcpyAsync(st->units, st->h_units.data(), sizeof(Unit) * st->n_units, cudaMemcpyHostToDevice, a.stream));
            PBSO_CUDA(cudaMemcpyAsync(st->cta_first, st->h_first.data(), sizeof(int) * (st->grid / st->pair + 1), cudaMemcpyHostToDevice, a.stream));
            if (a.n_events > 0) {
                st->h_ev_obj.resize(a.n_events);
                for (int o = 0; o < a.n_obj; ++o) for (int e = a.h_ev_off[o]; e < a.h_ev_off[o + 1]; ++e) st->h_ev_obj[e] = o;
                PBSO_CUDA(cudaMemcpyAsync(st->ev_obj, st->h_ev_obj.data(), sizeof(int) * a.n_events, cudaMemcpyHostToDevice, a.stream));
            }
            st->ev_ver = a.ev_ver; st->list_state = (a.v0r ? 1 : 0) + (a.d_stems ? 2 : 0); st->list_tiles = n_tiles; st->list_buf = a.buf_size; st->list_obj0 = o0; st->list_nobj = no;
        }
        // state blocks: carrier [n_it][no] | impulses of the batch
        const int e0 = a.h_ev_off[o0], ne = a.h_ev_off[o0 + no] - e0;
        if (st->n_units == 0) {
            // nothing for the contraction: silence (the mix is already zeroed) -- or impulses that land inside the LAST tile
            // of the render, whose samples are all head samples
            if (ne > 0 && a.buf_size % TCB_L != 0) { if (int rc = batch_event_heads(a, e0, ne, st->ev_obj)) return rc; ++*launches; }
            continue;
        }
        const size_t vneed = ((size_t)n_it * no + ne + 1) * mp;           // + one zero block: the idle half of a pair
        if (vneed > st->v_cap) {
            cudaFree(st->V); st->V = nullptr; st->v_cap = 0;
            PBSO_CUDA(cudaMalloc(&st->V, sizeof(float2) * vneed)); st->v_cap = vneed;
            st->v_nit = -1;
        }
        if (st->v_nit != n_it || st->v_no != no || st->v_ne != ne) {
            if (mp != a.n_modes) PBSO_CUDA(cudaMemsetAsync(st->V, 0, sizeof(float2) * vneed, a.stream));     // pad modes stay zero
            else PBSO_CUDA(cudaMemsetAsync(st->V + (vneed - mp), 0, sizeof(float2) * mp, a.stream));          // the zero block
        }
        st->v_nit = n_it; st->v_no = no; st->v_ne = ne;
        k_tc_carrier<<<(unsigned)(((size_t)no * a.n_modes + 255) / 256), 256, 0, a.stream>>>(no, a.n_modes, n_it, a.buf_size, mp, o0, a.lneps, a.theta, a.c3, a.cot,
                                                                                       a.trans, a.d_ev_off, a.d_ev_buf, a.d_ev_space, a.v0r, a.v0i, st->V);
        PBSO_CUDA(cudaGetLastError());
        ++*launches;
        if (ne > 0) {
            k_tc_impulse<<<(unsigned)(((size_t)ne * a.n_modes + 255) / 256), 256, 0, a.stream>>>(a.n_modes, mp, e0, ne, a.buf_size, st->ev_obj, a.d_ev_buf, a.lneps, a.theta, a.c3, a.cot, a.trans,
                                                                                              a.d_ev_space, st->V + (size_t)n_it * no * mp);
            PBSO_CUDA(cudaGetLastError());
            ++*launches;
            if (a.buf_size % TCB_L != 0) {                            // impulses inside a tile: their samples up to the next boundary
                if (int rc = batch_event_heads(a, e0, ne, st->ev_obj)) return rc;
                ++*launches;
            }
        }
        static const bool want_prof = getenv("PBSO_TC_PROF") != nullptr;
        unsigned long long* d_prof = nullptr;
        if (want_prof && !g_calibrating) { PBSO_CUDA(cudaMalloc(&d_prof, sizeof(unsigned long long) * 24 * 8)); PBSO_CUDA(cudaMemsetAsync(d_prof, 0, sizeof(unsigned long long) * 24 * 8, a.stream)); }
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(st->grid); cfg.blockDim = dim3(TCB_THREADS); cfg.dynamicSmemBytes = TCB_SMEM; cfg.stream = a.stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = st->pair; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (st->pair == 2)
                PBSO_CUDA(cudaLaunchKernelEx(&cfg, k_batch_tc<2>, a.n_modes, n_tiles, n_samples, (const int*)st->cta_first, (const Unit*)st->units, (const uint8_t*)st->tab,
                                             (const float2*)st->V, o0, a.d_mix, a.d_stems, inv_gain, ablate, d_prof, flush_chunks));
            else
                PBSO_CUDA(cudaLaunchKernelEx(&cfg, k_batch_tc<1>, a.n_modes, n_tiles, n_samples, (const int*)st->cta_first, (const Unit*)st->units, (const uint8_t*)st->tab,
                                             (const float2*)st->V, o0, a.d_mix, a.d_stems, inv_gain, ablate, d_prof, flush_chunks));
        }
        if (d_prof) {
            unsigned long long h[24 * 8];
            PBSO_CUDA(cudaMemcpyAsync(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost, a.stream));
            PBSO_CUDA(cudaStreamSynchronize(a.stream));
            cudaFree(d_prof);
            const long long chunks = std::max(1LL, (long long)(st->h_first[1] - st->h_first[0]) * cpu);
            static const char* role[24] = {"A0.0", "A0.1", "A0.2", "A0.3", "A1.0", "A1.1", "A1.2", "A1.3", "B0.0", "B0.1", "B0.2", "B0.3",
                                           "B1.0", "B1.1", "B1.2", "B1.3", "epi0", "epi1", "epi2", "epi3", "MMA0", "MMA1", "seed", "load"};
            fprintf(stderr, "[tc prof] CTA 0: %lld chunks; cycles per chunk: total | wait0 wait1 wait2 wait3\n", chunks);
            for (int w = 0; w < 24; ++w)
                fprintf(stderr, "[tc prof] %-5s %8.1f | %8.1f %8.1f %8.1f %8.1f\n", role[w], (double)h[w * 8] / chunks, (double)h[w * 8 + 1] / chunks,
                        (double)h[w * 8 + 2] / chunks, (double)h[w * 8 + 3] / chunks, (double)h[w * 8 + 4] / chunks);
        }
        PBSO_CUDA(cudaGetLastError());
        ++*launches;
    }
    return PBSO_OK;
}

// ---- gain calibration -----------------------------------------------------------------------------------------
// The tensor core adds into its FP32 accumulator with truncation; over a chain of 8 hi*hi MMAs that is a coherent
// gain error of about -2e-7 which belongs to the hardware, not to the workload.  Measured once per device: a
// synthetic batch (256 modes spread like SURVEY 8(d)'s, both materials, one impulse per object) rendered by
// k_batch_tc with gain 1 and by the FP64 direct-form kernel; gain = <y_tc, y_64> / <y_64, y_64>.
static int tc_calibrate(int device, int sm_count) {
    const char* fixed = getenv("PBSO_TC_GAIN");
    if (fixed) { g_inv_gain[device & 63] = 1.0 / atof(fixed); return PBSO_OK; }
    g_calibrating = true;
    const int n_obj = std::max(64, sm_count), n_modes = 256, buf = 256, n_buf = 128;     // two M-tiles of 128 tiles
    const double h = 1.0 / 44100.0;
    std::vector<double> a((size_t)n_obj * n_modes), b(a.size()), tr(a.size()), sp(a.size());
    unsigned s = 2463534242u;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return (double)(s >> 8) / 16777216.0; };
    const double two_pi = 6.283185307179586476925286766559;
    for (int o = 0; o < n_obj; ++o)
        for (int m = 0; m < n_modes; ++m) {
            const double f = 80.0 * std::pow(18000.0 / 80.0, ((double)m + rnd()) / n_modes), om = two_pi * f;
            const double alpha = (o & 1) ? 30.0 : 1.0, beta = (o & 1) ? 5e-7 : 1e-7;
            const double xi = 0.5 * (alpha / om + beta * om);
            a[(size_t)o * n_modes + m] = 2.0 * xi * om; b[(size_t)o * n_modes + m] = om * om;
            tr[(size_t)o * n_modes + m] = 0.1 + rnd(); sp[(size_t)o * n_modes + m] = 2.0 * rnd() - 1.0;
        }
    std::vector<int> obj(n_obj), bufi(n_obj);
    for (int o = 0; o < n_obj; ++o) { obj[o] = o; bufi[o] = (int)(rnd() * 40.0); }
    pbso_batch* bt = nullptr;
    int rc = pbso_batch_create(n_obj, n_modes, h, a.data(), b.data(), &bt);
    std::vector<double> y64((size_t)buf * n_buf), ytc(y64.size());
    if (!rc) rc = pbso_batch_set_transfer(bt, tr.data());
    if (!rc) rc = pbso_batch_set_impulses(bt, n_obj, obj.data(), bufi.data(), sp.data());
    if (!rc) rc = pbso_batch_render_mix(bt, buf, n_buf, PBSO_PREC_F64, 0, y64.data());
    if (!rc) rc = pbso_batch_render_mix(bt, buf, n_buf, PBSO_PREC_TC3X, 0, ytc.data());
    pbso_batch_destroy(bt);
    g_calibrating = false;
    if (rc) return rc;
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < y64.size(); ++i) { num += ytc[i] * y64[i]; den += y64[i] * y64[i]; }
    const double gain = den > 0.0 ? num / den : 1.0;
    if (!(gain > 0.999 && gain < 1.001)) return set_error(PBSO_ERR_CUDA, "tensor-core gain calibration out of range: %.9f", gain);
    g_inv_gain[device & 63] = 1.0 / gain;
    return PBSO_OK;
}
double tc_gain(int device) { return g_inv_gain[device & 63] != 0.0 ? 1.0 / g_inv_gain[device & 63] : 0.0; }

}  // namespace pbso
