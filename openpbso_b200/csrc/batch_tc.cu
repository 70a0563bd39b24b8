// =============================================================================
// batch_tc.cu -- offline batch synthesis on the 5th-generation tensor cores (PBSO_PREC_TC3X).
//
// Same job as k_batch_pow (batch.cu): every object is a ModalSolver::step loop (modal_solver.h:181-276) driven
// by PointForce messages, mixed down.  The pole-power form of the recurrence (modal_integrator.h:109-110)
//     q_m[(i L) + j] = Re v_m(i) * Im(w_m^j) + Im v_m(i) * Re(w_m^j),      v_m(i) = state at the start of tile i
// summed over modes with the transfer T_m (modal_solver.h:267-269) and over objects is ONE matrix product
//     Y[i][j] = sum_k A[i][k] B[j][k],    k = (object, mode, component),
//     A[i][(o,m,0..1)] = (Re, Im) v_{o,m}(i)          -- tile-start states  ("advance each mode with z^k")
//     B[j][(o,m,0..1)] = T_{o,m} (Im, Re) w_{o,m}^j   -- precomputed pole powers
// with i = tile index in time (L = 128 samples per tile), j = offset inside the tile.  For cfg5 that is a
// [3446 x 128] output contracted over K = 4096*512*2 = 4.2 M.  Neither operand ever exists in HBM: both are
// generated on the SM, straight into the UMMA shared-memory layout, and consumed by tcgen05.mma kind::tf32 with
// accumulators in TMEM.  Generation is FP32 (packed fma.rn.f32x2 complex products) from three-level tables of
// pole powers, row r = 16 blk + 4 t + c:  A[r] = v_base * W^(16 blk) * W^(4t) * W^c  with the table factors
// computed in FP64 and rounded once (k_tc_tables), and v_base from the FP64 carrier pass (k_tc_carrier).  The
// FP64 and F2F pipes are far too narrow to generate operands at tensor-core speed (first version: 3.9 k
// cycles per K chunk against 768 of MMA time), FP32 products of exactly-rounded factors hold ~1.5e-7.
//
// Precision ("3xTF32"): every FP32 element x is split x = hi + lo (hi = x truncated to TF32 -- the tensor core
// ignores the low 13 mantissa bits itself, so the raw x is stored as the hi operand -- and lo = x - hi rounded to
// nearest TF32), and three MMAs form hi*hi + hi*lo + lo*hi.  hi*hi goes to a "main" TMEM accumulator, the two
// small products to a separate "small" accumulator: the tensor core truncates when it adds into the FP32
// accumulator (~2.5e-8 relative per accumulating MMA, a coherent gain error), so main chains are kept to
// CHAIN*4 MMAs and promoted into FP32 registers with round-to-nearest adds (two-level accumulation); the
// registers are added to the FP64 mix every TCB_FLUSH_UNITS units (a longer FP32 running sum would lose
// 2.4e-8 sqrt(adds)), with the known mean truncation bias compensated at that point.
//
// Work decomposition.  M-tile = 128 consecutive time tiles (16 384 samples).  A unit is (M-tile, object) --
// the object's state at the M-tile start comes from the FP64 carrier pass k_tc_carrier -- or (M-tile, object,
// impulse) for an impulse landing inside the M-tile (rows before it are zero; linear superposition).  Units
// are sorted by M-tile and split evenly over one persistent CTA per SM; a CTA keeps its [128 x 128] partial
// mix in registers across units and adds it to the FP64 mix (RED.64) only when the M-tile changes.
//
// CTA = 16 warps: warp 0 issues the MMAs; warps 4-7 drain TMEM (epilogue); warps 8-11 generate A, warps 12-15
// generate B (thread = (mode of the 16-mode K chunk, 16-row block)); setmaxnreg moves registers from the idle
// warps to the epilogue.  Stage = K chunk of 16 modes: A_hi A_lo B_hi B_lo, [128 rows][32 fp32] each, 128-byte
// swizzle, K-major (64 KB; 3 stages).
// =============================================================================
#include "common.cuh"
#include "umma.cuh"
#include "batch_tc.cuh"
#include <algorithm>
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
constexpr int TCB_THREADS = 512;
constexpr int TCB_TMEM_COLS = 512;         // main[2] | small[2], 128 columns each
constexpr int TCB_FLUSH_UNITS = 4;         // units between FP64 flushes of the register accumulators

struct Unit { int it, obj, ev, pad; };     // ev < 0: carry unit (state from Vbase); else impulse index

struct Cplx { double x, y; };
__device__ __forceinline__ Cplx cmul(const Cplx a, const Cplx b) {
    Cplx r;
    r.x = fma(-a.y, b.y, a.x * b.x);
    r.y = fma(a.x, b.y, a.y * b.x);
    return r;
}

// ---- static per-(object, mode) tables of pole powers, FP64-computed, rounded once to FP32 --------------------
// tab[i][0..7] = P^(16 blk) (x T for operand B), [8..10] = P^4, P^8, P^12, [11..13] = P, P^2, P^3 with
// P = w^L (operand A: tile-to-tile) or w (operand B: sample-to-sample).  Operand B holds (Im, Re) = i conj(z) in
// its K columns; i conj(z1 z2) = (i conj z1) conj(z2), so tabB stores the block starts swapped and the step
// powers conjugated and the generator runs the same recurrence for both operands.
__device__ __forceinline__ float2 pole_pow(double le, double th, double k, double scale) {
    double s, c;
    sincos(k * th, &s, &c);
    const double e = scale * exp(k * le);
    return make_float2((float)(e * c), (float)(e * s));
}
// Layout: [object][K chunk][entry 0..15][mode within the chunk 0..15] float2 -- one 2 KB block per K chunk, which
// the kernel's loader thread brings into shared memory with a single bulk copy; entry-major so that the 16 lanes
// of a half-warp (= 16 modes) read 128 consecutive bytes.  Modes past n_modes are zero.
__global__ void k_tc_tables(int n_obj, int n_modes, int cpu, const double* __restrict__ lneps, const double* __restrict__ theta,
                            const double* __restrict__ trans, float2* __restrict__ tabA, float2* __restrict__ tabB) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;       // (object, padded mode)
    const size_t n = (size_t)n_obj * cpu * TCB_KMODES;
    if (i >= n) return;
    const int o = (int)(i / ((size_t)cpu * TCB_KMODES)), mp = (int)(i % ((size_t)cpu * TCB_KMODES));
    float2* ta = tabA + ((size_t)o * cpu + mp / TCB_KMODES) * 256 + (mp % TCB_KMODES);
    float2* tb = tabB + ((size_t)o * cpu + mp / TCB_KMODES) * 256 + (mp % TCB_KMODES);
    if (mp >= n_modes) {
#pragma unroll
        for (int e = 0; e < 16; ++e) ta[e * 16] = tb[e * 16] = make_float2(0.f, 0.f);
        return;
    }
    const size_t src = (size_t)o * n_modes + mp;
    const double le = lneps[src], th = theta[src], T = trans[src];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        ta[a * 16] = pole_pow(le, th, 16.0 * a * TCB_L, 1.0);
        { const float2 z = pole_pow(le, th, 16.0 * a, T); tb[a * 16] = make_float2(z.y, z.x); }
    }
#pragma unroll
    for (int t = 1; t < 4; ++t) {
        ta[(7 + t) * 16] = pole_pow(le, th, 4.0 * t * TCB_L, 1.0);
        ta[(10 + t) * 16] = pole_pow(le, th, (double)t * TCB_L, 1.0);
        { const float2 z = pole_pow(le, th, 4.0 * t, 1.0); tb[(7 + t) * 16] = make_float2(z.x, -z.y); }
        { const float2 z = pole_pow(le, th, (double)t, 1.0); tb[(10 + t) * 16] = make_float2(z.x, -z.y); }
    }
    ta[14 * 16] = ta[15 * 16] = tb[14 * 16] = tb[15 * 16] = make_float2(0.f, 0.f);
}

// ---- FP64 carrier: state of every (object, mode) at the start of every M-tile -------------------------------
// Vbase[it] excludes impulses landing at rows >= it*128 (those are impulse units of M-tile it).
__global__ void k_tc_carrier(int n_obj, int n_modes, int n_it, int tiles_per_buf_num, int tiles_per_buf_den,
                             const double* __restrict__ lneps, const double* __restrict__ theta,
                             const double* __restrict__ c3a, const double* __restrict__ cota,
                             const int* __restrict__ ev_off, const int* __restrict__ ev_buf,
                             const double* __restrict__ ev_space, float2* __restrict__ Vbase) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t npm = (size_t)n_obj * n_modes;
    if (idx >= npm) return;
    const int o = (int)(idx / n_modes), m = (int)(idx % n_modes);
    const double le = lneps[idx], th = theta[idx];
    double s, c;
    sincos((double)(TCB_L * TCB_ROWS) * th, &s, &c);
    const double e128 = exp((double)(TCB_L * TCB_ROWS) * le);
    const Cplx Wm{e128 * c, e128 * s};
    const double inji = c3a[idx], injr = inji * cota[idx];
    Cplx v{0.0, 0.0};
    int e = ev_off[o];
    const int e_end = ev_off[o + 1];
    for (int it = 0; it < n_it; ++it) {
        Vbase[(size_t)it * npm + idx] = make_float2((float)v.x, (float)v.y);
        v = cmul(v, Wm);
        const long long row_end = (long long)(it + 1) * TCB_ROWS;
        while (e < e_end) {
            const long long row = (long long)ev_buf[e] * tiles_per_buf_num / tiles_per_buf_den;
            if (row >= row_end) break;
            const double k = (double)(row_end - row) * TCB_L;               // samples from the impulse to the next M-tile start
            const double sp = ev_space[(size_t)e * n_modes + m];
            sincos(k * th, &s, &c);
            const double ek = exp(k * le);
            v.x += sp * (injr * (ek * c) - inji * (ek * s));
            v.y += sp * (injr * (ek * s) + inji * (ek * c));
            ++e;
        }
    }
}

// ---- operand generation ------------------------------------------------------------------------------------
// Complex numbers are (re, im) packed in one 64-bit register; products use mul/fma.rn.f32x2 (SASS FMUL2/FFMA2).
typedef unsigned long long c32;                                          // packed (re, im)
__device__ __forceinline__ c32 pk(float re, float im) { c32 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im)); return r; }
__device__ __forceinline__ void upk(c32 v, float& re, float& im) { asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(v)); }
__device__ __forceinline__ c32 pk2(float2 v) { return pk(v.x, v.y); }
// p * q = pr (qr, qi) + pi (-qi, qr): q is passed with its rotation qrot = i q = (-qi, qr) so that both packed
// instructions take p's components as broadcast scalars (no register-pair shuffling in the inner loops).
__device__ __forceinline__ c32 cmulf(c32 p, c32 q, c32 qrot) {
    float pr, pi; upk(p, pr, pi);
    const c32 pa = pk(pr, pr), pb = pk(pi, pi);
    c32 r;
    asm("{\n\t.reg .b64 t;\n\tmul.rn.f32x2 t, %1, %2;\n\tfma.rn.f32x2 %0, %3, %4, t;\n\t}" : "=l"(r) : "l"(pa), "l"(q), "l"(pb), "l"(qrot));
    return r;
}
__device__ __forceinline__ c32 rot(c32 v) { float a, b; upk(v, a, b); return pk(-b, a); }

template <int SPLIT>
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi_bits, uint32_t& lo_bits) {
    const uint32_t xb = __float_as_uint(x);
    if (SPLIT == 2) {                                                  // hi, lo both rounded to nearest
        const uint32_t h = (xb + 0x1000u) & 0xFFFFE000u;
        hi_bits = h;
        lo_bits = __float_as_uint(x - __uint_as_float(h)) + 0x1000u;
    } else {
        const uint32_t h = xb & 0xFFFFE000u;                           // what the tensor core reads of x
        hi_bits = xb;
        const float l = x - __uint_as_float(h);
        lo_bits = SPLIT == 1 ? __float_as_uint(l) + 0x1000u : __float_as_uint(l);   // +half ulp: truncation -> RN
    }
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// row b of the thread's 16-row block, K columns (2 m_l, 2 m_l + 1):
// byte offset (r/8)*1024 + (r%8)*128 + (((m_l/2) ^ (r%8)) * 16) + (m_l%2)*8 with r = 16 blk + b
template <int SPLIT>
__device__ __forceinline__ void store_row(uint32_t tile_hi, uint32_t tile_lo, uint32_t base, int b, c32 v) {
    float re, im; upk(v, re, im);
    uint32_t h0, l0, h1, l1;
    split_tf32<SPLIT>(re, h0, l0);
    split_tf32<SPLIT>(im, h1, l1);
    const uint32_t off = (base ^ ((uint32_t)(b & 7) * 16u)) + (uint32_t)(b >> 3) * 1024u + (uint32_t)(b & 7) * 128u;
    sts_v2(tile_hi + off, h0, h1);
    sts_v2(tile_lo + off, l0, l1);
}

// Fast path: rows 4t + c = x * Rt[t] * Rc[c]  (Rt[0] = Rc[0] = 1).  t123 / c123 hold entries 8..13 of the table.
template <int SPLIT>
__device__ __forceinline__ void gen_block(uint32_t tile_hi, uint32_t tile_lo, int blk, int m_l, c32 x, const c32 (&rt)[3], const c32 (&rc)[3]) {
    const uint32_t base = (uint32_t)blk * 2048u + (uint32_t)(m_l >> 1) * 16u + (uint32_t)(m_l & 1) * 8u;
    const c32 rcs[3] = {rot(rc[0]), rot(rc[1]), rot(rc[2])};   // i rc
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const c32 pt = t == 0 ? x : cmulf(x, rt[t - 1], rot(rt[t - 1]));
        store_row<SPLIT>(tile_hi, tile_lo, base, 4 * t, pt);
#pragma unroll
        for (int c = 1; c < 4; ++c) store_row<SPLIT>(tile_hi, tile_lo, base, 4 * t + c, cmulf(pt, rc[c - 1], rcs[c - 1]));
    }
}

// A operand from TMEM, B from shared memory:  D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Shared-memory map of k_batch_tc: B stages (B_hi | B_lo, 16 KB each), seed ring, barriers.
// TMEM map (512 columns): main[0] 0..127, main[1] 128..255, small 256..383, A stage s at 384 + 64 s (hi 32 | lo 32).
constexpr int TCB_BSTAGES = 4, TCB_ASTAGES = 2, TCB_SEEDS = 4;
constexpr int TCB_BSTAGE_BYTES = 2 * TCB_TILE_BYTES;
constexpr int TCB_RROW = 18 * 8;                                      // R row: 16 powers, a zero entry, pad (16-byte aligned)
constexpr int TCB_SEED_BYTES = TCB_KMODES * 64 + TCB_KMODES * TCB_RROW;   // X[16 modes][8 blk] then R[16 modes][18], float2
constexpr int TCB_TABS = 6, TCB_TAB_BYTES = 2 * 2048;                 // table ring: tabA block | tabB block of a chunk
constexpr int TCB_SMEM_TS = TCB_BSTAGES * TCB_BSTAGE_BYTES + TCB_SEEDS * TCB_SEED_BYTES + TCB_TABS * TCB_TAB_BYTES + 1024 + 512;
constexpr int TCB_SMALL_CHAIN = 8;                                   // chunks per small-accumulator chain

template <int SPLIT, int CHAIN>
__global__ void __launch_bounds__(TCB_THREADS, 1)
k_batch_tc(int n_obj, int n_modes, int n_tiles, const int* __restrict__ cta_first, const Unit* __restrict__ units,
           const float2* __restrict__ tabA, const float2* __restrict__ tabB, const float2* __restrict__ Vbase,
           const double* __restrict__ c3a, const double* __restrict__ cota, const int* __restrict__ ev_row,
           const double* __restrict__ ev_space, double* __restrict__ mix, int flush_units, int ablate) {
    static_assert(TCB_SMALL_CHAIN % CHAIN == 0, "small chains end on main chain ends");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* seeds = smem + TCB_BSTAGES * TCB_BSTAGE_BYTES;
    uint8_t* tabs = seeds + TCB_SEEDS * TCB_SEED_BYTES;
    uint64_t* bars = (uint64_t*)(tabs + TCB_TABS * TCB_TAB_BYTES);
    uint64_t* b_full = bars;                       // [4]  B stage written (4 generator warps)
    uint64_t* b_empty = b_full + TCB_BSTAGES;      // [4]  MMAs reading it retired
    uint64_t* a_full = b_empty + TCB_BSTAGES;      // [2]  A stage stored to TMEM (4 generator warps)
    uint64_t* a_empty = a_full + TCB_ASTAGES;      // [2]
    uint64_t* seed_full = a_empty + TCB_ASTAGES;   // [4]  seeds of a chunk written (1 seed warp)
    uint64_t* seed_empty = seed_full + TCB_SEEDS;  // [4]  consumed (4 A-generator warps)
    uint64_t* acc_full = seed_empty + TCB_SEEDS;   // [2]  main accumulator chain finished
    uint64_t* acc_empty = acc_full + 2;            // [2]  drained (128 epilogue threads)
    uint64_t* small_full = acc_empty + 2;          // [1]
    uint64_t* small_empty = small_full + 1;        // [1]
    uint64_t* tab_full = small_empty + 1;          // [6]  bulk copies of a chunk's table blocks landed (tx bytes)
    uint64_t* tab_empty = tab_full + TCB_TABS;     // [6]  consumed (1 seed warp + 4 B-generator warps)
    uint32_t* tmem_slot = (uint32_t*)(tab_empty + TCB_TABS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's contiguous range of units (sorted by M-tile; ranges of equal estimated cost, built on the host)
    const int u0 = cta_first[blockIdx.x], u1 = cta_first[blockIdx.x + 1];
    const int cpu = (n_modes + TCB_KMODES - 1) / TCB_KMODES;          // K chunks per unit
    const size_t npm = (size_t)n_obj * n_modes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TCB_BSTAGES; ++s) { mbar_init(&b_full[s], 4); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < TCB_ASTAGES; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < TCB_SEEDS; ++s) { mbar_init(&seed_full[s], 1); mbar_init(&seed_empty[s], 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        mbar_init(small_full, 1); mbar_init(small_empty, 128);
        for (int s = 0; s < TCB_TABS; ++s) { mbar_init(&tab_full[s], 1); mbar_init(&tab_empty[s], 5); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCB_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t n_chunks_total = (uint32_t)(u1 - u0) * (uint32_t)cpu;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        if (warp == 0) {
            // ---------------- MMA issuer ----------------
            // The whole warp runs the loop so that every operand stays warp-uniform (uniform registers feed
            // UTCHMMA directly); one elected lane issues.  B descriptors differ only in their 14-bit address field:
            // stage base + tile offset (16 KB -> +1024) + k step (32 B -> +2); A is a TMEM column address.
            constexpr uint32_t idesc = umma_idesc_tf32(TCB_ROWS, TCB_L);
            const uint64_t desc0 = umma_desc_k_sw128(smem_u32(smem));
            uint32_t g = 0, gs = 0;                                    // main / small chain counters
            int ch = 0;
            for (uint32_t q = 0; q < n_chunks_total; ++q) {
                const uint32_t sa = q % TCB_ASTAGES, pha = (q / TCB_ASTAGES) & 1;
                const uint32_t sb = q % TCB_BSTAGES, phb = (q / TCB_BSTAGES) & 1;
                const uint32_t buf = g & 1;
                const bool chain_start = (ch % CHAIN) == 0, chain_end = (ch % CHAIN) == CHAIN - 1 || ch == cpu - 1;
                const bool small_start = (ch % TCB_SMALL_CHAIN) == 0, small_end = (ch % TCB_SMALL_CHAIN) == TCB_SMALL_CHAIN - 1 || ch == cpu - 1;
                if (chain_start) mbar_wait(&acc_empty[buf], ((g >> 1) & 1) ^ 1);
                if (small_start) mbar_wait(small_empty, (gs & 1) ^ 1);
                mbar_wait(&a_full[sa], pha);
                mbar_wait(&b_full[sb], phb);
                tcgen05_fence_after();
                if (elect_one()) {
                    if (!(ablate & 8)) {
                    const uint32_t acc_main = tmem_base + buf * TCB_L, acc_small = tmem_base + 256;
                    const uint32_t a_hi = tmem_base + 384 + sa * 64, a_lo = a_hi + 32;
                    const uint64_t dB = desc0 + (uint64_t)(sb * (TCB_BSTAGE_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dBh = dB + 2 * k, dBl = dBh + (TCB_TILE_BYTES >> 4);
                        umma_tf32_ts(acc_main, a_hi + 8 * k, dBh, idesc, (chain_start && k == 0) ? 0u : 1u);
                        umma_tf32_ts(acc_small, a_hi + 8 * k, dBl, idesc, (small_start && k == 0) ? 0u : 1u);
                        umma_tf32_ts(acc_small, a_lo + 8 * k, dBh, idesc, 1u);
                    }
                    }
                    umma_commit(&a_empty[sa]);
                    umma_commit(&b_empty[sb]);
                    if (chain_end) umma_commit(&acc_full[buf]);
                    if (small_end) umma_commit(small_full);
                }
                __syncwarp();
                if (chain_end) ++g;
                if (small_end) ++gs;
                if (++ch == cpu) ch = 0;
            }
        } else if (warp == 3) {
            // ---------------- loader: one bulk copy per operand per chunk into the table ring ----------------
            if (lane == 0) {
                for (uint32_t q = 0; q < n_chunks_total; ++q) {
                    const uint32_t slot = q % TCB_TABS;
                    mbar_wait(&tab_empty[slot], ((q / TCB_TABS) & 1) ^ 1);
                    const size_t blk_idx = ((size_t)units[u0 + (int)(q / cpu)].obj * cpu + (q % cpu)) * 256;
                    const uint32_t dst = smem_u32(tabs) + slot * TCB_TAB_BYTES;
                    mbar_expect_tx(&tab_full[slot], TCB_TAB_BYTES);
                    bulk_g2s(dst, tabA + blk_idx, 2048, &tab_full[slot]);
                    bulk_g2s(dst + 2048, tabB + blk_idx, 2048, &tab_full[slot]);
                }
            }
        } else if (warp <= 2) {
            // ---------------- seed warps (chunks alternate between warps 1 and 2) ----------------
            // lanes 0-15: X[m][blk] = v_base * W^(16 blk), the state at the start of each 16-row block;
            // lanes 16-31: R[m][j] = W^j = W^(4t) W^c for the 16 rows of a block.
            const int m_l = lane & 15, half = lane >> 4;
            for (uint32_t q = warp - 1; q < n_chunks_total; q += 2) {
                const int u = u0 + (int)(q / cpu), ch = (int)(q % cpu);
                const Unit un = units[u];
                const int m = ch * TCB_KMODES + m_l;
                const bool valid = m < n_modes;
                const size_t idx = (size_t)un.obj * n_modes + (valid ? m : 0);
                const uint32_t tslot = q % TCB_TABS;
                mbar_wait(&tab_full[tslot], (q / TCB_TABS) & 1);
                const uint32_t tA = smem_u32(tabs) + tslot * TCB_TAB_BYTES + m_l * 8;     // entry e at tA + 128 e
                float2 out[16];
                if (half == 0) {
                    c32 ra[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) ra[e] = pk2(lds_f2(tA + 128 * e));
                    if (un.ev < 0) {
                        const c32 vb = valid ? pk2(__ldg(&Vbase[(size_t)un.it * npm + idx])) : pk(0.f, 0.f);
#pragma unroll
                        for (int blk = 0; blk < 8; ++blk) { float a, b; upk(cmulf(vb, ra[blk], rot(ra[blk])), a, b); out[blk] = make_float2(a, b); }
                    } else {
                        // impulse unit: zero before the impulse row `re`; block ae starts AT the impulse (rows are
                        // shifted by be in the row threads); later blocks start at u W^(16 (blk - ae) - be)
                        const int re = ev_row[un.ev] - un.it * TCB_ROWS, ae = re >> 4, be = re & 15;
                        const double inji = c3a[idx], injr = inji * cota[idx];
                        const double sp = valid ? ev_space[(size_t)un.ev * n_modes + m] : 0.0;
                        const c32 uimp = pk((float)(sp * injr), (float)(sp * inji));
#pragma unroll
                        for (int blk = 0; blk < 8; ++blk) {
                            c32 x = pk(0.f, 0.f);
                            if (blk == ae) x = uimp;
                            else if (blk > ae) {
                                const int d = 16 * (blk - ae) - be;
                                x = uimp;
                                if (d >> 4) { const c32 qq = pk2(lds_f2(tA + 128 * (d >> 4))); x = cmulf(x, qq, rot(qq)); }
                                if ((d >> 2) & 3) { const c32 qq = pk2(lds_f2(tA + 128 * (7 + ((d >> 2) & 3)))); x = cmulf(x, qq, rot(qq)); }
                                if (d & 3) { const c32 qq = pk2(lds_f2(tA + 128 * (10 + (d & 3)))); x = cmulf(x, qq, rot(qq)); }
                            }
                            float a, b; upk(x, a, b); out[blk] = make_float2(a, b);
                        }
                    }
                } else {
                    const c32 rt[4] = {pk(1.f, 0.f), pk2(lds_f2(tA + 128 * 8)), pk2(lds_f2(tA + 128 * 9)), pk2(lds_f2(tA + 128 * 10))};
                    const c32 rc[4] = {pk(1.f, 0.f), pk2(lds_f2(tA + 128 * 11)), pk2(lds_f2(tA + 128 * 12)), pk2(lds_f2(tA + 128 * 13))};
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const c32 v = (t == 0) ? rc[c] : (c == 0 ? rt[t] : cmulf(rt[t], rc[c], rot(rc[c])));
                            float a, b; upk(v, a, b); out[4 * t + c] = make_float2(a, b);
                        }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&tab_empty[tslot]);
                const uint32_t slot = q % TCB_SEEDS;
                mbar_wait(&seed_empty[slot], ((q / TCB_SEEDS) & 1) ^ 1);
                const uint32_t sbase = smem_u32(seeds) + slot * TCB_SEED_BYTES;
                if (half == 0) {
                    const uint32_t dst = sbase + m_l * 64;
#pragma unroll
                    for (int i = 0; i < 4; ++i) sts_v4(dst + 16 * i, out[2 * i].x, out[2 * i].y, out[2 * i + 1].x, out[2 * i + 1].y);
                } else {
                    const uint32_t dst = sbase + TCB_KMODES * 64 + m_l * TCB_RROW;
#pragma unroll
                    for (int i = 0; i < 8; ++i) sts_v4(dst + 16 * i, out[2 * i].x, out[2 * i].y, out[2 * i + 1].x, out[2 * i + 1].y);
                    sts_v4(dst + 128, 0.f, 0.f, 0.f, 0.f);             // R[m][16] = 0: rows before an impulse
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&seed_full[slot]);
            }
        }
    } else if (warp < 8) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ---------------- epilogue: promote finished chains into registers, flush to the FP64 mix ----------------
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        float acc[TCB_L];
#pragma unroll
        for (int j = 0; j < TCB_L; ++j) acc[j] = 0.f;
        uint32_t g = 0, gs = 0;
        int since_flush = 0;
        // mean of the tensor core's accumulate-with-truncation (2.5e-8 per main MMA of a chain, measured) and, when
        // lo is not re-centred (SPLIT < 2), of the dropped lo*lo term
        const double gain = 1.0 + 1.0e-7 * CHAIN + (SPLIT == 2 ? 0.0 : 0.6e-7);
        for (int u = u0; u < u1; ++u) {
            const int it = units[u].it;
            for (int ch = 0; ch < cpu; ++ch) {
                const bool chain_end = (ch % CHAIN) == CHAIN - 1 || ch == cpu - 1;
                const bool small_end = (ch % TCB_SMALL_CHAIN) == TCB_SMALL_CHAIN - 1 || ch == cpu - 1;
                if (chain_end) {
                    const int buf = g & 1;
                    mbar_wait(&acc_full[buf], (g >> 1) & 1);
                    tcgen05_fence_after();
                    if (!(ablate & 4))
#pragma unroll
                    for (int qd = 0; qd < TCB_L / 32; ++qd) {
                        uint32_t vm[32];
                        tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(buf * TCB_L + qd * 32), vm);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[qd * 32 + j] += __uint_as_float(vm[j]);
                    }
                    tcgen05_fence_before();
                    mbar_arrive(&acc_empty[buf]);
                    ++g;
                }
                if (small_end) {
                    mbar_wait(small_full, gs & 1);
                    tcgen05_fence_after();
                    if (!(ablate & 4))
#pragma unroll
                    for (int qd = 0; qd < TCB_L / 32; ++qd) {
                        uint32_t vs[32];
                        tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(256 + qd * 32), vs);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[qd * 32 + j] += __uint_as_float(vs[j]);
                    }
                    tcgen05_fence_before();
                    mbar_arrive(small_empty);
                    ++gs;
                }
            }
            ++since_flush;
            const bool flush = (u + 1 == u1) || (units[u + 1].it != it) || since_flush >= flush_units;
            if (flush) {
                since_flush = 0;
                const long long tile = (long long)it * TCB_ROWS + row;
                if (tile < n_tiles) {
                    double* dst = mix + tile * TCB_L;
#pragma unroll
                    for (int j = 0; j < TCB_L; ++j) atomicAdd(dst + j, (double)acc[j] * gain);
                }
#pragma unroll
                for (int j = 0; j < TCB_L; ++j) acc[j] = 0.f;
            }
        }
    } else if (warp < 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        // ---------------- A generators: thread = row (TMEM lane); A[row][2m..2m+1] = X[m][blk] * R[m][j] -----------
        const int row = (warp - 8) * 32 + lane, blk = row >> 4, j = row & 15;
        const uint32_t lane_off = (uint32_t)((warp - 8) * 32) << 16;
        for (uint32_t q = 0; q < n_chunks_total; ++q) {
            const int u = u0 + (int)(q / cpu);
            const Unit un = units[u];
            int jj = j;
            if (un.ev >= 0) {                                         // impulse unit: rows of block ae are shifted by be
                const int re = ev_row[un.ev] - un.it * TCB_ROWS;
                if (blk == (re >> 4)) jj = j - (re & 15);
            }
            const uint32_t slot = q % TCB_SEEDS, sa = q % TCB_ASTAGES;
            mbar_wait(&seed_full[slot], (q / TCB_SEEDS) & 1);
            const uint32_t sX = smem_u32(seeds) + slot * TCB_SEED_BYTES + blk * 8;
            const uint32_t sR = smem_u32(seeds) + slot * TCB_SEED_BYTES + TCB_KMODES * 64 + (jj < 0 ? 16 : jj) * 8;
            uint32_t hi[32], lo[32];
            if (ablate & 2) {
#pragma unroll
                for (int m = 0; m < 32; ++m) hi[m] = lo[m] = 0u;
            } else
#pragma unroll
            for (int m = 0; m < TCB_KMODES; ++m) {
                const float2 x = lds_f2(sX + m * 64), r = lds_f2(sR + m * TCB_RROW);
                float vr, vi; upk(cmulf(pk(x.x, x.y), pk(r.x, r.y), pk(-r.y, r.x)), vr, vi);
                split_tf32<SPLIT>(vr, hi[2 * m], lo[2 * m]);
                split_tf32<SPLIT>(vi, hi[2 * m + 1], lo[2 * m + 1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&seed_empty[slot]);
            mbar_wait(&a_empty[sa], ((q / TCB_ASTAGES) & 1) ^ 1);
            tcgen05_fence_after();
            const uint32_t a_col = tmem_base + lane_off + 384 + sa * 64;
            if (!(ablate & 2)) {
            tmem_st_32x32(a_col, hi);
            tmem_st_32x32(a_col + 32, lo);
            tmem_st_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[sa]);
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        // ---------------- B generators: thread = (mode of the chunk, 16-row block), pole powers T w^j -----------
        const int m_l = lane & 15, blk = (warp - 12) * 2 + (lane >> 4);
        for (uint32_t q = 0; q < n_chunks_total; ++q) {
            const uint32_t sb = q % TCB_BSTAGES, tslot = q % TCB_TABS;
            mbar_wait(&tab_full[tslot], (q / TCB_TABS) & 1);
            const uint32_t tB = smem_u32(tabs) + tslot * TCB_TAB_BYTES + 2048 + m_l * 8;   // entry e at tB + 128 e
            const c32 ra = pk2(lds_f2(tB + 128 * blk));
            const c32 rt[3] = {pk2(lds_f2(tB + 128 * 8)), pk2(lds_f2(tB + 128 * 9)), pk2(lds_f2(tB + 128 * 10))};
            const c32 rc[3] = {pk2(lds_f2(tB + 128 * 11)), pk2(lds_f2(tB + 128 * 12)), pk2(lds_f2(tB + 128 * 13))};
            mbar_wait(&b_empty[sb], ((q / TCB_BSTAGES) & 1) ^ 1);
            const uint32_t st = smem_u32(smem) + sb * TCB_BSTAGE_BYTES;
            if (!(ablate & 1)) gen_block<SPLIT>(st, st + TCB_TILE_BYTES, blk, m_l, ra, rt, rc);
            fence_proxy_async_smem();
            __syncwarp();
            // the table slot is released only here: the loads above are certainly complete once their values have
            // been used (an arrive issued right after the LDS can overtake them in the MIO queue)
            if (lane == 0) { mbar_arrive(&tab_empty[tslot]); mbar_arrive(&b_full[sb]); }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCB_TMEM_COLS));
}

__global__ void k_ev_rows(int n, const int* __restrict__ ev_buf, int num, int den, int* __restrict__ ev_row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ev_row[i] = (int)((long long)ev_buf[i] * num / den);
}

}  // namespace

namespace pbso {

struct TcState {
    float2 *tabA = nullptr, *tabB = nullptr, *Vbase = nullptr;
    Unit* units = nullptr; int* ev_row = nullptr; int* cta_first = nullptr; int grid = 0;
    size_t npm = 0, vbase_cap = 0, units_cap = 0, ev_cap = 0;
    unsigned tables_ver = ~0u, units_ev_ver = ~0u;
    int units_n_it = -1, units_buf_size = -1, n_units = 0;
    std::vector<Unit> h_units;
};

void tc_free(TcState* st) {
    if (!st) return;
    cudaFree(st->tabA); cudaFree(st->tabB); cudaFree(st->Vbase);
    cudaFree(st->units); cudaFree(st->ev_row); cudaFree(st->cta_first);
    delete st;
}

int tc_render(TcState** pst, const TcArgs& a, int* launches) {
    PBSO_REQUIRE(a.buf_size % TCB_L == 0, PBSO_ERR_UNSUPPORTED, "PBSO_PREC_TC3X needs buf_size to be a multiple of 128");
    if (!*pst) *pst = new TcState();
    TcState* st = *pst;
    const size_t npm = (size_t)a.n_obj * a.n_modes;
    const int cpu = div_up(a.n_modes, TCB_KMODES);
    const size_t ntab = (size_t)a.n_obj * cpu * 256;                  // float2 entries per operand table
    const long long n_samples = (long long)a.buf_size * a.n_buffers;
    const int n_tiles = (int)(n_samples / TCB_L);
    const int n_it = div_up(n_tiles, TCB_ROWS);
    const int tpb = a.buf_size / TCB_L;                               // tiles per buffer
    *launches = 0;
    // static tables (depend on a, b, the transfer vectors and L)
    if (st->npm != npm) {
        cudaFree(st->tabA); cudaFree(st->tabB);
        st->tabA = st->tabB = nullptr; st->npm = 0;
        PBSO_CUDA(cudaMalloc(&st->tabA, sizeof(float2) * ntab));
        PBSO_CUDA(cudaMalloc(&st->tabB, sizeof(float2) * ntab));
        st->npm = npm; st->tables_ver = ~0u;
    }
    if (st->tables_ver != a.trans_ver) {
        k_tc_tables<<<(unsigned)(((size_t)a.n_obj * cpu * TCB_KMODES + 255) / 256), 256, 0, a.stream>>>(a.n_obj, a.n_modes, cpu, a.lneps, a.theta, a.trans, st->tabA, st->tabB);
        PBSO_CUDA(cudaGetLastError());
        st->tables_ver = a.trans_ver; ++*launches;
    }
    // unit list (depends on the impulse script and the render length)
    if (st->units_ev_ver != a.ev_ver || st->units_n_it != n_it || st->units_buf_size != a.buf_size) {
        std::vector<Unit>& hu = st->h_units;
        hu.clear();
        std::vector<int> cursor(a.h_ev_off, a.h_ev_off + a.n_obj);    // first event of each object not yet placed
        for (int it = 0; it < n_it; ++it) {
            const long long row0 = (long long)it * TCB_ROWS, row1 = row0 + TCB_ROWS;
            for (int o = 0; o < a.n_obj; ++o) {
                const int e_begin = a.h_ev_off[o], e_end = a.h_ev_off[o + 1];
                if (e_begin < e_end && (long long)a.h_ev_buf[e_begin] * tpb < row0) hu.push_back(Unit{it, o, -1, 0});
                int& e = cursor[o];
                while (e < e_end && (long long)a.h_ev_buf[e] * tpb < row1) {
                    if ((long long)a.h_ev_buf[e] * tpb < n_tiles) hu.push_back(Unit{it, o, e, 0});
                    ++e;
                }
            }
        }
        st->n_units = (int)hu.size();
        if (hu.size() > st->units_cap) {
            cudaFree(st->units); st->units = nullptr; st->units_cap = 0;
            PBSO_CUDA(cudaMalloc(&st->units, sizeof(Unit) * hu.size())); st->units_cap = hu.size();
        }
        if ((size_t)std::max(a.n_events, 1) > st->ev_cap) {
            cudaFree(st->ev_row); st->ev_row = nullptr; st->ev_cap = 0;
            PBSO_CUDA(cudaMalloc(&st->ev_row, sizeof(int) * std::max(a.n_events, 1))); st->ev_cap = std::max(a.n_events, 1);
        }
        if (!hu.empty()) PBSO_CUDA(cudaMemcpyAsync(st->units, hu.data(), sizeof(Unit) * hu.size(), cudaMemcpyHostToDevice, a.stream));
        // contiguous ranges of equal estimated cost, one per CTA: an impulse unit costs more than a carry unit (one
        // of its 16-row blocks takes the general path), and impulse units cluster in the first M-tiles
        static const double imp_w = getenv("PBSO_TC_IMPW") ? atof(getenv("PBSO_TC_IMPW")) : 2.0;
        st->grid = std::max(1, std::min(a.sm_count, st->n_units));
        std::vector<int> first(st->grid + 1, st->n_units);
        double total = 0.0;
        for (const Unit& un : hu) total += un.ev >= 0 ? imp_w : 1.0;
        double acc_cost = 0.0; int c = 0;
        first[0] = 0;
        for (int i = 0; i < st->n_units; ++i) {
            while (c + 1 < st->grid && acc_cost >= total * (c + 1) / st->grid) first[++c] = i;
            acc_cost += hu[i].ev >= 0 ? imp_w : 1.0;
        }
        while (c + 1 <= st->grid) first[++c] = st->n_units;
        if (!st->cta_first) PBSO_CUDA(cudaMalloc(&st->cta_first, sizeof(int) * (a.sm_count + 1)));
        PBSO_CUDA(cudaMemcpyAsync(st->cta_first, first.data(), sizeof(int) * (st->grid + 1), cudaMemcpyHostToDevice, a.stream));
        if (a.n_events > 0) {
            k_ev_rows<<<div_up(a.n_events, 256), 256, 0, a.stream>>>(a.n_events, a.d_ev_buf, tpb, 1, st->ev_row);
            PBSO_CUDA(cudaGetLastError());
        }
        PBSO_CUDA(cudaStreamSynchronize(a.stream));                   // h_units may be rebuilt by the next call
        st->units_ev_ver = a.ev_ver; st->units_n_it = n_it; st->units_buf_size = a.buf_size;
    }
    if (st->n_units == 0) return PBSO_OK;                             // silence: the mix is already zeroed
    const size_t vneed = (size_t)n_it * npm;
    if (vneed > st->vbase_cap) {
        cudaFree(st->Vbase); st->Vbase = nullptr; st->vbase_cap = 0;
        PBSO_CUDA(cudaMalloc(&st->Vbase, sizeof(float2) * vneed)); st->vbase_cap = vneed;
    }
    k_tc_carrier<<<(unsigned)((npm + 255) / 256), 256, 0, a.stream>>>(a.n_obj, a.n_modes, n_it, tpb, 1, a.lneps, a.theta, a.c3, a.cot,
                                                                   a.d_ev_off, a.d_ev_buf, a.d_ev_space, st->Vbase);
    PBSO_CUDA(cudaGetLastError());
    ++*launches;
    static const int split = getenv("PBSO_TC_SPLIT") ? atoi(getenv("PBSO_TC_SPLIT")) : 1;
    static const int chain = getenv("PBSO_TC_CHAIN") ? atoi(getenv("PBSO_TC_CHAIN")) : 2;
    static const int flush_units = getenv("PBSO_TC_FLUSH") ? atoi(getenv("PBSO_TC_FLUSH")) : TCB_FLUSH_UNITS;
    static const int ablate = getenv("PBSO_TC_ABLATE") ? atoi(getenv("PBSO_TC_ABLATE")) : 0;
    const int grid = st->grid;
#define PBSO_TC_LAUNCH(S, C)                                                                                         \
    do {                                                                                                             \
        static bool attr[64] = {};      /* per device: function attributes do not carry across devices */          \
        int dev_ = 0; cudaGetDevice(&dev_);                                                                          \
        if (!attr[dev_ & 63]) { PBSO_CUDA(cudaFuncSetAttribute(k_batch_tc<S, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM_TS)); attr[dev_ & 63] = true; } \
        k_batch_tc<S, C><<<grid, TCB_THREADS, TCB_SMEM_TS, a.stream>>>(a.n_obj, a.n_modes, n_tiles, st->cta_first, st->units, st->tabA, \
            st->tabB, st->Vbase, a.c3, a.cot, st->ev_row, a.d_ev_space, a.d_mix, flush_units, ablate);                     \
    } while (0)
    if (split == 0 && chain == 2) PBSO_TC_LAUNCH(0, 2);
    else if (split == 2 && chain == 2) PBSO_TC_LAUNCH(2, 2);
    else if (split == 1 && chain == 1) PBSO_TC_LAUNCH(1, 1);
    else if (split == 1 && chain == 4) PBSO_TC_LAUNCH(1, 4);
    else if (split == 2 && chain == 4) PBSO_TC_LAUNCH(2, 4);
    else PBSO_TC_LAUNCH(1, 2);
#undef PBSO_TC_LAUNCH
    PBSO_CUDA(cudaGetLastError());
    ++*launches;
    return PBSO_OK;
}

}  // namespace pbso
