// =============================================================================
// project_tc.cu -- kernel K5: batched impulse projection  Y[b][m] = sum_k U[m][k] F[b][k]
// (tools/real_time_modal_sound.cpp:268-295 for B impulses with dense load vectors) on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA.
//
// Precision: a single TF32 product (10-bit mantissa) cannot hold the 1e-5 column tolerance, so each operand is
// split x = hi + lo with hi = x truncated to TF32 and lo = x - hi (exactly representable), and three MMAs
// accumulate  hi*hi + hi*lo + lo*hi  into the same FP32 TMEM accumulator ("3xTF32"); the dropped lo*lo term is
// O(2^-22).  U is static, so its split is done once at first use (from the FP64 master copy); F is split by a
// small pre-kernel per call.
//
// Tiling: CTA = (128-mode tile, 128-impulse tile, K split).  Per 32-wide K block a stage holds
// A_hi, A_lo [128][32] and B_hi, B_lo [128][32] fp32 = 64 KB, written by four TMA loads with the 128-byte
// swizzle that the UMMA shared-memory descriptors expect (K-major, 8-row groups 1024 B apart); 3 stages.
// Warp 0 lane 0 issues TMA, warp 1 lane 0 issues the 12 MMAs per stage and commits to the stage's "empty"
// mbarrier; four epilogue warps drain TMEM (tcgen05.ld 32x32b.x32) chunk by chunk into FP32 registers (two-level
// accumulation, see k_project_tc) and finally add their [128 x 128] partial tile into Y (FP32 reductions when K
// is split).  Y rows are impulses, so a warp's 32 lanes write 32 consecutive modes: coalesced.
// =============================================================================
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

using namespace pbso;
using namespace pbso::umma;

namespace {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;                 // 16 KB per operand tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;               // A_hi A_lo B_hi B_lo
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_TMEM_COLS = 256;                               // two 128-column accumulators

// Two-level accumulation.  The tensor core adds into its FP32 TMEM accumulator with truncation, so the error of a
// chain grows linearly with its length (measured ~3e-8 per accumulating MMA).  The K loop is therefore cut into
// chunks of TC_CHUNK K blocks (256 K elements = 96 MMAs): each chunk accumulates into one of two TMEM buffers
// from zero, and the four epilogue warps promote finished chunks into FP32 registers (round-to-nearest adds)
// while the MMA warp already works on the other buffer.
constexpr int TC_CHUNK = 8;
constexpr int TC_THREADS = 192;          // warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2-5 epilogue


__global__ void __launch_bounds__(TC_THREADS, 1)
k_project_tc(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
             const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
             float* __restrict__ Y, int M, int B, int kb_total, int kb_per_split, int use_atomics) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // SWIZZLE_128B needs 1024 B alignment
    uint64_t* full = (uint64_t*)(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* acc_full = empty + TC_STAGES;        // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2]
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BN;
    const int kb0 = blockIdx.z * kb_per_split;
    const int kb1 = min(kb0 + kb_per_split, kb_total);
    const int nkb = kb1 - kb0;
    const int nchunks = (nkb + TC_CHUNK - 1) / TC_CHUNK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (i / TC_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * TC_STAGE_BYTES;
                mbar_expect_tx(&full[s], TC_STAGE_BYTES);
                const int k = (kb0 + i) * TC_BK;
                tma_load_2d(st + 0 * TC_TILE_BYTES, &tmAhi, k, m0, &full[s]);
                tma_load_2d(st + 1 * TC_TILE_BYTES, &tmAlo, k, m0, &full[s]);
                tma_load_2d(st + 2 * TC_TILE_BYTES, &tmBhi, k, n0, &full[s]);
                tma_load_2d(st + 3 * TC_TILE_BYTES, &tmBlo, k, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, TC_BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (i / TC_STAGES) & 1;
                const int c = i / TC_CHUNK, buf = c & 1;
                const bool chunk_start = (i % TC_CHUNK) == 0;
                if (chunk_start) {
                    mbar_wait(&acc_empty[buf], (((uint32_t)c >> 1) & 1) ^ 1);      // epilogue has drained this buffer
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(buf * TC_BN);
                const uint32_t a_hi = smem_u32(smem + s * TC_STAGE_BYTES), a_lo = a_hi + TC_TILE_BYTES;
                const uint32_t b_hi = a_hi + 2 * TC_TILE_BYTES, b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {                  // UMMA_K = 8 tf32 = 32 bytes inside the swizzle atom
                    const uint32_t off = k * 32;
                    const uint64_t dAh = umma_desc_k_sw128(a_hi + off), dAl = umma_desc_k_sw128(a_lo + off);
                    const uint64_t dBh = umma_desc_k_sw128(b_hi + off), dBl = umma_desc_k_sw128(b_lo + off);
                    umma_tf32(acc, dAh, dBh, idesc, (chunk_start && k == 0) ? 0u : 1u);
                    umma_tf32(acc, dAh, dBl, idesc, 1u);
                    umma_tf32(acc, dAl, dBh, idesc, 1u);
                }
                umma_commit(&empty[s]);                               // frees the stage when these MMAs retire
                if ((i % TC_CHUNK) == TC_CHUNK - 1 || i == nkb - 1) umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ---------------- epilogue warps: promote chunks into registers, then write Y ----------------
        const int quarter = warp & 3;                               // TMEM lane quarter this warp may access
        const int m = m0 + quarter * 32 + lane;
        float accr[TC_BN];
#pragma unroll
        for (int j = 0; j < TC_BN; ++j) accr[j] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1;
            mbar_wait(&acc_full[buf], ((uint32_t)c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int q = 0; q < TC_BN / 32; ++q) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TC_BN + q * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) accr[q * 32 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[buf]);
        }
        if (m < M && nkb > 0) {
#pragma unroll
            for (int j = 0; j < TC_BN; ++j) {
                const int b = n0 + j;
                if (b < B) {
                    if (use_atomics) atomicAdd(&Y[(size_t)b * M + m], accr[j]);
                    else Y[(size_t)b * M + m] = accr[j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
}

// x (double or float) -> hi = x truncated to TF32 (low 13 mantissa bits cleared), lo = (float)(x - hi)
template <typename TIn>
__global__ void k_split_tf32(const TIn* __restrict__ src, int rows, int cols, int src_pitch,
                             float* __restrict__ hi, float* __restrict__ lo, int dst_pitch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= dst_pitch) return;
    float h = 0.f, l = 0.f;
    if (c < cols && r < rows) {
        const TIn x = src[(size_t)r * src_pitch + c];
        h = __uint_as_float(__float_as_uint((float)x) & 0xFFFFE000u);
        l = (float)(x - (TIn)h);
    }
    hi[(size_t)r * dst_pitch + c] = h;
    lo[(size_t)r * dst_pitch + c] = l;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(CUtensorMap* map, const float* base, int rows, int pitch_elems) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        PBSO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return set_error(PBSO_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)pitch_elems, (cuuint64_t)rows};          // innermost first
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};               // 32 floats = 128 B x 128 rows
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PBSO_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return PBSO_OK;
}

}  // namespace

namespace pbso {

int tc_pitch(int K) { return (K + TC_BK - 1) / TC_BK * TC_BK; }

// Splits a [rows][cols] FP64 matrix into zero-padded TF32 hi / lo planes with pitch tc_pitch(cols).
int tc_split_f64(const double* d_src, int rows, int cols, float* d_hi, float* d_lo, cudaStream_t s) {
    const int pitch = tc_pitch(cols);
    k_split_tf32<double><<<dim3(div_up(pitch, 256), rows), 256, 0, s>>>(d_src, rows, cols, cols, d_hi, d_lo, pitch);
    PBSO_CUDA(cudaGetLastError());
    return PBSO_OK;
}
int tc_split_f32(const float* d_src, int rows, int cols, float* d_hi, float* d_lo, cudaStream_t s) {
    const int pitch = tc_pitch(cols);
    k_split_tf32<float><<<dim3(div_up(pitch, 256), rows), 256, 0, s>>>(d_src, rows, cols, cols, d_hi, d_lo, pitch);
    PBSO_CUDA(cudaGetLastError());
    return PBSO_OK;
}

// Y[B][M] (float, overwritten) = U[M][K] F[B][K]^T from pre-split planes (pitch tc_pitch(K)).
int tc_project(const float* d_Uhi, const float* d_Ulo, int M, const float* d_Fhi, const float* d_Flo, int B, int K,
               float* d_Y, int sm_count, cudaStream_t s) {
    const int pitch = tc_pitch(K);
    CUtensorMap mAh, mAl, mBh, mBl;
    if (int rc = encode_map(&mAh, d_Uhi, M, pitch)) return rc;
    if (int rc = encode_map(&mAl, d_Ulo, M, pitch)) return rc;
    if (int rc = encode_map(&mBh, d_Fhi, B, pitch)) return rc;
    if (int rc = encode_map(&mBl, d_Flo, B, pitch)) return rc;
    static bool attr_set[64] = {};        // per device: function attributes do not carry across devices
    int dev_ = 0; cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        PBSO_CUDA(cudaFuncSetAttribute(k_project_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set[dev_ & 63] = true;
    }
    const int mt = div_up(M, TC_BM), nt = div_up(B, TC_BN), kb_total = pitch / TC_BK;
    // split K: smallest split count whose grid fills whole waves of SMs to >= 90 % (1 CTA per SM), every split
    // keeping at least two chunks of K
    const int tiles = mt * nt;
    const int max_splits = std::max(1, kb_total / (2 * TC_CHUNK));
    int splits = 1; double best = 0.0;
    for (int sp = 1; sp <= std::min(max_splits, 64); ++sp) {
        const double waves = (double)tiles * sp / sm_count;
        const double eff = waves / std::ceil(waves);
        if (eff > best + 1e-9) { best = eff; splits = sp; }
        if (eff >= 0.9 && waves >= 1.0) { splits = sp; break; }
    }
    const int kb_per = div_up(kb_total, splits);
    splits = div_up(kb_total, kb_per);
    if (splits > 1) PBSO_CUDA(cudaMemsetAsync(d_Y, 0, sizeof(float) * (size_t)B * M, s));
    k_project_tc<<<dim3(mt, nt, splits), TC_THREADS, TC_SMEM_BYTES, s>>>(mAh, mAl, mBh, mBl, d_Y, M, B, kb_total, kb_per,
                                                                      splits > 1 ? 1 : 0);
    PBSO_CUDA(cudaGetLastError());
    return PBSO_OK;
}

}  // namespace pbso
