// =============================================================================
// tc_probe.cu -- tcgen05 measurement helpers used by bench.py and the tests.
//
//  pbso_measure_tc_peak  bare tcgen05.mma loop (A operand in TMEM, B in shared memory, K-major 128-byte swizzle):
//                        one persistent CTA (cta_group::1) or CTA pair (cta_group::2) per SM issues MMAs back to
//                        back with nothing else running; optionally four more warps hammer shared memory with
//                        conflict-free stores at the same time (how much of the shared-memory pipe the tensor
//                        core's own B fetches leave free).  This is the denominator of the tensor roofline of
//                        k_batch_tc (batch_tc.cu) and k_project_tc (project_tc.cu): MEASURED_PEAKS.json holds a
//                        cuBLAS bf16 figure only.
//  pbso_tc_selftest      one [256 x 128] x K product on a CTA pair with integer-valued operands (exact in TF32
//                        and FP16): pins the operand placement the pair kernels rely on -- A rows in each CTA's
//                        TMEM lanes, B rows split between the two CTAs' shared memory, D rows in each CTA's TMEM.
// =============================================================================
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>
#include <algorithm>
#include <vector>

using namespace pbso;
using namespace pbso::umma;

namespace {

// byte offset of (row r, 16-byte chunk c) inside a K-major SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128(int r, int c) {
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}

template <int KIND, int CG, int N>
__global__ void __launch_bounds__(256, 1)
k_tc_peak(int iters, int stress, unsigned long long* __restrict__ cyc, unsigned long long* __restrict__ stress_ops) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int ROWS_B = N / CG;                                  // B rows held by this CTA
    uint8_t* scratch = smem + ROWS_B * 128;                         // 16 KB for the stress warps
    uint64_t* done = (uint64_t*)(scratch + 16384);
    uint32_t* tmem_slot = (uint32_t*)(done + 1);
    volatile uint32_t* stop = (volatile uint32_t*)(tmem_slot + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        mbar_init(done, 1);
        *stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    // benign operand values: B tile in shared memory, A in TMEM columns 256..287
    for (int i = threadIdx.x; i < ROWS_B * 32; i += blockDim.x) {
        const float v = 0.25f + 0.001f * (float)(i & 63);
        if (KIND == 0) ((float*)smem)[i] = v;
        else ((__half2*)smem)[i] = __floats2half2_rn(v, -v);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float f = 0.5f + 0.01f * (float)((lane + j) & 15);
            if (KIND == 0) v[j] = __float_as_uint(f);
            else { __half2 h = __floats2half2_rn(f, -f); v[j] = *reinterpret_cast<uint32_t*>(&h); }
        }
        tmem_st_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 256, v);
        tmem_st_wait();
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();

    if (warp == 0) {
        if (rank == 0) {
            constexpr uint32_t idesc = KIND == 0 ? umma_idesc_tf32(128 * CG, N) : umma_idesc_f16(128 * CG, N);
            const uint64_t dB = umma_desc_k_sw128(smem_u32(smem));
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_ts<KIND, CG>(tmem_base, tmem_base + 256 + 8 * k, dB + 2 * k, idesc, (it | k) ? 1u : 0u);
                }
                __syncwarp();
            }
            if (elect_one()) {
                if (CG == 1) umma_commit(done); else umma_commit_2cta(done, 3);
            }
            __syncwarp();
            mbar_wait(done, 0);
            const long long t1 = clock64();
            if (lane == 0) cyc[blockIdx.x / CG] = (unsigned long long)(t1 - t0);
        } else {
            mbar_wait(done, 0);
        }
        *stop = 1;
    } else if (warp >= 4 && stress) {
        // conflict-free 8-byte stores: a warp writes 256 consecutive bytes = 2 wavefronts per instruction
        const uint32_t base = smem_u32(scratch) + (uint32_t)((warp - 4) * 4096) + (uint32_t)lane * 8u;
        unsigned long long n = 0;
        while (*stop == 0) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + (uint32_t)r * 256u), "r"(lane), "r"(r) : "memory");
            n += 16;
        }
        if (lane == 0) atomicAdd(stress_ops, n);
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// D[256][128] = A[256][K] B[128][K]^T on one CTA pair; K = 32 (tf32) or 64 (f16): one 128-byte swizzle row, 4 MMAs.
template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    constexpr int K = KIND == 0 ? 32 : 64;
    __shared__ __align__(1024) uint8_t sB[64 * 128];
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // A: thread = row of this CTA's half = TMEM lane
    {
        const float* a = A + (size_t)(rank * 128 + threadIdx.x) * K;
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (KIND == 0) v[j] = __float_as_uint(a[j]);
            else { __half2 h = __floats2half2_rn(a[2 * j], a[2 * j + 1]); v[j] = *reinterpret_cast<uint32_t*>(&h); }
        }
        tmem_st_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 128, v);
        tmem_st_wait();
    }
    // B: this CTA holds rows [64 rank, 64 rank + 64) of the 128
    if (threadIdx.x < 64) {
        const int r = threadIdx.x;
        const float* b = B + (size_t)(rank * 64 + r) * K;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (KIND == 0) w[q] = __float_as_uint(b[4 * c + q]);
                else { __half2 h = __floats2half2_rn(b[8 * c + 2 * q], b[8 * c + 2 * q + 1]); w[q] = *reinterpret_cast<uint32_t*>(&h); }
            }
            *reinterpret_cast<uint4*>(sB + sw128(r, c)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            constexpr uint32_t idesc = KIND == 0 ? umma_idesc_tf32(256, 128) : umma_idesc_f16(256, 128);
            const uint64_t dB = umma_desc_k_sw128(smem_u32(sB));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ts<KIND, 2>(tmem_base, tmem_base + 128 + 8 * k, dB + 2 * k, idesc, k ? 1u : 0u);
            umma_commit_2cta(&done, 3);
        }
        __syncwarp();
    }
    mbar_wait(&done, 0);
    tcgen05_fence_after();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(32 * q), v);
        tmem_ld_wait();
        float* d = D + (size_t)(rank * 128 + threadIdx.x) * 128 + 32 * q;
#pragma unroll
        for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(v[j]);
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

template <int KIND, int CG, int N>
int run_peak(int iters, int stress, int sms, double* tflops, double* cycles_per_mma, double* stress_wf) {
    const int smem = (N / CG) * 128 + 16384 + 64 + 1024;
    PBSO_CUDA(cudaFuncSetAttribute(k_tc_peak<KIND, CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = (sms / CG) * CG;
    unsigned long long *d_cyc, *d_ops;
    PBSO_CUDA(cudaMalloc(&d_cyc, sizeof(unsigned long long) * grid));
    PBSO_CUDA(cudaMalloc(&d_ops, sizeof(unsigned long long)));
    cudaEvent_t e0, e1; PBSO_CUDA(cudaEventCreate(&e0)); PBSO_CUDA(cudaEventCreate(&e1));
    float best = 1e30f; std::vector<unsigned long long> cyc(grid / CG); unsigned long long ops = 0;
    for (int rep = 0; rep < 4; ++rep) {
        PBSO_CUDA(cudaMemset(d_ops, 0, sizeof(unsigned long long)));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        PBSO_CUDA(cudaEventRecord(e0));
        PBSO_CUDA(cudaLaunchKernelEx(&cfg, k_tc_peak<KIND, CG, N>, iters, stress, d_cyc, d_ops));
        PBSO_CUDA(cudaEventRecord(e1));
        PBSO_CUDA(cudaEventSynchronize(e1));
        float ms; PBSO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) {
            best = ms;
            PBSO_CUDA(cudaMemcpy(cyc.data(), d_cyc, sizeof(unsigned long long) * (grid / CG), cudaMemcpyDeviceToHost));
            PBSO_CUDA(cudaMemcpy(&ops, d_ops, sizeof(ops), cudaMemcpyDeviceToHost));
        }
    }
    PBSO_CUDA(cudaGetLastError());
    std::sort(cyc.begin(), cyc.end());
    const double med = (double)cyc[cyc.size() / 2];
    const double kk = KIND == 0 ? 8.0 : 16.0;
    const double flops = 2.0 * 128.0 * N * kk * 4.0 * iters * grid;          // every CTA computes 128 x N x K per MMA
    *tflops = flops / (best * 1e-3) / 1e12;
    if (cycles_per_mma) *cycles_per_mma = med / (4.0 * iters);
    if (stress_wf) *stress_wf = stress ? (double)ops * 2.0 / (double)grid / med : 0.0;   // wavefronts per cycle per SM
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_cyc); cudaFree(d_ops);
    return PBSO_OK;
}

}  // namespace

extern "C" {

int pbso_measure_tc_peak(int kind, int cta_group, int n, int stress, double* tflops, double* cycles_per_mma,
                         double* stress_wavefronts_per_cycle) {
    if (int rc = check_device()) return rc;
    PBSO_REQUIRE(tflops, PBSO_ERR_INVALID, "null output");
    PBSO_REQUIRE((kind == 0 || kind == 1) && (cta_group == 1 || cta_group == 2) && (n == 128 || n == 256), PBSO_ERR_INVALID,
                 "kind in {0: tf32, 1: f16}, cta_group in {1,2}, n in {128,256}");
    int dev; PBSO_CUDA(cudaGetDevice(&dev));
    int sms; PBSO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int iters = 20000;
#define PBSO_PEAK(K_, C_, N_) if (kind == K_ && cta_group == C_ && n == N_) return run_peak<K_, C_, N_>(iters, stress, sms, tflops, cycles_per_mma, stress_wavefronts_per_cycle)
    PBSO_PEAK(0, 1, 128); PBSO_PEAK(0, 1, 256); PBSO_PEAK(0, 2, 128); PBSO_PEAK(0, 2, 256);
    PBSO_PEAK(1, 1, 128); PBSO_PEAK(1, 1, 256); PBSO_PEAK(1, 2, 128); PBSO_PEAK(1, 2, 256);
#undef PBSO_PEAK
    return PBSO_ERR_INVALID;
}

// The same loop launched back to back for at least min_ms of device time: the tensor pipe's SUSTAINED rate under the
// board's power cap, the denominator for a kernel that is timed inside a long step (the burst figure above is for a
// kernel timed alone).
int pbso_measure_tc_peak_sustained(int kind, int cta_group, int n, double min_ms, double* tflops) {
    if (int rc = check_device()) return rc;
    PBSO_REQUIRE(tflops && min_ms > 0, PBSO_ERR_INVALID, "bad argument");
    double burst = 0.0, cyc = 0.0;
    if (int rc = pbso_measure_tc_peak(kind, cta_group, n, 0, &burst, &cyc, nullptr)) return rc;     // also validates the arguments
    int dev; PBSO_CUDA(cudaGetDevice(&dev));
    int sms; PBSO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int iters = 20000, grid = (sms / cta_group) * cta_group;
    const double flop_per_launch = 2.0 * 128.0 * n * (kind == 0 ? 8.0 : 16.0) * 4.0 * iters * grid;
    const int launches = std::max(4, (int)(min_ms / (flop_per_launch / (burst * 1e12) * 1e3)) + 1);
    cudaEvent_t e0, e1; PBSO_CUDA(cudaEventCreate(&e0)); PBSO_CUDA(cudaEventCreate(&e1));
    unsigned long long *d_cyc, *d_ops;
    PBSO_CUDA(cudaMalloc(&d_cyc, sizeof(unsigned long long) * grid));
    PBSO_CUDA(cudaMalloc(&d_ops, sizeof(unsigned long long)));
    const int smem = (n / cta_group) * 128 + 16384 + 64 + 1024;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cta_group; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    PBSO_CUDA(cudaEventRecord(e0));
    for (int l = 0; l < launches; ++l) {
#define PBSO_SUS(K_, C_, N_) if (kind == K_ && cta_group == C_ && n == N_) PBSO_CUDA(cudaLaunchKernelEx(&cfg, k_tc_peak<K_, C_, N_>, iters, 0, d_cyc, d_ops))
        PBSO_SUS(0, 1, 128); PBSO_SUS(0, 1, 256); PBSO_SUS(0, 2, 128); PBSO_SUS(0, 2, 256);
        PBSO_SUS(1, 1, 128); PBSO_SUS(1, 1, 256); PBSO_SUS(1, 2, 128); PBSO_SUS(1, 2, 256);
#undef PBSO_SUS
    }
    PBSO_CUDA(cudaEventRecord(e1));
    PBSO_CUDA(cudaEventSynchronize(e1));
    float ms; PBSO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *tflops = flop_per_launch * launches / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_cyc); cudaFree(d_ops);
    return PBSO_OK;
}

int pbso_tc_selftest(int kind, double* max_err) {
    if (int rc = check_device()) return rc;
    PBSO_REQUIRE(max_err && (kind == 0 || kind == 1), PBSO_ERR_INVALID, "bad argument");
    const int K = kind == 0 ? 32 : 64;
    std::vector<float> A(256 * K), B(128 * K), D(256 * 128), R(256 * 128, 0.f);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 24) % 9) - 4); };
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    for (int i = 0; i < 256; ++i)
        for (int j = 0; j < 128; ++j) {
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc += A[i * K + k] * B[j * K + k];
            R[i * 128 + j] = acc;
        }
    float *dA, *dB, *dD;
    PBSO_CUDA(cudaMalloc(&dA, A.size() * 4)); PBSO_CUDA(cudaMalloc(&dB, B.size() * 4)); PBSO_CUDA(cudaMalloc(&dD, D.size() * 4));
    PBSO_CUDA(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    PBSO_CUDA(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    PBSO_CUDA(cudaMemset(dD, 0xff, D.size() * 4));
    if (kind == 0) k_tc_selftest<0><<<2, 128>>>(dA, dB, dD); else k_tc_selftest<1><<<2, 128>>>(dA, dB, dD);
    PBSO_CUDA(cudaGetLastError());
    PBSO_CUDA(cudaDeviceSynchronize());
    PBSO_CUDA(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    double e = 0.0;
    for (size_t i = 0; i < D.size(); ++i) { const double d = std::fabs((double)D[i] - (double)R[i]); if (!(d <= e)) e = d; }
    *max_err = e;
    return PBSO_OK;
}

}  // extern "C"
