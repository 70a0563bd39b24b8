// =============================================================================
// integrator.cu -- ModalIntegrator<double> on the device + the per-buffer synthesis kernel.
//
//  K2  k_coeffs        modal_integrator.h:86-100   (a,b) -> (c1,c2,c3), FP64, once per object
//      k_build_ab      modal_integrator.h:62-67    (rho, w^2, alpha, beta) -> (a,b)
//      k_step          modal_integrator.h:103-123  single Step(Q) / Step()
//  K1  k_render_f64    modal_solver.h:261-272      BUF_SIZE steps + transfer-weighted modal sum
//                                                  + qnorm, state held in registers
//
// HBM layout (structure of arrays, one double array per quantity, length N):
//   c1 | c2 | c3 | q1 (= q_{k-1}) | q2 (= q_{k-2}) | transfer[L][n_transfer]
// One thread owns one mode for the whole buffer: 5 coalesced loads, T recurrence steps in
// registers, 2 coalesced stores.  The cross-mode sum y[i] = sum_m T_m q_m[i] is done 32 samples at
// a time with a register-transposing butterfly (31 shuffles per 32 outputs) so that no per-sample
// warp reduction is needed.
// =============================================================================
#include "common.cuh"
#include <cstring>
#include <utility>
#include <vector>

using namespace pbso;

struct pbso_integrator {
    int device = 0;
    int N = 0;
    double h = 0;
    cudaStream_t stream = nullptr;
    double* d_c = nullptr;        // c1 | c2 | c3
    double* d_q = nullptr;        // q1 | q2   (current state)
    double* d_q_alt = nullptr;    // ping-pong partner: K1 reads d_q and writes d_q_alt, then they swap
    double* d_in = nullptr;       // staging: space[N] | time[Tcap]
    double* d_out = nullptr;      // staging: y[L*Tcap] | qnorm[N]
    double* d_trans = nullptr;    // transfer[L][n_transfer]
    double* d_pos = nullptr;      // listener positions for set_transfer_ffat
    int pos_cap = 0;
    double* d_acc = nullptr;      // [L][T] cross-slab accumulators of K1 (all zero between launches)
    unsigned* d_cnt = nullptr;    // per listener group: slabs that have finished (zero between launches)
    size_t acc_cap = 0; int cnt_cap = 0;
    int n_transfer = 0, L = 0, Tcap = 0, Lcap = 0;
    double* h_in = nullptr;       // pinned mirrors of d_in / d_out
    double* h_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
};

// ---------------------------------------------------------------------------------------------
__global__ void k_build_ab(int N, double density, const double* __restrict__ omega2, double alpha,
                           double beta, double* __restrict__ a, double* __restrict__ b) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double omega = sqrt(omega2[i] / density);            // modal_integrator.h:63
    double xi = 0.5 * (alpha / omega + beta * omega);    // :64
    a[i] = 2.0 * xi * omega;                             // :65
    b[i] = omega * omega;                                // :66  pow(omega, 2)
}

__global__ void k_coeffs(int N, double h, const double* __restrict__ a, const double* __restrict__ b,
                         double* __restrict__ c1, double* __restrict__ c2, double* __restrict__ c3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double ai = a[i], bi = b[i];
    const double epsilon = exp(-ai / 2 * h);                       // modal_integrator.h:89
    const double theta = h * sqrt(bi - ai * ai / 4.0);             // :90
    const double gamma = asin(ai / (2.0 * sqrt(bi)));              // :91
    const double omega = sqrt(bi);                                 // :92
    const double omega_d = sqrt(bi - ai * ai / 4.0);               // :93
    c1[i] = 2.0 * epsilon * cos(theta);                            // :95
    c2[i] = -(epsilon * epsilon);                                  // :96
    double v = 2.0 * (epsilon * cos(theta + gamma) - epsilon * epsilon * cos(2.0 * theta + gamma));
    v /= (3.0 * omega * omega_d);                                  // :98
    c3[i] = v * 1E9;                                               // :99
}

__global__ void k_step(int N, const double* __restrict__ c, double* __restrict__ q,
                       const double* __restrict__ Q, double* __restrict__ q_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double q1 = q[i], q2 = q[N + i];
    double qk = c[i] * q1 + c[N + i] * q2;                          // modal_integrator.h:120
    if (Q) qk += c[2 * N + i] * Q[i];                               // :109-110
    q[i] = qk; q[N + i] = q1;                                       // ring rotate (:111)
    q_out[i] = qk;
}

// ---------------------------------------------------------------------------------------------
// K1: block = 256 modes; blockIdx.y = listener group (LPB listeners).  Blocks with blockIdx.y > 0
// recompute the recurrence (cheap) but only block row 0 writes state / qnorm.
// Measured for cfg4 (1024 modes, 64 listeners, 256 samples): 25.7 us.  A two-phase variant (one shared recurrence pass
// writing q[t][m], then the (L x M).(M x T) modal sum as an FP64 tile contraction) was tried and dropped: the recurrence
// alone is 12 us of FP64 dependency latency (256 steps x 3 dependent ops), and the contraction on FP64 FMA units was
// shared-memory-bandwidth bound at 26 us, so fusing the sum into the recurrence's shadow is the faster shape.
// ---------------------------------------------------------------------------------------------
constexpr int K1_TPB = 256;
constexpr int K1_TILE = 32;

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

// v[0..31] per lane -> returns sum over lanes of v[lane]
__device__ __forceinline__ double transpose_reduce32(double (&v)[K1_TILE], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const double keep = up ? v[i + s] : v[i];
            const double send = up ? v[i] : v[i + s];
            v[i] = keep + shfl_xor_f64(send, s);
        }
    }
    return v[0];
}

template <int LPB>
__global__ void __launch_bounds__(K1_TPB)
k_render_f64(int N, int T, int n_transfer, int L,
             const double* __restrict__ c, const double* __restrict__ q, double* __restrict__ q_next,
             const double* __restrict__ space, const double* __restrict__ time,
             const double* __restrict__ trans, double* __restrict__ y, double* __restrict__ qnorm,
             double* __restrict__ acc, unsigned* __restrict__ cnt) {
    extern __shared__ double smem[];
    double* s_time = smem;                                  // [T]
    double* s_part = smem + T;                              // [LPB][K1_TPB/32][32] per tile
    __shared__ bool s_last;
    const int m = blockIdx.x * K1_TPB + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l0 = blockIdx.y * LPB;
    for (int i = threadIdx.x; i < T; i += K1_TPB) s_time[i] = time[i];
    __syncthreads();

    const bool live = m < N;
    double c1 = 0, c2 = 0, c3s = 0, sp = 0, q1 = 0, q2 = 0, qsum = 0;
    if (live) {
        c1 = c[m]; c2 = c[N + m]; c3s = c[2 * N + m]; sp = space[m];
        q1 = q[m]; q2 = q[N + m];
    }
    double tl[LPB];
#pragma unroll
    for (int l = 0; l < LPB; ++l)
        tl[l] = (live && m < n_transfer && l0 + l < L) ? trans[(size_t)(l0 + l) * n_transfer + m] : 0.0;

    for (int t0 = 0; t0 < T; t0 += K1_TILE) {
        double qt[K1_TILE];
#pragma unroll
        for (int j = 0; j < K1_TILE; ++j) {
            const int i = t0 + j;
            double qk = 0.0;
            if (i < T) {
                const double Q = sp * s_time[i];            // modal_solver.h:266  space * time(ii)
                // modal_integrator.h:109-110  q_k = c1 q_{k-1} + c2 q_{k-2} + c3 Q.  Associated as c1 q_{k-1} + (c2 q_{k-2}
                // + c3 Q): the bracket does not depend on q_{k-1}, so the loop-carried dependency is ONE FP64 FMA per
                // sample instead of three (FP64 results return after ~29 cycles on this part: 12 us -> 4 us per 256-sample
                // buffer).  Same terms, one rounding placed differently; parity with the oracle stays ~1e-11 of full scale.
                const double older = c2 * q2 + c3s * Q;
                qk = c1 * q1 + older;
                q2 = q1; q1 = qk;
                qsum += qk * qk;                            // modal_solver.h:270
            }
            qt[j] = qk;
        }
        if (L > 0) {
#pragma unroll
            for (int l = 0; l < LPB; ++l) {
                double v[K1_TILE];
#pragma unroll
                for (int j = 0; j < K1_TILE; ++j) v[j] = tl[l] * qt[j];   // q.head(n).dot(transfer)
                const double r = transpose_reduce32(v, lane);
                s_part[(l * (K1_TPB / 32) + warp) * 32 + lane] = r;
            }
            __syncthreads();
            // K1_TPB/32 = 8 warps -> 8 partials per (listener, sample)
            for (int o = threadIdx.x; o < LPB * 32; o += K1_TPB) {
                const int l = o >> 5, j = o & 31;
                if (l0 + l < L && t0 + j < T) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < K1_TPB / 32; ++w) s += s_part[(l * (K1_TPB / 32) + w) * 32 + j];
                    if (gridDim.x == 1) y[(size_t)(l0 + l) * T + t0 + j] = s;
                    else atomicAdd(&acc[(size_t)(l0 + l) * T + t0 + j], s);
                }
            }
            __syncthreads();
        }
    }
    if (live && blockIdx.y == 0) {
        q_next[m] = q1; q_next[N + m] = q2;
        if (qnorm) qnorm[m] = sqrt(qsum);                    // modal_solver.h:272
    }
    // Several mode slabs: the last slab to finish a listener group moves the sums to y (which may be host memory
    // mapped into the device: no copy-engine launch on the way back) and leaves the accumulators zero for the next call.
    if (gridDim.x > 1 && L > 0) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&cnt[blockIdx.y], 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int o = threadIdx.x; o < LPB * T; o += K1_TPB) {
                const int l = l0 + o / T, i = o - (o / T) * T;
                if (l < L) {
                    y[(size_t)l * T + i] = __ldcg(&acc[(size_t)l * T + i]);
                    acc[(size_t)l * T + i] = 0.0;
                }
            }
            if (threadIdx.x == 0) cnt[blockIdx.y] = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
static int ensure_staging(pbso_integrator* it, int T, int L) {
    size_t in_need = (size_t)it->N + T, out_need = (size_t)(L > 0 ? L : 1) * T + it->N;
    if (in_need > it->in_cap) {
        if (it->d_in) cudaFree(it->d_in);
        if (it->h_in) cudaFreeHost(it->h_in);
        PBSO_CUDA(cudaMalloc(&it->d_in, in_need * sizeof(double)));
        PBSO_CUDA(cudaMallocHost(&it->h_in, in_need * sizeof(double)));
        it->in_cap = in_need;
    }
    if (out_need > it->out_cap) {
        if (it->d_out) cudaFree(it->d_out);
        if (it->h_out) cudaFreeHost(it->h_out);
        PBSO_CUDA(cudaMalloc(&it->d_out, out_need * sizeof(double)));
        PBSO_CUDA(cudaMallocHost(&it->h_out, out_need * sizeof(double)));
        it->out_cap = out_need;
    }
    return PBSO_OK;
}

static int launch_render(pbso_integrator* it, const double* d_space, const double* d_time, int T,
                         double* d_y, double* d_qnorm) {
    const int N = it->N, L = it->L;
    const int gx = div_up(N, K1_TPB);
    if (L > 0 && gx > 1) {
        const size_t need = (size_t)L * T; const int groups = L <= 1 ? 1 : div_up(L, 4);
        if (need > it->acc_cap || groups > it->cnt_cap) {
            PBSO_CUDA(cudaStreamSynchronize(it->stream));
            cudaFree(it->d_acc); cudaFree(it->d_cnt); it->d_acc = nullptr; it->d_cnt = nullptr; it->acc_cap = 0; it->cnt_cap = 0;
            PBSO_CUDA(cudaMalloc(&it->d_acc, need * sizeof(double)));
            PBSO_CUDA(cudaMalloc(&it->d_cnt, groups * sizeof(unsigned)));
            PBSO_CUDA(cudaMemsetAsync(it->d_acc, 0, need * sizeof(double), it->stream));   // once: every launch leaves them zero
            PBSO_CUDA(cudaMemsetAsync(it->d_cnt, 0, groups * sizeof(unsigned), it->stream));
            it->acc_cap = need; it->cnt_cap = groups;
        }
    }
    if (L <= 1) {
        size_t sm = sizeof(double) * ((size_t)T + 1 * (K1_TPB / 32) * 32);
        k_render_f64<1><<<dim3(gx, 1), K1_TPB, sm, it->stream>>>(
            N, T, it->n_transfer, L, it->d_c, it->d_q, it->d_q_alt, d_space, d_time, it->d_trans, d_y, d_qnorm, it->d_acc, it->d_cnt);
    } else {
        constexpr int LPB = 4;
        size_t sm = sizeof(double) * ((size_t)T + LPB * (K1_TPB / 32) * 32);
        k_render_f64<LPB><<<dim3(gx, div_up(L, LPB)), K1_TPB, sm, it->stream>>>(
            N, T, it->n_transfer, L, it->d_c, it->d_q, it->d_q_alt, d_space, d_time, it->d_trans, d_y, d_qnorm, it->d_acc, it->d_cnt);
    }
    PBSO_CUDA(cudaGetLastError());
    std::swap(it->d_q, it->d_q_alt);
    return PBSO_OK;
}

extern "C" {

int pbso_integrator_create(int N, double h, const double* a, const double* b, pbso_integrator** out) {
    PBSO_REQUIRE(out, PBSO_ERR_INVALID, "null output handle");
    *out = nullptr;
    PBSO_REQUIRE(N > 0 && a && b, PBSO_ERR_INVALID, "N must be > 0 and a, b non-null");
    if (int rc = check_device()) return rc;
    pbso_integrator* it = new pbso_integrator();
    it->N = N; it->h = h;
    PBSO_CUDA(cudaGetDevice(&it->device));
    PBSO_CUDA(cudaStreamCreateWithFlags(&it->stream, cudaStreamNonBlocking));
    PBSO_CUDA(cudaMalloc(&it->d_c, sizeof(double) * 3 * N));
    PBSO_CUDA(cudaMalloc(&it->d_q, sizeof(double) * 2 * N));
    PBSO_CUDA(cudaMalloc(&it->d_q_alt, sizeof(double) * 2 * N));
    PBSO_CUDA(cudaMemsetAsync(it->d_q, 0, sizeof(double) * 2 * N, it->stream));   // :76-78
    double* d_ab; PBSO_CUDA(cudaMalloc(&d_ab, sizeof(double) * 2 * N));
    PBSO_CUDA(cudaMemcpyAsync(d_ab, a, sizeof(double) * N, cudaMemcpyHostToDevice, it->stream));
    PBSO_CUDA(cudaMemcpyAsync(d_ab + N, b, sizeof(double) * N, cudaMemcpyHostToDevice, it->stream));
    k_coeffs<<<div_up(N, 256), 256, 0, it->stream>>>(N, h, d_ab, d_ab + N, it->d_c, it->d_c + N,
                                                     it->d_c + 2 * N);
    PBSO_CUDA(cudaGetLastError());
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    cudaFree(d_ab);
    // default transfer = TransMessage::setToUnit (modal_solver.h:89-92): ones * 1e7, one listener
    std::vector<double> unit(N, 1E7);
    *out = it;
    return pbso_integrator_set_transfer(it, unit.data(), N, 1);
}

int pbso_integrator_build(double density, const double* omega_squared, int n_omega, double alpha,
                          double beta, double h, int N, pbso_integrator** out) {
    PBSO_REQUIRE(out, PBSO_ERR_INVALID, "null output handle");
    *out = nullptr;
    PBSO_REQUIRE(omega_squared && n_omega > 0, PBSO_ERR_INVALID, "omega_squared empty");
    if (N < 0) N = n_omega;                                           // modal_integrator.h:53-54
    PBSO_REQUIRE(N <= n_omega && N > 0, PBSO_ERR_INVALID, "N for modal integrator invalid");  // :56
    if (int rc = check_device()) return rc;
    double *d_w, *d_ab;
    PBSO_CUDA(cudaMalloc(&d_w, sizeof(double) * N));
    PBSO_CUDA(cudaMalloc(&d_ab, sizeof(double) * 2 * N));
    PBSO_CUDA(cudaMemcpy(d_w, omega_squared, sizeof(double) * N, cudaMemcpyHostToDevice));
    k_build_ab<<<div_up(N, 256), 256>>>(N, density, d_w, alpha, beta, d_ab, d_ab + N);
    PBSO_CUDA(cudaGetLastError());
    std::vector<double> ab(2 * (size_t)N);
    PBSO_CUDA(cudaMemcpy(ab.data(), d_ab, sizeof(double) * 2 * N, cudaMemcpyDeviceToHost));
    cudaFree(d_w); cudaFree(d_ab);
    return pbso_integrator_create(N, h, ab.data(), ab.data() + N, out);
}

int pbso_integrator_destroy(pbso_integrator* it) {
    if (!it) return PBSO_OK;
    DeviceGuard g(it->device);
    if (it->stream) cudaStreamSynchronize(it->stream);
    cudaFree(it->d_c); cudaFree(it->d_q); cudaFree(it->d_q_alt); cudaFree(it->d_in); cudaFree(it->d_out);
    cudaFree(it->d_trans); cudaFree(it->d_pos); cudaFree(it->d_acc); cudaFree(it->d_cnt);
    if (it->h_in) cudaFreeHost(it->h_in);
    if (it->h_out) cudaFreeHost(it->h_out);
    if (it->stream) cudaStreamDestroy(it->stream);
    delete it;
    return PBSO_OK;
}

int pbso_integrator_stream(const pbso_integrator* it, void** cuda_stream) {
    PBSO_REQUIRE(it && cuda_stream, PBSO_ERR_INVALID, "null argument");
    *cuda_stream = (void*)it->stream;
    return PBSO_OK;
}

int pbso_integrator_listeners(const pbso_integrator* it, int* L) {
    PBSO_REQUIRE(it && L, PBSO_ERR_INVALID, "null argument");
    *L = it->L;
    return PBSO_OK;
}

int pbso_integrator_size(const pbso_integrator* it, int* N) {
    PBSO_REQUIRE(it && N, PBSO_ERR_INVALID, "null argument");
    *N = it->N;
    return PBSO_OK;
}

int pbso_integrator_get_coeffs(const pbso_integrator* it, double* c1, double* c2, double* c3) {
    PBSO_REQUIRE(it, PBSO_ERR_INVALID, "null handle");
    DeviceGuard g(it->device);
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    const size_t nb = sizeof(double) * it->N;
    if (c1) PBSO_CUDA(cudaMemcpy(c1, it->d_c, nb, cudaMemcpyDeviceToHost));
    if (c2) PBSO_CUDA(cudaMemcpy(c2, it->d_c + it->N, nb, cudaMemcpyDeviceToHost));
    if (c3) PBSO_CUDA(cudaMemcpy(c3, it->d_c + 2 * it->N, nb, cudaMemcpyDeviceToHost));
    return PBSO_OK;
}

int pbso_integrator_step(pbso_integrator* it, const double* Q, double* q_out) {
    PBSO_REQUIRE(it && q_out, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(it->device);
    if (int rc = ensure_staging(it, 1, 1)) return rc;
    const int N = it->N;
    if (Q) {
        std::memcpy(it->h_in, Q, sizeof(double) * N);
        PBSO_CUDA(cudaMemcpyAsync(it->d_in, it->h_in, sizeof(double) * N, cudaMemcpyHostToDevice, it->stream));
    }
    k_step<<<div_up(N, 256), 256, 0, it->stream>>>(N, it->d_c, it->d_q, Q ? it->d_in : nullptr, it->d_out);
    PBSO_CUDA(cudaGetLastError());
    PBSO_CUDA(cudaMemcpyAsync(it->h_out, it->d_out, sizeof(double) * N, cudaMemcpyDeviceToHost, it->stream));
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    std::memcpy(q_out, it->h_out, sizeof(double) * N);
    return PBSO_OK;
}

int pbso_integrator_get_state(const pbso_integrator* it, double* q_km1, double* q_km2) {
    PBSO_REQUIRE(it && q_km1 && q_km2, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(it->device);
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    PBSO_CUDA(cudaMemcpy(q_km1, it->d_q, sizeof(double) * it->N, cudaMemcpyDeviceToHost));
    PBSO_CUDA(cudaMemcpy(q_km2, it->d_q + it->N, sizeof(double) * it->N, cudaMemcpyDeviceToHost));
    return PBSO_OK;
}

int pbso_integrator_set_state(pbso_integrator* it, const double* q_km1, const double* q_km2) {
    PBSO_REQUIRE(it && q_km1 && q_km2, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(it->device);
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    PBSO_CUDA(cudaMemcpy(it->d_q, q_km1, sizeof(double) * it->N, cudaMemcpyHostToDevice));
    PBSO_CUDA(cudaMemcpy(it->d_q + it->N, q_km2, sizeof(double) * it->N, cudaMemcpyHostToDevice));
    return PBSO_OK;
}

int pbso_integrator_set_transfer(pbso_integrator* it, const double* transfer, int n_transfer, int L) {
    PBSO_REQUIRE(it, PBSO_ERR_INVALID, "null handle");
    PBSO_REQUIRE(L >= 0 && n_transfer >= 0 && n_transfer <= it->N, PBSO_ERR_INVALID,
                 "transfer size must be <= N (q.head(n).dot(transfer), modal_solver.h:268)");
    PBSO_REQUIRE(L == 0 || transfer, PBSO_ERR_INVALID, "null transfer");
    DeviceGuard g(it->device);
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    const size_t need = (size_t)L * n_transfer;
    if (need > (size_t)it->Lcap) {
        // only the table grows here: d_pos / d_acc / d_cnt keep their own capacities (launch_render and
        // set_transfer_ffat re-size them when they need to)
        cudaFree(it->d_trans); it->d_trans = nullptr; it->Lcap = 0;
        PBSO_CUDA(cudaMalloc(&it->d_trans, sizeof(double) * need));
        it->Lcap = (int)need;
    }
    if (need) PBSO_CUDA(cudaMemcpy(it->d_trans, transfer, sizeof(double) * need, cudaMemcpyHostToDevice));
    it->n_transfer = n_transfer; it->L = L;
    return PBSO_OK;
}

int pbso_integrator_set_transfer_ffat(pbso_integrator* it, const pbso_ffat* maps, int n_transfer, const double* pos, int L) {
    PBSO_REQUIRE(it && maps, PBSO_ERR_INVALID, "null handle");
    PBSO_REQUIRE(L > 0 && pos, PBSO_ERR_INVALID, "need at least one listener position");
    PBSO_REQUIRE(n_transfer > 0 && n_transfer <= it->N, PBSO_ERR_INVALID,
                 "transfer size must be <= N (q.head(n).dot(transfer), modal_solver.h:268)");
    DeviceGuard g(it->device);
    const size_t need = (size_t)L * n_transfer;
    if (need > (size_t)it->Lcap) {                       // grow: wait for renders that still read the old table
        PBSO_CUDA(cudaStreamSynchronize(it->stream));
        cudaFree(it->d_trans); it->d_trans = nullptr; it->Lcap = 0;
        PBSO_CUDA(cudaMalloc(&it->d_trans, sizeof(double) * need));
        it->Lcap = (int)need;
    }
    if (3 * L > it->pos_cap) {
        PBSO_CUDA(cudaStreamSynchronize(it->stream));
        cudaFree(it->d_pos); it->d_pos = nullptr; it->pos_cap = 0;
        PBSO_CUDA(cudaMalloc(&it->d_pos, sizeof(double) * 3 * L));
        it->pos_cap = 3 * L;
    }
    // stream order does the rest: K3 writes the table after earlier renders have read it and before the next one
    PBSO_CUDA(cudaMemcpyAsync(it->d_pos, pos, sizeof(double) * 3 * L, cudaMemcpyHostToDevice, it->stream));
    if (int rc = pbso_ffat_eval_device(maps, n_transfer, it->d_pos, L, it->d_trans, it->stream)) return rc;
    it->n_transfer = n_transfer; it->L = L;
    return PBSO_OK;
}

int pbso_render_buffer_device(pbso_integrator* it, const double* d_space, const double* d_time, int T,
                              double* d_y, double* d_qnorm) {
    PBSO_REQUIRE(it && d_space && d_time && T > 0, PBSO_ERR_INVALID, "bad argument");
    PBSO_REQUIRE(it->L == 0 || d_y, PBSO_ERR_INVALID, "null output");
    DeviceGuard g(it->device);
    return launch_render(it, d_space, d_time, T, d_y, d_qnorm);
}

int pbso_render_buffer(pbso_integrator* it, const double* space, const double* time, int T,
                       double* y_out, double* qnorm_out) {
    PBSO_REQUIRE(it && space && time && T > 0, PBSO_ERR_INVALID, "bad argument");
    PBSO_REQUIRE(it->L == 0 || y_out, PBSO_ERR_INVALID, "null output");
    DeviceGuard g(it->device);
    const int N = it->N, L = it->L;
    if (int rc = ensure_staging(it, T, L)) return rc;
    std::memcpy(it->h_in, space, sizeof(double) * N);
    std::memcpy(it->h_in + N, time, sizeof(double) * T);
    // Zero-copy: the pinned staging buffers are mapped into the device (unified addressing), so K1 reads its 10 KB of
    // input and writes y / qnorm straight across PCIe -- one launch and one synchronisation per buffer, no copy-engine
    // operations (they cost more in launch latency than the bytes do in transfer time).  Larger outputs (many listeners)
    // still go through the copy engine.
    const size_t out_n = (size_t)(L > 0 ? L : 1) * T + (qnorm_out ? N : 0);
    const bool zero_copy = out_n * sizeof(double) <= (64u << 10);
    if (zero_copy) {
        double* m_y = it->h_out;
        double* m_qn = it->h_out + (size_t)(L > 0 ? L : 1) * T;
        if (int rc = launch_render(it, it->h_in, it->h_in + N, T, m_y, qnorm_out ? m_qn : nullptr)) return rc;
    } else {
        PBSO_CUDA(cudaMemcpyAsync(it->d_in, it->h_in, sizeof(double) * (N + T), cudaMemcpyHostToDevice, it->stream));
        double* d_y = it->d_out;
        double* d_qn = it->d_out + (size_t)(L > 0 ? L : 1) * T;
        if (int rc = launch_render(it, it->d_in, it->d_in + N, T, d_y, qnorm_out ? d_qn : nullptr)) return rc;
        PBSO_CUDA(cudaMemcpyAsync(it->h_out, it->d_out, sizeof(double) * out_n, cudaMemcpyDeviceToHost, it->stream));
    }
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    if (L > 0) std::memcpy(y_out, it->h_out, sizeof(double) * (size_t)L * T);
    if (qnorm_out) std::memcpy(qnorm_out, it->h_out + (size_t)(L > 0 ? L : 1) * T, sizeof(double) * N);
    return PBSO_OK;
}

int pbso_integrator_sync(pbso_integrator* it) {
    PBSO_REQUIRE(it, PBSO_ERR_INVALID, "null handle");
    DeviceGuard g(it->device);
    PBSO_CUDA(cudaStreamSynchronize(it->stream));
    return PBSO_OK;
}

}  // extern "C"
