// =============================================================================
// batch.cu -- offline batch renderer: many independent sound objects x long audio.
//
// Each object is its own ModalSolver::step loop (modal_solver.h:181-276) driven by PointForce
// messages (forces.h:81-90) and one static TransMessage; the product is the mix of all objects.
//
//  k_batch_setup   per (object, mode) constants from (a, b, h), FP64 (modal_integrator.h:86-100)
//  k_batch_f64     reference arithmetic: FP64 direct-form recurrence, one thread per mode
//  k_batch_pow     fast path ("pole-power tiles"):
//      The recurrence q_k = c1 q_{k-1} + c2 q_{k-2} has poles w, conj(w), w = eps e^{i theta}, and
//      q[k] = Im v[k] for the complex state v[k] = w v[k-1] (+ c3 Q[k] (wr/wi + i) when forced).
//      Between impulses  q[s+j] = Re v[s] * Im(w^j) + Im v[s] * Re(w^j), so a tile of L samples is
//      a rank-2 update from PRECOMPUTED pole powers -- 2 FMA per mode-sample instead of the
//      recurrence's 4, and no FP32 recurrence at all (FP32 recurrences cannot hold the 1e-6
//      tolerance): the carrier v lives in FP64 and is advanced once per tile with
//      w^L; only the tile evaluation is FP32.
//      Mapping: a warp owns 16 modes; lane j owns sample offsets {j, j+32} of every 64-sample tile
//      and keeps T_m Im(w_m^j), T_m Re(w_m^j) for its 16 modes in registers (64 regs).  Lanes 0..15
//      double as the FP64 carrier owners of the warp's modes (one thread per mode, state in
//      registers) and publish the tile-start states through shared memory as broadcast float4s.
//      The modal sum is then in-thread (no shuffles); warps (mode blocks) meet in shared memory.
//
// HBM layout (SoA, [n_obj][n_modes] doubles): lneps | theta | c1 | c2 | c3 | cot | trans.
// Events: CSR by object (ev_off[n_obj+1], ev_buf[], ev_space[n_events][n_modes]).
// =============================================================================
#include "common.cuh"
#include "batch_tc.cuh"
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

using namespace pbso;

struct pbso_batch {
    int device = 0;
    int n_obj = 0, n_modes = 0;
    double h = 0;
    cudaStream_t stream = nullptr;      // stream in use
    cudaStream_t own_stream = nullptr;  // the handle's own stream
    double* d_par = nullptr;      // 7 arrays of n_obj*n_modes: lneps, theta, c1, c2, c3, cot, trans
    int* d_ev_off = nullptr; int* d_ev_buf = nullptr; double* d_ev_space = nullptr;
    int* d_ev_src = nullptr; double* d_ev_stage = nullptr; size_t ev_cap = 0, ev_stage_cap = 0;
    int n_events = 0;
    double* d_mix = nullptr; size_t mix_cap = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int last_launches = 0;
    int sm_count = 148;
    std::vector<int> h_ev_off, h_ev_buf;      // host copy of the impulse CSR (unit list of the tensor-core path)
    unsigned trans_ver = 0, ev_ver = 0;
    // state at the start of the render (stateful range renders): q[-1] | q[-2] as ModalIntegrator keeps them, and the
    // same state as the complex carrier v0 (q[j] = Im(v0 w^j) for the free response) | 4 arrays of n_obj*n_modes
    double* d_state = nullptr; bool has_state = false;
    TcState* tc = nullptr;
    size_t npm() const { return (size_t)n_obj * n_modes; }
    double* lneps() const { return d_par; }
    double* theta() const { return d_par + npm(); }
    double* c1() const { return d_par + 2 * npm(); }
    double* c2() const { return d_par + 3 * npm(); }
    double* c3() const { return d_par + 4 * npm(); }
    double* cot() const { return d_par + 5 * npm(); }
    double* trans() const { return d_par + 6 * npm(); }
    const double* q10() const { return has_state ? d_state : nullptr; }
    const double* q20() const { return has_state ? d_state + npm() : nullptr; }
    const double* v0r() const { return has_state ? d_state + 2 * npm() : nullptr; }
    const double* v0i() const { return has_state ? d_state + 3 * npm() : nullptr; }
};

// (q[-1], q[-2]) -> carrier at sample 0: with u = x + i q[-1] the state one sample earlier, Im(u / w) = q[-2] gives
// x = (q[-1] cos(theta) - eps q[-2]) / sin(theta), and v0 = u w.  Im(v0) = c1 q[-1] + c2 q[-2], the recurrence's q[0].
__global__ void k_batch_state_to_carrier(size_t n, const double* __restrict__ lneps, const double* __restrict__ theta,
                                         const double* __restrict__ q1, const double* __restrict__ q2,
                                         double* __restrict__ v0r, double* __restrict__ v0i) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c; sincos(theta[i], &s, &c);
    const double eps = exp(lneps[i]);
    const double x = (q1[i] * c - eps * q2[i]) / s;
    v0r[i] = eps * (x * c - q1[i] * s);
    v0i[i] = eps * (x * s + q1[i] * c);
}

// State after n_samples: v_N = v0 w^N + sum_e inj_e w^(N - t_e) in closed form (FP64 exp / sincos of the total angle per
// term), then q[N-1] = Im(v_N / w), q[N-2] = Im(v_N / w^2) -- what ModalIntegrator would hold after the same steps.
__global__ void k_batch_end_state(int n_obj, int n_modes, long long n_samples, int BUF,
                                  const double* __restrict__ lneps, const double* __restrict__ theta,
                                  const double* __restrict__ c3a, const double* __restrict__ cota,
                                  const int* __restrict__ ev_off, const int* __restrict__ ev_buf, const double* __restrict__ ev_space,
                                  const double* __restrict__ v0r, const double* __restrict__ v0i,
                                  double* __restrict__ q1, double* __restrict__ q2) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)n_obj * n_modes) return;
    const int o = (int)(i / n_modes), m = (int)(i % n_modes);
    const double le = lneps[i], th = theta[i];
    auto powr = [&](double k, double& pr, double& pi) { double s, c; sincos(k * th, &s, &c); const double e = exp(k * le); pr = e * c; pi = e * s; };
    double vr = 0.0, vi = 0.0, pr, pi;
    if (v0r) { powr((double)n_samples, pr, pi); vr = v0r[i] * pr - v0i[i] * pi; vi = v0r[i] * pi + v0i[i] * pr; }
    const double inji = c3a[i], injr = inji * cota[i];
    for (int e = ev_off[o]; e < ev_off[o + 1]; ++e) {
        const long long t = (long long)ev_buf[e] * BUF;
        if (t >= n_samples) break;
        powr((double)(n_samples - t), pr, pi);
        const double sp = ev_space[(size_t)e * n_modes + m];
        vr += sp * (injr * pr - inji * pi); vi += sp * (injr * pi + inji * pr);
    }
    powr(-1.0, pr, pi); q1[i] = vr * pi + vi * pr;
    powr(-2.0, pr, pi); q2[i] = vr * pi + vi * pr;
}

// row r of dst = row src[r] of stage (impulse scripts that arrive unsorted)
__global__ void k_gather_rows(int n_modes, const int* __restrict__ src, const double* __restrict__ stage,
                              double* __restrict__ dst) {
    const size_t r = blockIdx.x;
    const double* in = stage + (size_t)src[r] * n_modes;
    for (int m = threadIdx.x; m < n_modes; m += blockDim.x) dst[r * n_modes + m] = in[m];
}

__global__ void k_batch_setup(size_t n, double h, const double* __restrict__ a, const double* __restrict__ b,
                              double* __restrict__ lneps, double* __restrict__ theta, double* __restrict__ c1,
                              double* __restrict__ c2, double* __restrict__ c3, double* __restrict__ cot) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ai = a[i], bi = b[i];
    const double le = -ai / 2 * h;                                   // log of modal_integrator.h:89
    const double epsilon = exp(le);
    const double th = h * sqrt(bi - ai * ai / 4.0);                  // :90
    const double gamma = asin(ai / (2.0 * sqrt(bi)));                // :91
    const double omega = sqrt(bi), omega_d = sqrt(bi - ai * ai / 4.0);   // :92-93
    double s, c; sincos(th, &s, &c);
    lneps[i] = le; theta[i] = th;
    c1[i] = 2.0 * epsilon * c;                                       // :95
    c2[i] = -(epsilon * epsilon);                                    // :96
    double v = 2.0 * (epsilon * cos(th + gamma) - epsilon * epsilon * cos(2.0 * th + gamma));   // :97
    v /= (3.0 * omega * omega_d);                                    // :98
    c3[i] = v * 1E9;                                                 // :99
    cot[i] = c / s;                                                  // wr/wi
}

// ---------------------------------------------------------------------------------------------
// FP64 reference-arithmetic kernel: CTA = (object, slab of 256 modes); thread = mode.
// ---------------------------------------------------------------------------------------------
constexpr int BF_TPB = 256;

__device__ __forceinline__ double bshfl_xor_f64(double v, int mask) {
    int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v), mask);
    int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v), mask);
    return __hiloint2double(hi, lo);
}

__global__ void __launch_bounds__(BF_TPB)
k_batch_f64(int n_modes, int slabs, int BUF, int n_buf,
            const double* __restrict__ c1a, const double* __restrict__ c2a, const double* __restrict__ c3a,
            const double* __restrict__ trans, const int* __restrict__ ev_off, const int* __restrict__ ev_buf,
            const double* __restrict__ ev_space, const double* __restrict__ q10, const double* __restrict__ q20,
            double* __restrict__ mix, float* __restrict__ stems) {
    __shared__ double s_part[BF_TPB / 32][32];
    const int obj = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const int m = slab * BF_TPB + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = m < n_modes;
    const size_t pm = (size_t)obj * n_modes + (live ? m : 0);
    const double c1 = live ? c1a[pm] : 0.0, c2 = live ? c2a[pm] : 0.0, c3 = live ? c3a[pm] : 0.0;
    const double T = live ? trans[pm] : 0.0;
    double q1 = (live && q10) ? q10[pm] : 0.0, q2 = (live && q20) ? q20[pm] : 0.0;
    int ev = ev_off[obj];
    const int ev_end = ev_off[obj + 1];
    int next_buf = ev < ev_end ? ev_buf[ev] : INT_MAX;
    for (int bi = 0; bi < n_buf; ++bi) {
        double Q0 = 0.0;                                   // space * time(0), PointForce: time(0) = 1
        if (bi == next_buf) {
            if (live) Q0 = ev_space[(size_t)ev * n_modes + m];
            ++ev; next_buf = ev < ev_end ? ev_buf[ev] : INT_MAX;
        }
        for (int t0 = 0; t0 < BUF; t0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                double qk = 0.0;
                if (t0 + j < BUF) {
                    const double Q = (t0 + j == 0) ? Q0 : 0.0;
                    qk = c1 * q1 + c2 * q2 + c3 * Q;       // modal_integrator.h:109-110
                    q2 = q1; q1 = qk;
                }
                v[j] = T * qk;                             // modal_solver.h:267-269
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const bool up = (lane & s) != 0;
#pragma unroll
                for (int i = 0; i < s; ++i) {
                    const double keep = up ? v[i + s] : v[i];
                    const double send = up ? v[i] : v[i + s];
                    v[i] = keep + bshfl_xor_f64(send, s);
                }
            }
            s_part[warp][lane] = v[0];
            __syncthreads();
            if (threadIdx.x < 32 && t0 + threadIdx.x < BUF) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < BF_TPB / 32; ++w) s += s_part[w][threadIdx.x];
                const size_t i = (size_t)bi * BUF + t0 + threadIdx.x;
                if (mix) atomicAdd(&mix[i], s);
                if (stems) atomicAdd(&stems[(size_t)obj * n_buf * BUF + i], (float)s);
            }
            __syncthreads();
        }
    }
}

// The "scalar tail" of the tensor-core path when buf_size is not a multiple of its 128-sample tiles (the reference's
// default FRAMES_PER_BUFFER is 513): an impulse then lands inside a tile, and its samples up to the next tile boundary --
// at most 127 -- are rendered here with the reference recurrence, one thread per mode, CTA = (event, slab of 256 modes);
// from the boundary on the impulse is part of the contraction (k_tc_impulse carries its state there).
__global__ void __launch_bounds__(BF_TPB)
k_batch_event_heads(int n_modes, int slabs, int BUF, long long n_samples, int e0,
                    const double* __restrict__ c1a, const double* __restrict__ c2a, const double* __restrict__ c3a,
                    const double* __restrict__ trans, const int* __restrict__ ev_obj, const int* __restrict__ ev_buf,
                    const double* __restrict__ ev_space, double* __restrict__ mix, float* __restrict__ stems) {
    __shared__ double s_part[BF_TPB / 32][32];
    const int e = e0 + blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const long long t_e = (long long)ev_buf[e] * BUF;
    const int s0 = (int)(t_e % 128);
    if (s0 == 0 || t_e >= n_samples) return;
    const int len = (int)min((long long)(128 - s0), n_samples - t_e);
    const int obj = ev_obj[e];
    const int m = slab * BF_TPB + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = m < n_modes;
    const size_t pm = (size_t)obj * n_modes + (live ? m : 0);
    const double c1 = live ? c1a[pm] : 0.0, c2 = live ? c2a[pm] : 0.0, c3 = live ? c3a[pm] : 0.0, T = live ? trans[pm] : 0.0;
    const double Q0 = live ? ev_space[(size_t)e * n_modes + m] : 0.0;
    double q1 = 0.0, q2 = 0.0;
    for (int t0 = 0; t0 < len; t0 += 32) {
        double v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            double qk = 0.0;
            if (t0 + j < len) {
                qk = c1 * q1 + c2 * q2 + c3 * ((t0 + j == 0) ? Q0 : 0.0);       // modal_integrator.h:109-110
                q2 = q1; q1 = qk;
            }
            v[j] = T * qk;
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < s; ++i) {
                const double keep = up ? v[i + s] : v[i];
                const double send = up ? v[i] : v[i + s];
                v[i] = keep + bshfl_xor_f64(send, s);
            }
        }
        s_part[warp][lane] = v[0];
        __syncthreads();
        if (threadIdx.x < 32 && t0 + threadIdx.x < len) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < BF_TPB / 32; ++w) s += s_part[w][threadIdx.x];
            const long long i = t_e + t0 + threadIdx.x;
            if (mix) atomicAdd(&mix[i], s);
            if (stems) atomicAdd(&stems[(size_t)obj * n_samples + i], (float)s);
        }
        __syncthreads();
    }
}

int pbso::batch_event_heads(const TcArgs& a, int e0, int ne, const int* d_ev_obj) {
    const int slabs = div_up(a.n_modes, BF_TPB);
    k_batch_event_heads<<<ne * slabs, BF_TPB, 0, a.stream>>>(a.n_modes, slabs, a.buf_size, (long long)a.buf_size * a.n_buffers, e0, a.c1, a.c2, a.c3,
                                                             a.trans, d_ev_obj, a.d_ev_buf, a.d_ev_space, a.d_mix, a.d_stems);
    PBSO_CUDA(cudaGetLastError());
    return PBSO_OK;
}

// ---------------------------------------------------------------------------------------------
// Fast path: pole-power tiles.  CTA = (object, slab of 256 modes), 16 warps x 16 modes.
// ---------------------------------------------------------------------------------------------
constexpr int FB_WARPS = 16;
constexpr int FB_MB = 16;                       // modes per warp
constexpr int FB_SLAB = FB_WARPS * FB_MB;       // modes per CTA
constexpr int FB_L = 64;                        // tile length (2 sample offsets per lane)

template <int TT>                               // tiles per buffer: BUF = 64 * TT
__global__ void __launch_bounds__(FB_WARPS * 32, 1)
k_batch_pow(int n_modes, int slabs, int n_buf,
            const double* __restrict__ lneps, const double* __restrict__ theta,
            const double* __restrict__ c3a, const double* __restrict__ cota, const double* __restrict__ trans,
            const int* __restrict__ ev_off, const int* __restrict__ ev_buf, const double* __restrict__ ev_space,
            const double* __restrict__ v0r, const double* __restrict__ v0i,
            double* __restrict__ mix, float* __restrict__ stems) {
    constexpr int BUF = FB_L * TT;
    __shared__ __align__(16) float sV[FB_WARPS][FB_MB][2 * TT];      // tile-start states (re[TT] | im[TT])
    __shared__ __align__(16) float sY[2][FB_WARPS][BUF];             // per-warp partial sums, double-buffered
    const int obj = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_base = slab * FB_SLAB + warp * FB_MB;
    const size_t obase = (size_t)obj * n_modes;

    // ---- pole-power tables for this lane's two sample offsets, FP64 -> FP32 ----------------
    float A[FB_MB][2], B[FB_MB][2];
#pragma unroll
    for (int i = 0; i < FB_MB; ++i) {
        const int m = m_base + i;
        if (m < n_modes) {
            const double le = lneps[obase + m], th = theta[obase + m], T = trans[obase + m];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const double j = (double)(lane + 32 * jj);
                double s, c; sincos(j * th, &s, &c);
                const double e = T * exp(j * le);
                A[i][jj] = (float)(e * s);                 // T Im(w^j)
                B[i][jj] = (float)(e * c);                 // T Re(w^j)
            }
        } else {
            A[i][0] = A[i][1] = B[i][0] = B[i][1] = 0.f;
        }
    }
    // ---- FP64 carrier: lanes 0..15 own the warp's modes ----------------------------------
    const int my_m = m_base + (lane & (FB_MB - 1));
    const bool owner = lane < FB_MB && my_m < n_modes;
    double vr = 0.0, vi = 0.0, Wr = 0.0, Wi = 0.0, injr = 0.0, inji = 0.0;
    if (owner) {
        const double le = lneps[obase + my_m], th = theta[obase + my_m];
        double s, c; sincos((double)FB_L * th, &s, &c);
        const double e = exp((double)FB_L * le);
        Wr = e * c; Wi = e * s;                            // w^L
        inji = c3a[obase + my_m];                          // Im part of c3 (wr/wi + i)
        injr = inji * cota[obase + my_m];
        if (v0r) { vr = v0r[obase + my_m]; vi = v0i[obase + my_m]; }   // stateful range render
    }
    int ev = ev_off[obj];
    const int ev_end = ev_off[obj + 1];
    int next_buf = ev < ev_end ? ev_buf[ev] : INT_MAX;

    for (int bi = 0; bi < n_buf; ++bi) {
        if (bi == next_buf) {                              // PointForce lands on sample 0 (forces.h:87)
            if (owner) {
                const double sp = ev_space[(size_t)ev * n_modes + my_m];
                vr = fma(injr, sp, vr); vi = fma(inji, sp, vi);
            }
            ++ev; next_buf = ev < ev_end ? ev_buf[ev] : INT_MAX;
        }
        if (lane < FB_MB) {
            float re[TT], im[TT];
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                re[t] = (float)vr; im[t] = (float)vi;
                const double nr = vr * Wr - vi * Wi;       // v <- v w^L
                vi = vr * Wi + vi * Wr; vr = nr;
            }
#pragma unroll
            for (int t = 0; t < TT; ++t) { sV[warp][lane][t] = re[t]; sV[warp][lane][TT + t] = im[t]; }
        }
        __syncwarp();
        float acc[TT][2];
#pragma unroll
        for (int t = 0; t < TT; ++t) acc[t][0] = acc[t][1] = 0.f;
#pragma unroll
        for (int i = 0; i < FB_MB; ++i) {
            float re[TT], im[TT];
            if (TT == 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(&sV[warp][i][0]);
                const float4 i4 = *reinterpret_cast<const float4*>(&sV[warp][i][TT]);
                re[0] = r4.x; re[1 % TT] = r4.y; re[2 % TT] = r4.z; re[3 % TT] = r4.w;
                im[0] = i4.x; im[1 % TT] = i4.y; im[2 % TT] = i4.z; im[3 % TT] = i4.w;
            } else {
#pragma unroll
                for (int t = 0; t < TT; ++t) { re[t] = sV[warp][i][t]; im[t] = sV[warp][i][TT + t]; }
            }
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                acc[t][0] = fmaf(re[t], A[i][0], acc[t][0]);
                acc[t][1] = fmaf(re[t], A[i][1], acc[t][1]);
                acc[t][0] = fmaf(im[t], B[i][0], acc[t][0]);
                acc[t][1] = fmaf(im[t], B[i][1], acc[t][1]);
            }
        }
        __syncwarp();
        float* yb = &sY[bi & 1][warp][0];
#pragma unroll
        for (int t = 0; t < TT; ++t) { yb[t * FB_L + lane] = acc[t][0]; yb[t * FB_L + 32 + lane] = acc[t][1]; }
        __syncthreads();
        if (threadIdx.x < BUF) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < FB_WARPS; ++w) s += sY[bi & 1][w][threadIdx.x];
            const size_t i = (size_t)bi * BUF + threadIdx.x;
            if (mix) atomicAdd(&mix[i], (double)s);
            if (stems) {
                if (slabs == 1) stems[(size_t)obj * n_buf * BUF + i] = s;
                else atomicAdd(&stems[(size_t)obj * n_buf * BUF + i], s);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Generalised pole-power kernel (BUF = 256): WARPS x MB modes per CTA, JJ sample offsets per lane
// (tile length L = 32 JJ, TT = 256 / L tiles per buffer).  F2 packs sample offsets (j, j+32) into
// fma.rn.f32x2 (SASS FFMA2) with the tile-start state as the broadcast scalar operand.
// Shared-memory wavefronts per FMA fall as 1/JJ (each broadcast state word feeds JJ FMAs per lane);
// table registers grow as 2 JJ MB, which is what bounds JJ.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long p, float v) {
    unsigned long long vv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(p), "l"(vv));
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <int WARPS, int MB, int JJ, bool F2, int MINB = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_batch_pow_g(int n_modes, int slabs, int n_buf, int n_chunks, int bufs_per_chunk,
              const double* __restrict__ lneps, const double* __restrict__ theta,
              const double* __restrict__ c3a, const double* __restrict__ cota, const double* __restrict__ trans,
              const int* __restrict__ ev_off, const int* __restrict__ ev_buf, const double* __restrict__ ev_space,
              const double* __restrict__ v0r, const double* __restrict__ v0i,
              double* __restrict__ mix, float* __restrict__ stems) {
    constexpr int BUF = 256, L = 32 * JJ, TT = BUF / L, SLAB = WARPS * MB, NT = WARPS * 32;
    static_assert(BUF % L == 0 && (JJ % 2 == 0) && MB <= 32, "bad tile configuration");
    __shared__ __align__(16) float sV[WARPS][MB][2 * TT];
    __shared__ __align__(16) float sY[2][WARPS][BUF];
    // blockIdx.x = (object, slab, time chunk): chunks are independent because a chunk's start state follows in
    // closed form from the impulses before it (v = sum_e inj_e w^(256 (b0 - b_e))), evaluated in FP64 below.
    const int chunk = blockIdx.x % n_chunks;
    const int obj = (blockIdx.x / n_chunks) / slabs, slab = (blockIdx.x / n_chunks) % slabs;
    const int b_begin = chunk * bufs_per_chunk;
    const int b_end = min(n_buf, b_begin + bufs_per_chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_base = slab * SLAB + warp * MB;
    const size_t obase = (size_t)obj * n_modes;

    float A[MB][JJ], B[MB][JJ];
#pragma unroll
    for (int i = 0; i < MB; ++i) {
        const int m = m_base + i;
        if (m < n_modes) {
            const double le = lneps[obase + m], th = theta[obase + m], T = trans[obase + m];
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const double j = (double)(lane + 32 * jj);
                double s, c; sincos(j * th, &s, &c);
                const double e = T * exp(j * le);
                A[i][jj] = (float)(e * s); B[i][jj] = (float)(e * c);
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) A[i][jj] = B[i][jj] = 0.f;
        }
    }
    unsigned long long A2[MB][JJ / 2], B2[MB][JJ / 2];
    if (F2) {
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
            for (int p = 0; p < JJ / 2; ++p) { A2[i][p] = pack2(A[i][2 * p], A[i][2 * p + 1]); B2[i][p] = pack2(B[i][2 * p], B[i][2 * p + 1]); }
    }
    const int my_m = m_base + (lane % MB);
    const bool owner = lane < MB && my_m < n_modes;
    double vr = 0.0, vi = 0.0, Wr = 0.0, Wi = 0.0, injr = 0.0, inji = 0.0;
    if (owner) {
        const double le = lneps[obase + my_m], th = theta[obase + my_m];
        double s, c; sincos((double)L * th, &s, &c);
        const double e = exp((double)L * le);
        Wr = e * c; Wi = e * s;
        inji = c3a[obase + my_m];
        injr = inji * cota[obase + my_m];
    }
    int ev = ev_off[obj];
    const int ev_end = ev_off[obj + 1];
    if (owner && v0r) {                                               // stateful range render: v0 w^(samples before the chunk)
        const double n = (double)BUF * (double)b_begin;
        double s, c; sincos(n * theta[obase + my_m], &s, &c);
        const double e = exp(n * lneps[obase + my_m]);
        const double ar = v0r[obase + my_m], ai = v0i[obase + my_m];
        vr = e * (ar * c - ai * s); vi = e * (ar * s + ai * c);
    }
    // impulses before this chunk: advance each with the exact pole power (FP64 exp / sincos of the total angle)
    while (ev < ev_end && ev_buf[ev] < b_begin) {
        if (owner) {
            const double n = (double)BUF * (double)(b_begin - ev_buf[ev]);
            const double le = lneps[obase + my_m], th = theta[obase + my_m];
            double s, c; sincos(n * th, &s, &c);
            const double e = exp(n * le);
            const double pr = e * c, pi = e * s;                    // w^n
            const double sp = ev_space[(size_t)ev * n_modes + my_m];
            const double ir = injr * sp, ii = inji * sp;
            vr += ir * pr - ii * pi; vi += ir * pi + ii * pr;
        }
        ++ev;
    }
    int next_buf = ev < ev_end ? ev_buf[ev] : INT_MAX;

    for (int bi = b_begin; bi < b_end; ++bi) {
        if (bi == next_buf) {
            if (owner) {
                const double sp = ev_space[(size_t)ev * n_modes + my_m];
                vr = fma(injr, sp, vr); vi = fma(inji, sp, vi);
            }
            ++ev; next_buf = ev < ev_end ? ev_buf[ev] : INT_MAX;
        }
        if (lane < MB) {
            float st[2 * TT];
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                st[t] = (float)vr; st[TT + t] = (float)vi;
                const double nr = vr * Wr - vi * Wi;
                vi = vr * Wi + vi * Wr; vr = nr;
            }
            if (2 * TT == 4) *reinterpret_cast<float4*>(&sV[warp][lane][0]) = make_float4(st[0], st[1], st[2], st[3]);
            else {
#pragma unroll
                for (int k = 0; k < 2 * TT; ++k) sV[warp][lane][k] = st[k];
            }
        }
        __syncwarp();
        float acc[TT][JJ];
        unsigned long long acc2[TT][JJ / 2];
#pragma unroll
        for (int t = 0; t < TT; ++t) {
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) acc[t][jj] = 0.f;
#pragma unroll
            for (int p = 0; p < JJ / 2; ++p) acc2[t][p] = 0ull;
        }
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            float st[2 * TT];
            if (2 * TT == 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(&sV[warp][i][0]);
                st[0] = v4.x; st[1] = v4.y; st[2 % (2 * TT)] = v4.z; st[3 % (2 * TT)] = v4.w;
            } else if (2 * TT == 8) {
                const float4 v4 = *reinterpret_cast<const float4*>(&sV[warp][i][0]);
                const float4 w4 = *reinterpret_cast<const float4*>(&sV[warp][i][4]);
                st[0] = v4.x; st[1] = v4.y; st[2] = v4.z; st[3] = v4.w;
                st[4 % (2 * TT)] = w4.x; st[5 % (2 * TT)] = w4.y; st[6 % (2 * TT)] = w4.z; st[7 % (2 * TT)] = w4.w;
            } else {
                const float2 v2 = *reinterpret_cast<const float2*>(&sV[warp][i][0]);
                st[0] = v2.x; st[1 % (2 * TT)] = v2.y;
            }
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (F2) {
#pragma unroll
                    for (int p = 0; p < JJ / 2; ++p) { ffma2(acc2[t][p], A2[i][p], st[t]); ffma2(acc2[t][p], B2[i][p], st[TT + t]); }
                } else {
#pragma unroll
                    for (int jj = 0; jj < JJ; ++jj) {
                        acc[t][jj] = fmaf(st[t], A[i][jj], acc[t][jj]);
                        acc[t][jj] = fmaf(st[TT + t], B[i][jj], acc[t][jj]);
                    }
                }
            }
        }
        __syncwarp();
        float* yb = &sY[bi & 1][warp][0];
#pragma unroll
        for (int t = 0; t < TT; ++t) {
            if (F2) {
#pragma unroll
                for (int p = 0; p < JJ / 2; ++p) unpack2(acc2[t][p], acc[t][2 * p], acc[t][2 * p + 1]);
            }
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) yb[t * L + 32 * jj + lane] = acc[t][jj];
        }
        __syncthreads();
        for (int o = threadIdx.x; o < BUF; o += NT) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) s += sY[bi & 1][w][o];
            const size_t i = (size_t)bi * BUF + o;
            if (mix) atomicAdd(&mix[i], (double)s);
            if (stems) {
                if (slabs == 1) stems[(size_t)obj * n_buf * BUF + i] = s;
                else atomicAdd(&stems[(size_t)obj * n_buf * BUF + i], s);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
static int launch_render(pbso_batch* bt, int buf_size, int n_buffers, int precision, int n_chunks, double* d_mix, float* d_stems) {
    PBSO_REQUIRE(buf_size > 0 && n_buffers > 0, PBSO_ERR_INVALID, "buf_size and n_buffers must be > 0");
    PBSO_REQUIRE(bt->d_ev_off, PBSO_ERR_INVALID, "no impulse script: call pbso_batch_set_impulses first");
    const size_t ns = (size_t)buf_size * n_buffers;
    if (d_mix) PBSO_CUDA(cudaMemsetAsync(d_mix, 0, sizeof(double) * ns, bt->stream));
    if (d_stems) PBSO_CUDA(cudaMemsetAsync(d_stems, 0, sizeof(float) * ns * bt->n_obj, bt->stream));
    PBSO_CUDA(cudaEventRecord(bt->e0, bt->stream));
    if (precision == PBSO_PREC_F32_TILED) {
        const int slabs = div_up(bt->n_modes, FB_SLAB);
        const int grid = bt->n_obj * slabs;
#define PBSO_LAUNCH_POW(TT)                                                                              \
        k_batch_pow<TT><<<grid, FB_WARPS * 32, 0, bt->stream>>>(bt->n_modes, slabs, n_buffers, bt->lneps(), \
            bt->theta(), bt->c3(), bt->cot(), bt->trans(), bt->d_ev_off, bt->d_ev_buf, bt->d_ev_space, bt->v0r(), bt->v0i(), d_mix, d_stems)
        // 16 warps x 16 modes, 64-sample tiles, packed FFMA2: the best of the variants measured in round 1
        // (profiles/r1_k_batch_pow.md); the others are gone
#define PBSO_LAUNCH_G(W, MB, JJ, ...)                                                                      \
        do { const int sl = div_up(bt->n_modes, (W) * (MB));                                               \
             /* time chunks: enough CTAs for ~4 waves of SMs when there are few objects; >= 8 buffers each */ \
             int nc = n_chunks > 0 ? n_chunks : div_up(4 * bt->sm_count, bt->n_obj * sl);                  \
             nc = std::max(1, std::min(nc, div_up(n_buffers, 8)));                                         \
             const int bpc = div_up(n_buffers, nc); nc = div_up(n_buffers, bpc);                           \
             k_batch_pow_g<W, MB, JJ, __VA_ARGS__><<<bt->n_obj * sl * nc, (W) * 32, 0, bt->stream>>>(bt->n_modes, sl, n_buffers, nc, bpc, \
                 bt->lneps(), bt->theta(), bt->c3(), bt->cot(), bt->trans(), bt->d_ev_off, bt->d_ev_buf,  \
                 bt->d_ev_space, bt->v0r(), bt->v0i(), d_mix, d_stems); } while (0)
        if (buf_size == 256) PBSO_LAUNCH_G(16, 16, 2, true);
        else if (buf_size == 64) PBSO_LAUNCH_POW(1);
        else if (buf_size == 128) PBSO_LAUNCH_POW(2);
        else return set_error(PBSO_ERR_UNSUPPORTED, "PBSO_PREC_F32_TILED needs buf_size in {64,128,256}; got %d (use PBSO_PREC_F64)", buf_size);
#undef PBSO_LAUNCH_POW
    } else if (precision == PBSO_PREC_TC3X) {

        TcArgs ta{bt->n_obj, bt->n_modes, buf_size, n_buffers, bt->sm_count, bt->lneps(), bt->theta(), bt->c1(), bt->c2(), bt->c3(), bt->cot(), bt->trans(),
                  bt->h_ev_off.data(), bt->h_ev_buf.data(), bt->d_ev_off, bt->d_ev_buf, bt->d_ev_space, bt->n_events,
                  bt->trans_ver, bt->ev_ver, bt->v0r(), bt->v0i(), d_mix, d_stems, bt->stream};
        int nl = 0;
        if (int rc = tc_render(&bt->tc, ta, &nl)) return rc;
        PBSO_CUDA(cudaEventRecord(bt->e1, bt->stream));
        bt->last_launches = nl;
        return PBSO_OK;
    } else if (precision == PBSO_PREC_F64) {
        const int slabs = div_up(bt->n_modes, BF_TPB);
        k_batch_f64<<<bt->n_obj * slabs, BF_TPB, 0, bt->stream>>>(bt->n_modes, slabs, buf_size, n_buffers, bt->c1(),
            bt->c2(), bt->c3(), bt->trans(), bt->d_ev_off, bt->d_ev_buf, bt->d_ev_space, bt->q10(), bt->q20(), d_mix, d_stems);
    } else {
        return set_error(PBSO_ERR_INVALID, "unknown precision %d", precision);
    }
    PBSO_CUDA(cudaGetLastError());
    PBSO_CUDA(cudaEventRecord(bt->e1, bt->stream));
    bt->last_launches = 1;
    return PBSO_OK;
}

extern "C" {

int pbso_batch_create(int n_obj, int n_modes, double h, const double* a, const double* b, pbso_batch** out) {
    PBSO_REQUIRE(out, PBSO_ERR_INVALID, "null output handle");
    *out = nullptr;
    PBSO_REQUIRE(n_obj > 0 && n_modes > 0 && a && b, PBSO_ERR_INVALID, "bad argument");
    if (int rc = check_device()) return rc;
    pbso_batch* bt = new pbso_batch();
    bt->n_obj = n_obj; bt->n_modes = n_modes; bt->h = h;
    PBSO_CUDA(cudaGetDevice(&bt->device));
    PBSO_CUDA(cudaDeviceGetAttribute(&bt->sm_count, cudaDevAttrMultiProcessorCount, bt->device));
    PBSO_CUDA(cudaStreamCreateWithFlags(&bt->own_stream, cudaStreamNonBlocking));
    bt->stream = bt->own_stream;
    PBSO_CUDA(cudaEventCreate(&bt->e0)); PBSO_CUDA(cudaEventCreate(&bt->e1));
    const size_t n = bt->npm();
    PBSO_CUDA(cudaMalloc(&bt->d_par, sizeof(double) * 7 * n));
    double* d_ab; PBSO_CUDA(cudaMalloc(&d_ab, sizeof(double) * 2 * n));
    PBSO_CUDA(cudaMemcpyAsync(d_ab, a, sizeof(double) * n, cudaMemcpyHostToDevice, bt->stream));
    PBSO_CUDA(cudaMemcpyAsync(d_ab + n, b, sizeof(double) * n, cudaMemcpyHostToDevice, bt->stream));
    k_batch_setup<<<(unsigned)((n + 255) / 256), 256, 0, bt->stream>>>(n, h, d_ab, d_ab + n, bt->lneps(), bt->theta(),
                                                                        bt->c1(), bt->c2(), bt->c3(), bt->cot());
    PBSO_CUDA(cudaGetLastError());
    // default transfer: TransMessage::setToUnit (modal_solver.h:89-92)
    std::vector<double> unit(n, 1E7);
    PBSO_CUDA(cudaMemcpyAsync(bt->trans(), unit.data(), sizeof(double) * n, cudaMemcpyHostToDevice, bt->stream));
    PBSO_CUDA(cudaStreamSynchronize(bt->stream));
    cudaFree(d_ab);
    *out = bt;
    return PBSO_OK;
}

int pbso_batch_destroy(pbso_batch* bt) {
    if (!bt) return PBSO_OK;
    DeviceGuard g(bt->device);
    if (bt->stream) cudaStreamSynchronize(bt->stream);
    cudaFree(bt->d_par); cudaFree(bt->d_ev_off); cudaFree(bt->d_ev_buf); cudaFree(bt->d_ev_space); cudaFree(bt->d_mix); cudaFree(bt->d_ev_src); cudaFree(bt->d_ev_stage); cudaFree(bt->d_state);
    tc_free(bt->tc);
    if (bt->e0) cudaEventDestroy(bt->e0);
    if (bt->e1) cudaEventDestroy(bt->e1);
    if (bt->own_stream) cudaStreamDestroy(bt->own_stream);
    delete bt;
    return PBSO_OK;
}

static int set_transfer_impl(pbso_batch* bt, const double* trans, bool wait) {
    PBSO_REQUIRE(bt && trans, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(bt->device);
    PBSO_CUDA(cudaMemcpyAsync(bt->trans(), trans, sizeof(double) * bt->npm(), cudaMemcpyHostToDevice, bt->stream));
    if (wait) PBSO_CUDA(cudaStreamSynchronize(bt->stream));          // the caller may reuse its buffer on return
    ++bt->trans_ver;
    return PBSO_OK;
}
int pbso_batch_set_transfer(pbso_batch* bt, const double* trans) { return set_transfer_impl(bt, trans, true); }
int pbso_batch_set_transfer_async(pbso_batch* bt, const double* trans) { return set_transfer_impl(bt, trans, false); }

static int set_impulses_impl(pbso_batch* bt, int n_events, const int* obj, const int* buf, const double* space, bool wait);
int pbso_batch_set_impulses(pbso_batch* bt, int n_events, const int* obj, const int* buf, const double* space) {
    return set_impulses_impl(bt, n_events, obj, buf, space, true);
}
int pbso_batch_set_impulses_async(pbso_batch* bt, int n_events, const int* obj, const int* buf, const double* space) {
    return set_impulses_impl(bt, n_events, obj, buf, space, false);
}

static int set_impulses_impl(pbso_batch* bt, int n_events, const int* obj, const int* buf, const double* space, bool wait) {
    PBSO_REQUIRE(bt && n_events >= 0 && (n_events == 0 || (obj && buf && space)), PBSO_ERR_INVALID, "bad argument");
    DeviceGuard g(bt->device);
    // sort events by (object, buffer) into CSR; a ModalSolver dequeues at most one ForceMessage per
    // step (modal_solver.h:184), so two messages of one object cannot share a buffer.
    std::vector<int> order(n_events);
    std::iota(order.begin(), order.end(), 0);
    for (int e = 0; e < n_events; ++e) {
        if (obj[e] < 0 || obj[e] >= bt->n_obj) return set_error(PBSO_ERR_RANGE, "event %d: object %d out of range", e, obj[e]);
        if (buf[e] < 0) return set_error(PBSO_ERR_RANGE, "event %d: negative buffer index", e);
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        return obj[x] != obj[y] ? obj[x] < obj[y] : buf[x] < buf[y]; });
    for (int i = 1; i < n_events; ++i)
        if (obj[order[i]] == obj[order[i - 1]] && buf[order[i]] == buf[order[i - 1]])
            return set_error(PBSO_ERR_INVALID, "object %d has two messages for buffer %d: step() dequeues one per buffer",
                             obj[order[i]], buf[order[i]]);
    std::vector<int> off(bt->n_obj + 1, 0), sbuf(std::max(n_events, 1)), src(std::max(n_events, 1));
    bool identity = true;
    for (int i = 0; i < n_events; ++i) {
        const int e = order[i];
        off[obj[e] + 1]++;
        sbuf[i] = buf[e]; src[i] = e;
        identity &= (e == i);
    }
    for (int o = 0; o < bt->n_obj; ++o) off[o + 1] += off[o];
    // device buffers are kept across calls and only grow; the rows of `space` go straight from the
    // caller's (ideally pinned) memory to the device -- when the events arrive unsorted they are
    // permuted there, never through a pageable host copy.
    const size_t rows = (size_t)std::max(n_events, 1);
    if (!bt->d_ev_off) PBSO_CUDA(cudaMalloc(&bt->d_ev_off, sizeof(int) * off.size()));
    if (rows > bt->ev_cap) {
        PBSO_CUDA(cudaStreamSynchronize(bt->stream));                // renders in flight still read the old buffers
        cudaFree(bt->d_ev_buf); cudaFree(bt->d_ev_space); cudaFree(bt->d_ev_src);
        bt->d_ev_buf = nullptr; bt->d_ev_space = nullptr; bt->d_ev_src = nullptr; bt->ev_cap = 0;
        PBSO_CUDA(cudaMalloc(&bt->d_ev_buf, sizeof(int) * rows));
        PBSO_CUDA(cudaMalloc(&bt->d_ev_src, sizeof(int) * rows));
        PBSO_CUDA(cudaMalloc(&bt->d_ev_space, sizeof(double) * rows * bt->n_modes));
        bt->ev_cap = rows;
    }
    PBSO_CUDA(cudaMemcpyAsync(bt->d_ev_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, bt->stream));
    PBSO_CUDA(cudaMemcpyAsync(bt->d_ev_buf, sbuf.data(), sizeof(int) * sbuf.size(), cudaMemcpyHostToDevice, bt->stream));
    if (n_events > 0) {
        const size_t bytes = sizeof(double) * (size_t)n_events * bt->n_modes;
        if (identity) {
            PBSO_CUDA(cudaMemcpyAsync(bt->d_ev_space, space, bytes, cudaMemcpyHostToDevice, bt->stream));
        } else {
            if ((size_t)n_events > bt->ev_stage_cap) {
                cudaFree(bt->d_ev_stage); bt->d_ev_stage = nullptr; bt->ev_stage_cap = 0;
                PBSO_CUDA(cudaMalloc(&bt->d_ev_stage, bytes)); bt->ev_stage_cap = n_events;
            }
            PBSO_CUDA(cudaMemcpyAsync(bt->d_ev_stage, space, bytes, cudaMemcpyHostToDevice, bt->stream));
            PBSO_CUDA(cudaMemcpyAsync(bt->d_ev_src, src.data(), sizeof(int) * n_events, cudaMemcpyHostToDevice, bt->stream));
            k_gather_rows<<<n_events, 256, 0, bt->stream>>>(bt->n_modes, bt->d_ev_src, bt->d_ev_stage, bt->d_ev_space);
            PBSO_CUDA(cudaGetLastError());
        }
    }
    // off / sbuf / src are pageable: cudaMemcpyAsync has staged them when it returns; `space` may be pinned
    if (wait) PBSO_CUDA(cudaStreamSynchronize(bt->stream));          // the caller may reuse its buffers on return
    bt->h_ev_off = off; bt->h_ev_buf.assign(sbuf.begin(), sbuf.begin() + n_events); ++bt->ev_ver;
    bt->n_events = n_events;
    return PBSO_OK;
}

int pbso_batch_render_mix_device(pbso_batch* bt, int buf_size, int n_buffers, int precision, int n_chunks, double* d_mix) {
    PBSO_REQUIRE(bt && d_mix, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(bt->device);
    return launch_render(bt, buf_size, n_buffers, precision, n_chunks, d_mix, nullptr);
}

int pbso_batch_render_mix(pbso_batch* bt, int buf_size, int n_buffers, int precision, int n_chunks, double* mix) {
    PBSO_REQUIRE(bt && mix, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(bt->device);
    const size_t ns = (size_t)buf_size * n_buffers;
    if (ns > bt->mix_cap) { cudaFree(bt->d_mix); PBSO_CUDA(cudaMalloc(&bt->d_mix, sizeof(double) * ns)); bt->mix_cap = ns; }
    if (int rc = launch_render(bt, buf_size, n_buffers, precision, n_chunks, bt->d_mix, nullptr)) return rc;
    PBSO_CUDA(cudaMemcpyAsync(mix, bt->d_mix, sizeof(double) * ns, cudaMemcpyDeviceToHost, bt->stream));
    PBSO_CUDA(cudaStreamSynchronize(bt->stream));
    return PBSO_OK;
}

int pbso_batch_render_stems(pbso_batch* bt, int buf_size, int n_buffers, int precision, float* stems) {
    PBSO_REQUIRE(bt && stems, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(bt->device);
    const size_t ns = (size_t)buf_size * n_buffers * bt->n_obj;
    float* d_stems; PBSO_CUDA(cudaMalloc(&d_stems, sizeof(float) * ns));
    int rc = launch_render(bt, buf_size, n_buffers, precision, 0, nullptr, d_stems);
    if (rc == PBSO_OK) {
        cudaError_t e = cudaMemcpyAsync(stems, d_stems, sizeof(float) * ns, cudaMemcpyDeviceToHost, bt->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(bt->stream);
        if (e != cudaSuccess) rc = set_error(PBSO_ERR_CUDA, "stems copy failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d_stems);
    return rc;
}

int pbso_batch_set_state(pbso_batch* bt, const double* q_km1, const double* q_km2) {
    PBSO_REQUIRE(bt, PBSO_ERR_INVALID, "null handle");
    PBSO_REQUIRE((q_km1 == nullptr) == (q_km2 == nullptr), PBSO_ERR_INVALID, "q_km1 and q_km2 go together");
    DeviceGuard g(bt->device);
    if (!q_km1) { bt->has_state = false; return PBSO_OK; }          // back to the zero state of a fresh solver
    const size_t n = bt->npm();
    if (!bt->d_state) PBSO_CUDA(cudaMalloc(&bt->d_state, sizeof(double) * 4 * n));
    PBSO_CUDA(cudaMemcpyAsync(bt->d_state, q_km1, sizeof(double) * n, cudaMemcpyHostToDevice, bt->stream));
    PBSO_CUDA(cudaMemcpyAsync(bt->d_state + n, q_km2, sizeof(double) * n, cudaMemcpyHostToDevice, bt->stream));
    k_batch_state_to_carrier<<<(unsigned)((n + 255) / 256), 256, 0, bt->stream>>>(n, bt->lneps(), bt->theta(), bt->d_state, bt->d_state + n,
                                                                                   bt->d_state + 2 * n, bt->d_state + 3 * n);
    PBSO_CUDA(cudaGetLastError());
    PBSO_CUDA(cudaStreamSynchronize(bt->stream));                    // the caller may reuse its buffers on return
    bt->has_state = true;
    return PBSO_OK;
}

int pbso_batch_get_end_state(pbso_batch* bt, int buf_size, int n_buffers, double* q_km1, double* q_km2) {
    PBSO_REQUIRE(bt && q_km1 && q_km2 && buf_size > 0 && n_buffers > 0, PBSO_ERR_INVALID, "bad argument");
    DeviceGuard g(bt->device);
    const size_t n = bt->npm();
    if (!bt->d_ev_off) { if (int rc = pbso_batch_set_impulses(bt, 0, nullptr, nullptr, nullptr)) return rc; }
    double* d_q; PBSO_CUDA(cudaMalloc(&d_q, sizeof(double) * 2 * n));
    k_batch_end_state<<<(unsigned)((n + 127) / 128), 128, 0, bt->stream>>>(bt->n_obj, bt->n_modes, (long long)buf_size * n_buffers, buf_size,
        bt->lneps(), bt->theta(), bt->c3(), bt->cot(), bt->d_ev_off, bt->d_ev_buf, bt->d_ev_space, bt->v0r(), bt->v0i(), d_q, d_q + n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(q_km1, d_q, sizeof(double) * n, cudaMemcpyDeviceToHost, bt->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q_km2, d_q + n, sizeof(double) * n, cudaMemcpyDeviceToHost, bt->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(bt->stream);
    cudaFree(d_q);
    if (e != cudaSuccess) return set_error(PBSO_ERR_CUDA, "end state failed: %s", cudaGetErrorString(e));
    return PBSO_OK;
}

int pbso_batch_sync(pbso_batch* bt) {
    PBSO_REQUIRE(bt, PBSO_ERR_INVALID, "null handle");
    DeviceGuard g(bt->device);
    PBSO_CUDA(cudaStreamSynchronize(bt->stream));
    return PBSO_OK;
}

int pbso_batch_set_stream(pbso_batch* bt, void* cuda_stream) {
    PBSO_REQUIRE(bt, PBSO_ERR_INVALID, "null handle");
    DeviceGuard g(bt->device);
    PBSO_CUDA(cudaStreamSynchronize(bt->stream));
    bt->stream = cuda_stream ? (cudaStream_t)cuda_stream : bt->own_stream;
    return PBSO_OK;
}

int pbso_tc_gain(double* gain) {
    PBSO_REQUIRE(gain, PBSO_ERR_INVALID, "null output");
    int dev = 0; PBSO_CUDA(cudaGetDevice(&dev));
    *gain = tc_gain(dev);
    return PBSO_OK;
}

int pbso_batch_last_kernel_ms(pbso_batch* bt, float* ms, int* launches) {
    PBSO_REQUIRE(bt && ms, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(bt->device);
    PBSO_CUDA(cudaEventSynchronize(bt->e1));
    PBSO_CUDA(cudaEventElapsedTime(ms, bt->e0, bt->e1));
    if (launches) *launches = bt->last_launches;
    return PBSO_OK;
}

}  // extern "C"
