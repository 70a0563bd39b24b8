// =============================================================================
// modes.cu -- ModeData<double>::_modes on the device and the impulse projection U^T f.
//
//  K4s k_project_sparse   tools/real_time_modal_sound.cpp:236-295  GetModalForceVertex / Face
//                         (3 or 9 reads per mode; B impulses per launch)
//  K4  k_gemv_f64         same contraction with a dense load vector f (K = 3V): HBM-bound GEMV
//  K5  k_gemm_f64         dense batch Y = U F in FP64 (exact-parity path; the tensor-core 3xTF32
//                         kernel lives in project_tc.cu)
//
// HBM layout: U is mode-major [M][K] doubles exactly as ModeData stores it (ModeData.h:23-24,
// 61-83), so a mode's DOFs are contiguous: the GEMV streams rows with 16-byte loads, and the sparse
// gather touches one 32-byte sector per (mode, vertex).
// =============================================================================
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <fstream>
#include <vector>

using namespace pbso;

struct pbso_modes {
    int device = 0;
    int M = 0, K = 0;
    double* d_U = nullptr;
    std::vector<double> omega2;       // only when read from a .modes file
    cudaStream_t stream = nullptr;
    void* d_scratch = nullptr; size_t scratch_cap = 0;
    // tensor-core path (project_tc.cu): TF32 hi/lo planes of U, built on first use
    float* d_Uhi = nullptr; float* d_Ulo = nullptr;
    float* d_Fsplit = nullptr; size_t fsplit_cap = 0;      // F_hi | F_lo planes
    int sm_count = 148;
    cudaEvent_t e0 = nullptr, e1 = nullptr;                // bracket the last projection kernel(s)
    // contact-storm pipeline (pbso_modes_storm_buffer): dense load vectors, projected loads, summed load, delta profile, output
    void* d_storm = nullptr; size_t storm_cap = 0;
    double* h_storm = nullptr; size_t h_storm_cap = 0;     // pinned staging
};

namespace pbso {
int tc_pitch(int K);
int tc_split_f64(const double* d_src, int rows, int cols, float* d_hi, float* d_lo, cudaStream_t s);
int tc_split_f32(const float* d_src, int rows, int cols, float* d_hi, float* d_lo, cudaStream_t s);
int tc_project(const float* d_Uhi, const float* d_Ulo, int M, const float* d_Fhi, const float* d_Flo, int B, int K,
               float* d_Y, int sm_count, cudaStream_t s);
}

// out[b][m] = sum_{j<nv} coords[b][j] * (vn[b] . U_m[3 vid[b][j] .. +2])
// nv = 1, coords = 1 reproduces GetModalForceVertex (:276-280); nv = 3 GetModalForceFace (:245-251).
__global__ void __launch_bounds__(128)
k_project_sparse(int M, int K, int B, int nv, const double* __restrict__ U, const int* __restrict__ vids,
                 const double* __restrict__ coords, const double* __restrict__ vn, double* __restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double* mode = U + (size_t)m * K;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {      // gridDim.y is capped at 65535: stride over the batch
        const double n0 = vn[3 * b], n1 = vn[3 * b + 1], n2 = vn[3 * b + 2];
        double acc = 0.0;
        if (nv == 1) {
            const int v = vids[b];
            acc = n0 * mode[v * 3 + 0] + n1 * mode[v * 3 + 1] + n2 * mode[v * 3 + 2];
        } else {
            for (int j = 0; j < nv; ++j) {
                const int v = vids[(size_t)b * nv + j];
                const double cj = coords[(size_t)b * nv + j];
                acc += n0 * mode[v * 3 + 0] * cj + n1 * mode[v * 3 + 1] * cj + n2 * mode[v * 3 + 2] * cj;
            }
        }
        out[(size_t)b * M + m] = acc;
    }
}

// One block per mode row; 16-byte loads; FP64 accumulate; block reduce.
__global__ void __launch_bounds__(256)
k_gemv_f64(int K, const double* __restrict__ U, const double* __restrict__ f, double* __restrict__ y) {
    const int m = blockIdx.x;
    const double2* row = reinterpret_cast<const double2*>(U + (size_t)m * K);
    const double2* f2 = reinterpret_cast<const double2*>(f);
    const int K2 = K >> 1;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int i = threadIdx.x;
    for (; i + 768 < K2; i += 1024) {
        const double2 u0 = row[i], u1 = row[i + 256], u2 = row[i + 512], u3 = row[i + 768];
        const double2 g0 = f2[i], g1 = f2[i + 256], g2 = f2[i + 512], g3 = f2[i + 768];
        a0 = fma(u0.x, g0.x, a0); a0 = fma(u0.y, g0.y, a0);
        a1 = fma(u1.x, g1.x, a1); a1 = fma(u1.y, g1.y, a1);
        a2 = fma(u2.x, g2.x, a2); a2 = fma(u2.y, g2.y, a2);
        a3 = fma(u3.x, g3.x, a3); a3 = fma(u3.y, g3.y, a3);
    }
    for (; i < K2; i += 256) { const double2 u = row[i], g = f2[i]; a0 = fma(u.x, g.x, a0); a0 = fma(u.y, g.y, a0); }
    if ((K & 1) && threadIdx.x == 0) a1 = fma(U[(size_t)m * K + K - 1], f[K - 1], a1);
    double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        int lo = __shfl_xor_sync(0xffffffffu, __double2loint(acc), s);
        int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(acc), s);
        acc += __hiloint2double(hi, lo);
    }
    __shared__ double s_part[8];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_part[w];
        y[m] = s;
    }
}

// Y[B][M] = U[M][K] F[B][K]^T, FP64, 64x64 tile per block, 4x4 per thread.
__global__ void __launch_bounds__(256)
k_gemm_f64(int M, int K, int B, const double* __restrict__ U, const double* __restrict__ F, double* __restrict__ Y) {
    __shared__ double As[16][64 + 1];
    __shared__ double Bs[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int mm = e >> 4, kk = e & 15;                 // U row-contiguous in k
            As[kk][mm] = (m0 + mm < M && k0 + kk < K) ? U[(size_t)(m0 + mm) * K + k0 + kk] : 0.0;
            Bs[kk][mm] = (b0 + mm < B && k0 + kk < K) ? F[(size_t)(b0 + mm) * K + k0 + kk] : 0.0;   // F row-contiguous in k
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, b = b0 + tx * 4 + j;
            if (m < M && b < B) Y[(size_t)b * M + m] = acc[i][j];
        }
}

static int ensure_scratch(pbso_modes* md, size_t bytes) {
    if (bytes > md->scratch_cap) {
        cudaFree(md->d_scratch);
        PBSO_CUDA(cudaMalloc(&md->d_scratch, bytes));
        md->scratch_cap = bytes;
    }
    return PBSO_OK;
}

static int project_sparse(pbso_modes* md, int force_dim, int B, int nv, const int* vids, const double* coords,
                          const double* vn, double* out) {
    PBSO_REQUIRE(md && vids && vn && out && B > 0, PBSO_ERR_INVALID, "bad argument");
    PBSO_REQUIRE(force_dim >= 0 && force_dim <= md->M, PBSO_ERR_RANGE, "forceDim exceeds number of modes (modes.mode(mm) .at())");
    for (int i = 0; i < B * nv; ++i)
        if (vids[i] < 0 || (long long)vids[i] * 3 + 2 >= md->K)
            return set_error(PBSO_ERR_RANGE, "vertex id %d outside the %d DOFs (std::vector::at)", vids[i], md->K);
    if (force_dim == 0) return PBSO_OK;
    DeviceGuard g(md->device);
    const size_t nb_v = sizeof(int) * B * nv, nb_c = sizeof(double) * B * nv, nb_n = sizeof(double) * 3 * B,
                 nb_o = sizeof(double) * (size_t)B * force_dim;
    const size_t off_c = (nb_v + 15) & ~(size_t)15, off_n = off_c + ((nb_c + 15) & ~(size_t)15),
                 off_o = off_n + ((nb_n + 15) & ~(size_t)15);
    if (int rc = ensure_scratch(md, off_o + nb_o)) return rc;
    char* s = (char*)md->d_scratch;
    PBSO_CUDA(cudaMemcpyAsync(s, vids, nb_v, cudaMemcpyHostToDevice, md->stream));
    if (coords) PBSO_CUDA(cudaMemcpyAsync(s + off_c, coords, nb_c, cudaMemcpyHostToDevice, md->stream));
    PBSO_CUDA(cudaMemcpyAsync(s + off_n, vn, nb_n, cudaMemcpyHostToDevice, md->stream));
    k_project_sparse<<<dim3(div_up(force_dim, 128), std::min(B, 65535)), 128, 0, md->stream>>>(
        force_dim, md->K, B, nv, md->d_U, (const int*)s, (const double*)(s + off_c), (const double*)(s + off_n),
        (double*)(s + off_o));
    PBSO_CUDA(cudaGetLastError());
    PBSO_CUDA(cudaMemcpyAsync(out, s + off_o, nb_o, cudaMemcpyDeviceToHost, md->stream));
    PBSO_CUDA(cudaStreamSynchronize(md->stream));
    return PBSO_OK;
}

extern "C" {

int pbso_modes_upload(const double* U, int M, int K, pbso_modes** out) {
    PBSO_REQUIRE(out, PBSO_ERR_INVALID, "null output handle");
    *out = nullptr;
    PBSO_REQUIRE(U && M > 0 && K > 0, PBSO_ERR_INVALID, "bad argument");
    if (int rc = check_device()) return rc;
    pbso_modes* md = new pbso_modes();
    md->M = M; md->K = K;
    PBSO_CUDA(cudaGetDevice(&md->device));
    PBSO_CUDA(cudaDeviceGetAttribute(&md->sm_count, cudaDevAttrMultiProcessorCount, md->device));
    PBSO_CUDA(cudaStreamCreateWithFlags(&md->stream, cudaStreamNonBlocking));
    PBSO_CUDA(cudaEventCreate(&md->e0)); PBSO_CUDA(cudaEventCreate(&md->e1));
    PBSO_CUDA(cudaMalloc(&md->d_U, sizeof(double) * (size_t)M * K));
    PBSO_CUDA(cudaMemcpy(md->d_U, U, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice));
    *out = md;
    return PBSO_OK;
}

int pbso_modes_read_file(const char* filename, pbso_modes** out, int* M, int* K) {
    PBSO_REQUIRE(out && filename, PBSO_ERR_INVALID, "null argument");
    *out = nullptr;
    std::ifstream fin(filename, std::ios::binary);
    if (!fin.good()) return set_error(PBSO_ERR_IO, "cannot open file for reading modes: %s", filename);
    int nDOF = 0, nModes = 0;                                    // ModeData.h:66-68
    fin.read((char*)&nDOF, sizeof(int));
    fin.read((char*)&nModes, sizeof(int));
    if (!fin.good() || nDOF <= 0 || nModes <= 0) return set_error(PBSO_ERR_FORMAT, "%s: bad header", filename);
    std::vector<double> w2(nModes), U((size_t)nModes * nDOF);
    fin.read((char*)w2.data(), sizeof(double) * nModes);         // :71-72
    fin.read((char*)U.data(), sizeof(double) * (size_t)nModes * nDOF);   // :75-79 (rows are contiguous)
    if (!fin.good()) return set_error(PBSO_ERR_FORMAT, "%s: truncated", filename);
    if (int rc = pbso_modes_upload(U.data(), nModes, nDOF, out)) return rc;
    (*out)->omega2 = std::move(w2);
    if (M) *M = nModes;
    if (K) *K = nDOF;
    return PBSO_OK;
}

int pbso_modes_omega_squared(const pbso_modes* md, double* omega_squared) {
    PBSO_REQUIRE(md && omega_squared, PBSO_ERR_INVALID, "null argument");
    PBSO_REQUIRE(!md->omega2.empty(), PBSO_ERR_UNSUPPORTED, "handle was not read from a .modes file");
    std::memcpy(omega_squared, md->omega2.data(), sizeof(double) * md->omega2.size());
    return PBSO_OK;
}

int pbso_modes_destroy(pbso_modes* md) {
    if (!md) return PBSO_OK;
    DeviceGuard g(md->device);
    if (md->stream) cudaStreamSynchronize(md->stream);
    cudaFree(md->d_U); cudaFree(md->d_scratch); cudaFree(md->d_Uhi); cudaFree(md->d_Ulo); cudaFree(md->d_Fsplit); cudaFree(md->d_storm);
    if (md->h_storm) cudaFreeHost(md->h_storm);
    if (md->e0) cudaEventDestroy(md->e0);
    if (md->e1) cudaEventDestroy(md->e1);
    if (md->stream) cudaStreamDestroy(md->stream);
    delete md;
    return PBSO_OK;
}

int pbso_modes_project_vertex(const pbso_modes* md, int force_dim, int vid, const double* vn3, double* out) {
    return project_sparse(const_cast<pbso_modes*>(md), force_dim, 1, 1, &vid, nullptr, vn3, out);
}

int pbso_modes_project_face(const pbso_modes* md, int force_dim, const int* vids3, const double* coords3,
                            const double* vn3, double* out) {
    PBSO_REQUIRE(coords3, PBSO_ERR_INVALID, "null coords");
    return project_sparse(const_cast<pbso_modes*>(md), force_dim, 1, 3, vids3, coords3, vn3, out);
}

int pbso_modes_project_vertices(const pbso_modes* md, int force_dim, int B, const int* vids, const double* vn,
                                double* out) {
    return project_sparse(const_cast<pbso_modes*>(md), force_dim, B, 1, vids, nullptr, vn, out);
}

static int ensure_u_planes(pbso_modes* md) {
    if (md->d_Uhi) return PBSO_OK;
    const size_t n = (size_t)md->M * tc_pitch(md->K);
    PBSO_CUDA(cudaMalloc(&md->d_Uhi, sizeof(float) * n));
    PBSO_CUDA(cudaMalloc(&md->d_Ulo, sizeof(float) * n));
    return tc_split_f64(md->d_U, md->M, md->K, md->d_Uhi, md->d_Ulo, md->stream);
}
static int ensure_f_planes(pbso_modes* md, int B) {
    const size_t n = 2 * (size_t)B * tc_pitch(md->K);
    if (n > md->fsplit_cap) {
        cudaFree(md->d_Fsplit);
        PBSO_CUDA(cudaMalloc(&md->d_Fsplit, sizeof(float) * n));
        md->fsplit_cap = n;
    }
    return PBSO_OK;
}

int pbso_modes_project_dense(const pbso_modes* mdc, int force_dim, const double* F, int B, double* Y, int precision) {
    pbso_modes* md = const_cast<pbso_modes*>(mdc);
    PBSO_REQUIRE(md && F && Y && B > 0, PBSO_ERR_INVALID, "bad argument");
    PBSO_REQUIRE(force_dim > 0 && force_dim <= md->M, PBSO_ERR_RANGE, "forceDim exceeds number of modes");
    DeviceGuard g(md->device);
    const int K = md->K;
    const size_t nb_f = sizeof(double) * (size_t)K * B, nb_y = sizeof(double) * (size_t)force_dim * B;
    const size_t off_y = (nb_f + 255) & ~(size_t)255;
    if (int rc = ensure_scratch(md, off_y + nb_y)) return rc;
    char* s = (char*)md->d_scratch;
    PBSO_CUDA(cudaMemcpyAsync(s, F, nb_f, cudaMemcpyHostToDevice, md->stream));
    PBSO_CUDA(cudaEventRecord(md->e0, md->stream));
    if (precision == PBSO_PREC_F64) {
        if (B == 1 && (K % 2 == 0)) {
            k_gemv_f64<<<force_dim, 256, 0, md->stream>>>(K, md->d_U, (const double*)s, (double*)(s + off_y));
        } else {
            k_gemm_f64<<<dim3(div_up(B, 64), div_up(force_dim, 64)), 256, 0, md->stream>>>(
                force_dim, K, B, md->d_U, (const double*)s, (double*)(s + off_y));
        }
        PBSO_CUDA(cudaGetLastError());
        PBSO_CUDA(cudaEventRecord(md->e1, md->stream));
        PBSO_CUDA(cudaMemcpyAsync(Y, s + off_y, nb_y, cudaMemcpyDeviceToHost, md->stream));
        PBSO_CUDA(cudaStreamSynchronize(md->stream));
        return PBSO_OK;
    }
    PBSO_REQUIRE(precision == PBSO_PREC_TF32X3, PBSO_ERR_INVALID, "precision must be PBSO_PREC_F64 or PBSO_PREC_TF32X3");
    if (int rc = ensure_u_planes(md)) return rc;
    if (int rc = ensure_f_planes(md, B)) return rc;
    const size_t plane = (size_t)B * tc_pitch(K);
    if (int rc = tc_split_f64((const double*)s, B, K, md->d_Fsplit, md->d_Fsplit + plane, md->stream)) return rc;
    float* d_Y = (float*)(s + off_y);
    if (int rc = tc_project(md->d_Uhi, md->d_Ulo, force_dim, md->d_Fsplit, md->d_Fsplit + plane, B, K, d_Y, md->sm_count, md->stream)) return rc;
    PBSO_CUDA(cudaEventRecord(md->e1, md->stream));
    std::vector<float> yf((size_t)force_dim * B);
    PBSO_CUDA(cudaMemcpyAsync(yf.data(), d_Y, sizeof(float) * yf.size(), cudaMemcpyDeviceToHost, md->stream));
    PBSO_CUDA(cudaStreamSynchronize(md->stream));
    for (size_t i = 0; i < yf.size(); ++i) Y[i] = (double)yf[i];
    return PBSO_OK;
}

int pbso_modes_last_kernel_ms(const pbso_modes* md, float* ms) {
    PBSO_REQUIRE(md && ms, PBSO_ERR_INVALID, "null argument");
    DeviceGuard g(md->device);
    PBSO_CUDA(cudaEventSynchronize(md->e1));
    PBSO_CUDA(cudaEventElapsedTime(ms, md->e0, md->e1));
    return PBSO_OK;
}

int pbso_modes_project_dense_device(const pbso_modes* mdc, int force_dim, const float* d_F, int B, float* d_Y, void* cuda_stream) {
    pbso_modes* md = const_cast<pbso_modes*>(mdc);
    PBSO_REQUIRE(md && d_F && d_Y && B > 0, PBSO_ERR_INVALID, "bad argument");
    PBSO_REQUIRE(force_dim > 0 && force_dim <= md->M, PBSO_ERR_RANGE, "forceDim exceeds number of modes");
    DeviceGuard g(md->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : md->stream;
    if (!md->d_Uhi) { if (int rc = ensure_u_planes(md)) return rc; PBSO_CUDA(cudaStreamSynchronize(md->stream)); }
    if (int rc = ensure_f_planes(md, B)) return rc;
    const size_t plane = (size_t)B * tc_pitch(md->K);
    if (int rc = tc_split_f32(d_F, B, md->K, md->d_Fsplit, md->d_Fsplit + plane, st)) return rc;
    return tc_project(md->d_Uhi, md->d_Ulo, force_dim, md->d_Fsplit, md->d_Fsplit + plane, B, md->K, d_Y, md->sm_count, st);
}

// ---- cfg3: contact storm -> projection -> rank-1 load -> integrator, one buffer, all on the integrator's stream ----------
// F[b][3 vid_b + d] = vn_b[d]: the dense load vector of a vertex impulse (what GetModalForceVertex contracts with U)
}  // extern "C" (kernels below)
__global__ void k_storm_scatter(int B, int K, const int* __restrict__ vids, const double* __restrict__ vn, float* __restrict__ F) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float* row = F + (size_t)b * K + 3 * (size_t)vids[b];
    row[0] = (float)vn[3 * b]; row[1] = (float)vn[3 * b + 1]; row[2] = (float)vn[3 * b + 2];
}
// space[m] = sum_b Y[b][m]  (ModalSolver::step sums the spatial loads of the active forces, modal_solver.h:206-221)
template <typename T>
__global__ void k_storm_sum(int B, int M, const T* __restrict__ Y, double* __restrict__ space) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int b = 0;
    for (; b + 3 < B; b += 4) {
        a0 += (double)Y[(size_t)b * M + m]; a1 += (double)Y[(size_t)(b + 1) * M + m];
        a2 += (double)Y[(size_t)(b + 2) * M + m]; a3 += (double)Y[(size_t)(b + 3) * M + m];
    }
    for (; b < B; ++b) a0 += (double)Y[(size_t)b * M + m];
    space[m] = (a0 + a1) + (a2 + a3);
}
extern "C" {

int pbso_modes_storm_buffer(pbso_modes* md, pbso_integrator* it, int force_dim, int B, const int* vids, const double* vn,
                            int T, double* y_out, double* qnorm_out, int precision) {
    PBSO_REQUIRE(md && it && vids && vn && y_out && B > 0 && T > 0, PBSO_ERR_INVALID, "bad argument");
    PBSO_REQUIRE(precision == PBSO_PREC_F64 || precision == PBSO_PREC_TF32X3, PBSO_ERR_INVALID, "precision must be PBSO_PREC_F64 or PBSO_PREC_TF32X3");
    int N = 0; if (int rc = pbso_integrator_size(it, &N)) return rc;
    PBSO_REQUIRE(force_dim == N && force_dim <= md->M, PBSO_ERR_RANGE, "forceDim must equal the integrator's mode count and not exceed the mode shapes");
    for (int i = 0; i < B; ++i)
        if (vids[i] < 0 || (long long)vids[i] * 3 + 2 >= md->K)
            return set_error(PBSO_ERR_RANGE, "vertex id %d outside the %d DOFs (std::vector::at)", vids[i], md->K);
    DeviceGuard g(md->device);
    void* sv = nullptr; if (int rc = pbso_integrator_stream(it, &sv)) return rc;
    cudaStream_t st = (cudaStream_t)sv;
    const int K = md->K, M = force_dim;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const bool dense = precision == PBSO_PREC_TF32X3;
    const size_t o_vid = 0, o_vn = al(sizeof(int) * B), o_F = o_vn + al(sizeof(double) * 3 * B);
    const size_t o_Y = o_F + (dense ? al(sizeof(float) * (size_t)B * K) : 0);
    const size_t o_sp = o_Y + al((dense ? sizeof(float) : sizeof(double)) * (size_t)B * M);
    const size_t o_tm = o_sp + al(sizeof(double) * M), o_y = o_tm + al(sizeof(double) * T);
    const size_t o_q = o_y + al(sizeof(double) * 64 * (size_t)T), total = o_q + al(sizeof(double) * M);
    if (total > md->storm_cap) {
        PBSO_CUDA(cudaStreamSynchronize(st));
        cudaFree(md->d_storm); md->d_storm = nullptr; md->storm_cap = 0;
        PBSO_CUDA(cudaMalloc(&md->d_storm, total)); md->storm_cap = total;
    }
    const size_t h_need = (size_t)4 * B + 64 * (size_t)T + M + T;
    if (h_need > md->h_storm_cap) {
        if (md->h_storm) cudaFreeHost(md->h_storm);
        md->h_storm = nullptr; md->h_storm_cap = 0;
        PBSO_CUDA(cudaMallocHost(&md->h_storm, sizeof(double) * h_need)); md->h_storm_cap = h_need;
    }
    char* d = (char*)md->d_storm;
    // pinned staging: vids | vn | delta profile in, y | qnorm out
    int* h_vid = (int*)md->h_storm; double* h_vn = md->h_storm + (B + 1) / 2; double* h_tm = h_vn + 3 * B; double* h_y = h_tm + T; double* h_q = h_y + 64 * (size_t)T;
    std::memcpy(h_vid, vids, sizeof(int) * B); std::memcpy(h_vn, vn, sizeof(double) * 3 * B);
    std::memset(h_tm, 0, sizeof(double) * T); h_tm[0] = 1.0;                     // PointForce::Add (forces.h:87)
    PBSO_CUDA(cudaMemcpyAsync(d + o_vid, h_vid, sizeof(int) * B, cudaMemcpyHostToDevice, st));
    PBSO_CUDA(cudaMemcpyAsync(d + o_vn, h_vn, sizeof(double) * 3 * B, cudaMemcpyHostToDevice, st));
    PBSO_CUDA(cudaMemcpyAsync(d + o_tm, h_tm, sizeof(double) * T, cudaMemcpyHostToDevice, st));
    PBSO_CUDA(cudaEventRecord(md->e0, st));
    if (dense) {
        PBSO_CUDA(cudaMemsetAsync(d + o_F, 0, sizeof(float) * (size_t)B * K, st));
        k_storm_scatter<<<div_up(B, 128), 128, 0, st>>>(B, K, (const int*)(d + o_vid), (const double*)(d + o_vn), (float*)(d + o_F));
        PBSO_CUDA(cudaGetLastError());
        if (int rc = pbso_modes_project_dense_device(md, M, (const float*)(d + o_F), B, (float*)(d + o_Y), (void*)st)) return rc;
        k_storm_sum<float><<<div_up(M, 128), 128, 0, st>>>(B, M, (const float*)(d + o_Y), (double*)(d + o_sp));
    } else {
        k_project_sparse<<<dim3(div_up(M, 128), std::min(B, 65535)), 128, 0, st>>>(M, K, B, 1, md->d_U, (const int*)(d + o_vid), nullptr,
                                                                               (const double*)(d + o_vn), (double*)(d + o_Y));
        PBSO_CUDA(cudaGetLastError());
        k_storm_sum<double><<<div_up(M, 128), 128, 0, st>>>(B, M, (const double*)(d + o_Y), (double*)(d + o_sp));
    }
    PBSO_CUDA(cudaGetLastError());
    PBSO_CUDA(cudaEventRecord(md->e1, st));
    if (int rc = pbso_render_buffer_device(it, (const double*)(d + o_sp), (const double*)(d + o_tm), T, (double*)(d + o_y), qnorm_out ? (double*)(d + o_q) : nullptr)) return rc;
    int L = 1; pbso_integrator_listeners(it, &L);
    PBSO_REQUIRE(L >= 0 && L <= 64, PBSO_ERR_UNSUPPORTED, "the storm entry stages at most 64 listeners");
    if (L > 0) PBSO_CUDA(cudaMemcpyAsync(h_y, d + o_y, sizeof(double) * (size_t)L * T, cudaMemcpyDeviceToHost, st));
    if (qnorm_out) PBSO_CUDA(cudaMemcpyAsync(h_q, d + o_q, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
    PBSO_CUDA(cudaStreamSynchronize(st));
    if (L > 0) std::memcpy(y_out, h_y, sizeof(double) * (size_t)L * T);
    if (qnorm_out) std::memcpy(qnorm_out, h_q, sizeof(double) * M);
    return PBSO_OK;
}

}  // extern "C"
