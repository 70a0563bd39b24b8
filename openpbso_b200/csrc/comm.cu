// =============================================================================
// comm.cu -- the one exchange step of the synthesis path: sound objects (and the mode blocks of one
// large object) are independent, so every rank renders its block into its own FP64 mix and ONE sum-reduce of the
// audio lands the track on the root.  NCCL is called from here, on the stream the render kernels run on, so that a
// C++ caller (tools/pbso_render --gpus N) and bench.py share the product's own multi-GPU path.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): single-GPU users need no NCCL at all, and inside a process
// that already loaded an NCCL under that soname (PyTorch ships one) the same library is reused.
// =============================================================================
#include "common.cuh"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

using namespace pbso;

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 opaque bytes passed by value)
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8 };
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_destroy)(nccl_comm);
typedef int (*fn_reduce)(const void*, void*, size_t, int, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_version)(int*);

struct Nccl {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr; fn_init_rank init_rank = nullptr; fn_destroy destroy = nullptr;
    fn_reduce reduce = nullptr; fn_allreduce allreduce = nullptr; fn_errstr errstr = nullptr; fn_version version = nullptr;
};
Nccl g_nccl;
std::once_flag g_once;

int load_nccl() {
    std::call_once(g_once, [] {
        const char* names[] = {getenv("PBSO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (!g_nccl.lib) return;
        g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
        g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.lib, "ncclCommInitRank");
        g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
        g_nccl.reduce = (fn_reduce)dlsym(g_nccl.lib, "ncclReduce");
        g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
        g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
        g_nccl.version = (fn_version)dlsym(g_nccl.lib, "ncclGetVersion");
    });
    if (!g_nccl.lib || !g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.reduce || !g_nccl.allreduce)
        return set_error(PBSO_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    return PBSO_OK;
}

#define PBSO_NCCL(expr)                                                                         \
    do {                                                                                        \
        int _r = (expr);                                                                        \
        if (_r != 0) return set_error(PBSO_ERR_CUDA, "%s failed: %s", #expr, g_nccl.errstr ? g_nccl.errstr(_r) : "nccl error"); \
    } while (0)

}  // namespace

struct pbso_comm {
    int nranks = 1, rank = 0, device = 0;
    nccl_comm comm = nullptr;
};

extern "C" {

int pbso_comm_unique_id(unsigned char* id128) {
    PBSO_REQUIRE(id128, PBSO_ERR_INVALID, "null id");
    if (int rc = load_nccl()) return rc;
    nccl_uid u;
    PBSO_NCCL(g_nccl.get_uid(&u));
    std::memcpy(id128, u.internal, 128);
    return PBSO_OK;
}

int pbso_comm_init(int nranks, int rank, const unsigned char* id128, pbso_comm** out) {
    PBSO_REQUIRE(out, PBSO_ERR_INVALID, "null output handle");
    *out = nullptr;
    PBSO_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, PBSO_ERR_INVALID, "rank outside [0, nranks)");
    if (int rc = check_device()) return rc;
    pbso_comm* c = new pbso_comm();
    c->nranks = nranks; c->rank = rank;
    PBSO_CUDA(cudaGetDevice(&c->device));
    if (nranks > 1) {
        PBSO_REQUIRE(id128, PBSO_ERR_INVALID, "null id");
        if (int rc = load_nccl()) { delete c; return rc; }
        nccl_uid u; std::memcpy(u.internal, id128, 128);
        int r = g_nccl.init_rank(&c->comm, nranks, u, rank);
        if (r != 0) { delete c; return set_error(PBSO_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.errstr ? g_nccl.errstr(r) : "nccl error"); }
    }
    *out = c;
    return PBSO_OK;
}

int pbso_comm_destroy(pbso_comm* c) {
    if (!c) return PBSO_OK;
    if (c->comm) { DeviceGuard g(c->device); g_nccl.destroy(c->comm); }
    delete c;
    return PBSO_OK;
}

int pbso_comm_info(const pbso_comm* c, int* nranks, int* rank, int* nccl_version) {
    PBSO_REQUIRE(c, PBSO_ERR_INVALID, "null handle");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (nccl_version) { *nccl_version = 0; if (g_nccl.version) g_nccl.version(nccl_version); }
    return PBSO_OK;
}

int pbso_comm_shard(const pbso_comm* c, long long n_units, long long* lo, long long* hi) {
    PBSO_REQUIRE(c && lo && hi && n_units >= 0, PBSO_ERR_INVALID, "bad argument");
    const long long per = (n_units + c->nranks - 1) / c->nranks;
    *lo = std::min(c->rank * per, n_units);
    *hi = std::min(*lo + per, n_units);
    return PBSO_OK;
}

int pbso_comm_reduce_audio(pbso_comm* c, double* d_audio, size_t n, int root, void* cuda_stream) {
    PBSO_REQUIRE(c && d_audio, PBSO_ERR_INVALID, "null argument");
    PBSO_REQUIRE(root >= -1 && root < c->nranks, PBSO_ERR_INVALID, "root outside [-1, nranks)");
    if (c->nranks == 1 || n == 0) return PBSO_OK;                     // nothing to exchange
    DeviceGuard g(c->device);
    if (root < 0) PBSO_NCCL(g_nccl.allreduce(d_audio, d_audio, n, NCCL_FLOAT64, NCCL_SUM, c->comm, (cudaStream_t)cuda_stream));
    else PBSO_NCCL(g_nccl.reduce(d_audio, d_audio, n, NCCL_FLOAT64, NCCL_SUM, root, c->comm, (cudaStream_t)cuda_stream));
    return PBSO_OK;
}

int pbso_comm_reduce_audio_host(pbso_comm* c, double* audio, size_t n, int root) {
    PBSO_REQUIRE(c && audio, PBSO_ERR_INVALID, "null argument");
    PBSO_REQUIRE(root >= -1 && root < c->nranks, PBSO_ERR_INVALID, "root outside [-1, nranks)");
    if (c->nranks == 1 || n == 0) return PBSO_OK;
    DeviceGuard g(c->device);
    double* d = nullptr;
    PBSO_CUDA(cudaMalloc(&d, sizeof(double) * n));
    cudaError_t e = cudaMemcpy(d, audio, sizeof(double) * n, cudaMemcpyHostToDevice);
    int rc = PBSO_OK;
    if (e == cudaSuccess) {
        rc = pbso_comm_reduce_audio(c, d, n, root, nullptr);           // the default stream: ordered with the copies around it
        if (rc == PBSO_OK) e = cudaStreamSynchronize(nullptr);
        if (rc == PBSO_OK && e == cudaSuccess && (root < 0 || root == c->rank)) e = cudaMemcpy(audio, d, sizeof(double) * n, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (rc != PBSO_OK) return rc;
    if (e != cudaSuccess) return set_error(PBSO_ERR_CUDA, "audio reduce staging failed: %s", cudaGetErrorString(e));
    return PBSO_OK;
}

}  // extern "C"
