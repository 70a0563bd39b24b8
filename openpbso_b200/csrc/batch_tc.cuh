// Interface between batch.cu (handle, impulse script, C ABI) and batch_tc.cu (tensor-core render).
#pragma once
#include <cuda_runtime.h>

namespace pbso {

struct TcState;      // device tables / unit list cached across renders, owned by a pbso_batch

struct TcArgs {
    int n_obj, n_modes, buf_size, n_buffers, sm_count;
    const double *lneps, *theta, *c1, *c2, *c3, *cot, *trans;  // [n_obj][n_modes], device
    const int *h_ev_off, *h_ev_buf;                            // impulse CSR, host copy
    const int *d_ev_off, *d_ev_buf; const double* d_ev_space;  // impulse CSR, device
    int n_events;
    unsigned trans_ver, ev_ver;                                // bumped by set_transfer / set_impulses
    const double *v0r, *v0i;                                   // carrier at sample 0 (stateful range render), or NULL
    double* d_mix;                                             // zeroed by the caller (or NULL)
    float* d_stems;                                            // [n_obj][n_buffers*buf_size], zeroed by the caller (or NULL)
    cudaStream_t stream;
};

int tc_render(TcState** st, const TcArgs& a, int* launches);
// batch.cu: FP64 direct-form samples of events [e0, e0 + ne) from their impulse to the next 128-sample boundary
int batch_event_heads(const TcArgs& a, int e0, int ne, const int* d_ev_obj);
void tc_free(TcState* st);
double tc_gain(int device);   // calibrated accumulate-truncation gain of the tensor-core path (0 = not calibrated yet)

}  // namespace pbso
