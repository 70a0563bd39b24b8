/* =============================================================================
 * pbso_b200.h -- C ABI of libpbso_b200.so: the B200 (sm_100a) implementation of openpbso's
 * modal-synthesis hot path  (U^T f projection -> per-mode IIR -> FFAT-weighted modal sum).
 *
 * The reference (jhwang7628/openpbso) has no FFI layer: its API is a set of header-only C++
 * templates.  include/openpbso/ mirrors those headers and forwards to the entry points
 * below; each entry point cites the reference interface it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; host pointers unless a name ends in _device / d_*
 *  - every function returns an int status (PBSO_OK == 0); never throws, never spins
 *  - handles are opaque; each owns one CUDA stream and its device/pinned buffers; a handle is
 *    bound to the device that was current (pbso_set_device) when it was created
 *  - all arithmetic visible at this boundary is IEEE double, like the reference
 *    (ModalSolver<double>, tools/real_time_modal_sound.cpp:188); the offline batch renderer
 *    additionally offers an FP32-tile fast path that stays within the parity tolerance
 *  - there is NO CPU fallback: without a CUDA device every compute entry returns
 *    PBSO_ERR_NO_DEVICE
 * ============================================================================= */
#ifndef PBSO_B200_H
#define PBSO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBSO_ABI_VERSION 1

enum {
    PBSO_OK = 0,
    PBSO_ERR_INVALID = 1,     /* bad argument: null pointer, size mismatch (reference: assert) */
    PBSO_ERR_CUDA = 2,        /* CUDA runtime failure; text in pbso_last_error() */
    PBSO_ERR_IO = 3,          /* file / directory missing or unreadable */
    PBSO_ERR_FORMAT = 4,      /* malformed .fatcube / .modes contents */
    PBSO_ERR_RANGE = 5,       /* index or mode id out of range (reference: std::out_of_range) */
    PBSO_ERR_NO_DEVICE = 6,   /* no CUDA device: the product has no CPU path */
    PBSO_ERR_UNSUPPORTED = 7
};

/* arithmetic selectors */
#define PBSO_PREC_F64 0        /* FP64: the reference's arithmetic */
#define PBSO_PREC_F32_TILED 1  /* batch renderer: FP32 pole-power tiles evaluated from an FP64 carrier */
#define PBSO_PREC_TF32X3 2     /* batched projection: tcgen05 tensor cores, 3xTF32 split, FP32 accumulate */
#define PBSO_PREC_TC3X 3       /* batch renderer: pole-power synthesis as one tcgen05 3xTF32 contraction (any buf_size, 513 included) */

typedef struct pbso_integrator pbso_integrator;  /* ModalIntegrator<double> + per-buffer renderer */
typedef struct pbso_ffat pbso_ffat;              /* std::map<int, FFAT_Map<double,3>> */
typedef struct pbso_modes pbso_modes;            /* ModeData<double>::_modes on the device */
typedef struct pbso_batch pbso_batch;            /* many independent sound objects, offline */
typedef struct pbso_comm pbso_comm;              /* the ranks of a multi-GPU render (one NCCL communicator) */

/* ---- library / device ----------------------------------------------------------------- */
int pbso_abi_version(void);
const char* pbso_last_error(void);               /* thread-local, valid until the next call */
int pbso_device_count(int* n);
int pbso_set_device(int device);                 /* like cudaSetDevice, per calling thread */
int pbso_device_info(int* sm_count, int* cc_major, int* cc_minor, double* hbm_gib);

/* ---- modal integrator: modal_integrator.h --------------------------------------------- */
/* ModalIntegrator<T>::Build (modal_integrator.h:39-42, 47-70).  N < 0 means n_omega. */
int pbso_integrator_build(double density, const double* omega_squared, int n_omega,
                          double alpha, double beta, double h, int N, pbso_integrator** out);
/* ModalIntegrator<T>::ModalIntegrator(N, h, a, b) (modal_integrator.h:37-38, 72-101):
 * coefficient kernel K2 runs in FP64 on the device. */
int pbso_integrator_create(int N, double h, const double* a, const double* b,
                           pbso_integrator** out);
int pbso_integrator_destroy(pbso_integrator* it);
int pbso_integrator_size(const pbso_integrator* it, int* N);
/* Listeners of the resident transfer table (the L of the last set_transfer call), and the handle's CUDA stream
 * (cudaStream_t as void*): device-resident callers order their own work with the renders through it. */
int pbso_integrator_listeners(const pbso_integrator* it, int* L);
int pbso_integrator_stream(const pbso_integrator* it, void** cuda_stream);
/* _c1/_c2/_c3 (modal_integrator.h:32-34, 95-99); any pointer may be NULL. */
int pbso_integrator_get_coeffs(const pbso_integrator* it, double* c1, double* c2, double* c3);
/* Step(Q) / Step() (modal_integrator.h:43-44, 103-123); Q == NULL selects Step().
 * q_out[N] receives q_k (the reference returns a reference into its ring). */
int pbso_integrator_step(pbso_integrator* it, const double* Q, double* q_out);
/* Ring contents as (q_{k-1}, q_{k-2}) (modal_integrator.h:24,29): checkpoint / chunked render. */
int pbso_integrator_get_state(const pbso_integrator* it, double* q_km1, double* q_km2);
int pbso_integrator_set_state(pbso_integrator* it, const double* q_km1, const double* q_km2);

/* ---- per-buffer synthesis: ModalSolver::step hot loop (modal_solver.h:261-272) ---------
 * For i in [0,T):  q = Step(space * time[i]);  y[l][i] = sum_{m < n_transfer} q[m]*transfer[l][m];
 * qnorm[m] = sqrt(sum_i q[m]^2).  Kernel K1 (FP64 direct form, state kept in registers).
 *   space[N], time[T]; y_out[L*T] (listener-major); qnorm_out[N] or NULL.
 * The transfer table is resident: set it with pbso_integrator_set_transfer (TransMessage,
 * modal_solver.h:84-98, 250-252); transfer is column-major n_transfer x L like the reference's
 * batched call site (tools/real_time_modal_sound.cpp:921-927).  L == 0 renders q only. */
int pbso_integrator_set_transfer(pbso_integrator* it, const double* transfer, int n_transfer,
                                 int L);
int pbso_render_buffer(pbso_integrator* it, const double* space, const double* time, int T,
                       double* y_out, double* qnorm_out);
/* Moving listeners: computeTransfer(pos, T*) for L positions (modal_solver.h:303-315; batched call
 * site tools/real_time_modal_sound.cpp:921-927) evaluated by kernel K3 straight into the resident transfer table on
 * the integrator's stream -- the n_transfer x L table never visits the host.  pos is L x 3; maps must hold mode ids
 * 0..n_transfer-1 (PBSO_ERR_RANGE otherwise, like _ffat_maps->at()).  Takes effect for the next render call. */
int pbso_integrator_set_transfer_ffat(pbso_integrator* it, const pbso_ffat* maps, int n_transfer,
                                      const double* pos, int L);
/* Same, enqueue-only on the handle's stream with device-resident inputs/outputs (no copies,
 * no sync).  d_y must hold L*T doubles. */
int pbso_render_buffer_device(pbso_integrator* it, const double* d_space, const double* d_time,
                              int T, double* d_y, double* d_qnorm);
int pbso_integrator_sync(pbso_integrator* it);

/* ---- FFAT maps: ffat_solver.h runtime subset + ffat_map_serialize.h ------------------- */
/* FFAT_Map_Serialize_Double::LoadAll (ffat_map_serialize.h:267-279) incl. ListDirFiles(dir,
 * ".fatcube") (io.cpp:18-35).  A missing directory yields an empty set and PBSO_ERR_IO. */
int pbso_ffat_load_dir(const char* dirname, pbso_ffat** out);
/* FFAT_Map_Serialize_Double::Load (ffat_map_serialize.h:166-254) of a single file. */
int pbso_ffat_load_file(const char* filename, pbso_ffat** out);
/* Build a set from arrays (the fields ffat_map_serialize.h:55-79 keeps), one record per map:
 * geom[32] = {cellSize, lowCorners[6][3], shell centre[3], bboxLow[3], bboxTop[3], map centre[3], k}
 * igeom[18] = {N_elements[6][2], strides[6]};  psi = n_maps blocks of psi_len doubles (column 0). */
int pbso_ffat_create(int n_maps, const int* mode_ids, const double* geom, const int* igeom,
                     const double* psi, int psi_len, const unsigned char* is_compressed,
                     pbso_ffat** out);
int pbso_ffat_destroy(pbso_ffat* f);
int pbso_ffat_num_maps(const pbso_ffat* f, int* n);
int pbso_ffat_mode_ids(const pbso_ffat* f, int* ids);
/* Field access by mode id, for FFAT_Map_Serialize_Double::Check-style comparisons
 * (ffat_map_serialize.h:281-329).  psi may be NULL to query psi_len only. */
int pbso_ffat_get_map(const pbso_ffat* f, int mode_id, double* geom32, int* igeom18,
                      int* psi_len, int* psi_cols, int* is_compressed, double* psi);
/* FFAT_Map_Serialize_Double::Save (ffat_map_serialize.h:90-164). */
int pbso_ffat_save_file(const pbso_ffat* f, int mode_id, const char* filename);
/* The LEGACY .fatcube form -- libigl's igl::serialize of the FFAT_Map<T,3> object, what FFAT_Map<T,3>::Save / Load / LoadAll
 * (ffat_solver.h:1066-1085) write and read -- is recognised by pbso_ffat_load_file / pbso_ffat_load_dir by its first chunk
 * header ("serial_map_ch3") and read without libigl (the format is restated in csrc/fatcube_codec.h); this writes it:
 * shell 2 three times (GetMapVal reads shell 2 only), g++ type strings. */
int pbso_ffat_save_legacy_file(const pbso_ffat* f, int mode_id, const char* filename);
/* ModalSolver::computeTransfer (modal_solver.h:286-315) for L listeners at once (kernel K3):
 * out[l*n_modes + m] = |maps.at(m).GetMapVal(pos_l, use_compressed)|, m in [0, n_modes)
 * (ffat_solver.h:1180-1206 -> Intersect :676-712 -> Interpolate :736-803 -> Reconstruct :899-906).
 * pos is L x 3.  A missing mode id in [0, n_modes) returns PBSO_ERR_RANGE (.at() throws). */
int pbso_ffat_eval(const pbso_ffat* f, int n_modes, const double* pos, int L, int use_compressed,
                   double* out);
/* FFAT_Map<T,3>::Compress (ffat_solver.h:1125-1178), split where the reference goes through a JPEG file on disk:
 *   pbso_ffat_quantise        :1128-1147  per face A *= 255/maxAmp, cv::Mat::convertTo(CV_8U) (cvRound = round half to
 *                             even, saturated to [0,255]) -> q8[psi_len], max_amp6[6]; returns maxAmp_global (:1132-1136).
 *   (the reference's cv::imwrite/cv::imread JPEG round trip of every face image, :1149-1158, belongs between the two
 *    calls; this library has no image codec -- a caller who wants the lossy step runs it on q8 there)
 *   pbso_ffat_set_compressed_u8 :1159-1173 _compressed_Psi = q8 * (maxAmp/255.) per face, _is_compressed = true.  _Psi
 *                             stays, so both views of GetMapVal(p, getCompressed) are there afterwards.
 *   pbso_ffat_compress        both halves back to back for one map or (mode_id -1) all; max_amp_global gets one
 *                             value per compressed map in mode-id order (may be NULL).
 * The device keeps the compressed view as ONE BYTE per texel + maxAmp/255 per (map, face) and evaluates
 * w * ((double)q * scale): bit for bit what reading the reference's stored doubles gives, at an eighth of the bytes.
 * A map LOADED compressed (ffat_map_serialize.h:238-252) has only its doubles; it cannot be compressed again. */
int pbso_ffat_quantise(const pbso_ffat* f, int mode_id, unsigned char* q8, double* max_amp6,
                       double* max_amp_global);
int pbso_ffat_set_compressed_u8(pbso_ffat* f, int mode_id, const unsigned char* q8, int len,
                                const double* max_amp6);
int pbso_ffat_compress(pbso_ffat* f, int mode_id, double* max_amp_global);
/* q8[psi_len], maxAmp per face and/or the doubles of _compressed_Psi (any may be NULL); q8 and max_amp6 are what
 * pbso_ffat_set_compressed_u8 takes, so a map's compressed view can be carried into another set. */
int pbso_ffat_get_compressed(const pbso_ffat* f, int mode_id, unsigned char* q8, double* max_amp6,
                             double* compressed_psi);
/* Device-resident variant, enqueue only (cuda_stream NULL = the handle's own stream).  A handle keeps per-call scratch
 * (listener stencils, per-tile bins and their ping-pong counters), so calls on one handle must come from one thread and
 * must not overlap on different streams; successive calls on one stream are ordered by the stream. */
int pbso_ffat_eval_device(const pbso_ffat* f, int n_modes, const double* d_pos, int L,
                          double* d_out, void* cuda_stream);
/* Same with GetMapVal's getCompressed switch (pbso_ffat_eval_device reads the uncompressed view). */
int pbso_ffat_eval_device_view(const pbso_ffat* f, int n_modes, const double* d_pos, int L,
                               int use_compressed, double* d_out, void* cuda_stream);

/* ---- FFAT map construction: ffat_solver.h:944-1069 ------------------
 * The step before the synthesis path: fit the run-time map Psi from the Dirichlet pressure an acoustic solver
 * sampled on nested cube shells.  A fitter holds what FFAT_Map<T,3>(modeId, cellSize, V, N_elements)
 * (ffat_solver.h:944-989; shells :399-428) derives from the mesh -- it is shared by every mode of an object.
 *   V          n_rows x 3 row-major quad vertices, 4 per quad, shells back to back (CubemapMesh order, :334-397);
 *              only each face's first vertex (its low corner) is read, as in the reference
 *   n_elements [n_shells][6][2] quads per face (+x,-x,+y,-y,+z,-z);  n_shells in [3,8]: shell 2 is the one the
 *              run-time map keeps (_shells.at(2), :982, :1189)
 * The reference takes the shell bounding box as min/max against uninitialised members (:423-428, undefined
 * behaviour); here the bounds start from the first low corner, which equals a zero-filled start whenever the box
 * straddles the origin. */
typedef struct pbso_ffat_fitter pbso_ffat_fitter;
int pbso_ffat_fitter_create(double cell_size, const double* V, int n_rows, const int* n_elements,
                            int n_shells, pbso_ffat_fitter** out);
int pbso_ffat_fitter_destroy(pbso_ffat_fitter* f);
/* _N_elements_total, _N_directions, FFAT_Map<T,3>::_strides (:956-965, :983-986); any pointer may be NULL. */
int pbso_ffat_fitter_info(const pbso_ffat_fitter* f, int* n_shells, int* n_elements_total,
                          int* n_directions, int* shell_strides);
/* One shell as the geom[32]/igeom[18] record pbso_ffat_create takes (k = -1 until solved; map centre = shell
 * 2's centre): shell 2's record plus a solved Psi and its k is a complete run-time map. */
int pbso_ffat_fitter_shell(const pbso_ffat_fitter* f, int shell, double* geom32, int* igeom18);
/* FFAT_Map<T,3>::Solve(k, dirichletPressure, powerScaling) (:1007-1069 -> FFAT_Solver<T,3>::Solve :872-897,
 * Scaling :909-929) for n_maps modes at once (kernel K6):  k[n_maps];  pressure = n_maps vectors of
 * 2*n_elements_total complex doubles (re,im interleaved; two entries per quad, the even one is read, :1054);
 * psi[n_maps][n_directions];  scale[n_maps] or NULL receives Scaling's return value (1 without power scaling). */
int pbso_ffat_fitter_solve(pbso_ffat_fitter* f, int n_maps, const double* k, const double* pressure,
                           int flags, double* psi, double* scale);
/* `flags` (the reference's bool powerScaling is values 0 and 1):
 *   PBSO_FIT_POWER_SCALING  Scaling (:909-929) is applied: psi *= sqrt(sum |p_0|^2 / sum (psi/(k r_0))^2)
 *   PBSO_FIT_DEFER_SCALE    with POWER_SCALING: psi is left unscaled and scale[] (required) holds the factor -- for a consumer
 *                           that folds it in (K3 evaluates |psi / (k r)|: dividing k by the factor does it) and saves the
 *                           second pass over Psi
 *   PBSO_FIT_PACKED         pressure holds ONE complex per quad, n_elements_total per mode -- entry i is the reference
 *                           layout's entry 2 i (the odd entries, the second triangle of every quad, are never read,
 *                           :1054-1056): half the bytes and half the sectors for a solver that can write it that way */
#define PBSO_FIT_POWER_SCALING 1
#define PBSO_FIT_DEFER_SCALE 2
#define PBSO_FIT_PACKED 4
/* Device-resident variant, enqueue only (cuda_stream NULL = the fitter's own stream); d_scale may be NULL. */
int pbso_ffat_fitter_solve_device(pbso_ffat_fitter* f, int n_maps, const double* d_k,
                                  const double* d_pressure, int flags, double* d_psi,
                                  double* d_scale, void* cuda_stream);
/* CUDA-event time of the kernels of the last pbso_ffat_fitter_solve call (copies excluded). */
int pbso_ffat_fitter_last_kernel_ms(const pbso_ffat_fitter* f, float* ms);

/* ---- mode shapes and impulse projection U^T f ---------------------------------------- */
/* ModeData<REAL>::_modes (ModeData.h:23-24), mode-major U[M][K]. */
int pbso_modes_upload(const double* U, int M, int K, pbso_modes** out);
/* ModeData<REAL>::read (ModeData.h:61-83) straight to the device; omega_squared[M] may be NULL. */
int pbso_modes_read_file(const char* filename, pbso_modes** out, int* M, int* K);
int pbso_modes_omega_squared(const pbso_modes* md, double* omega_squared);
int pbso_modes_destroy(pbso_modes* md);
/* GetModalForceVertex (tools/real_time_modal_sound.cpp:268-280): out[m] = vn . U_m[3vid..3vid+2] */
int pbso_modes_project_vertex(const pbso_modes* md, int force_dim, int vid, const double* vn3,
                              double* out);
/* GetModalForceFace (tools/real_time_modal_sound.cpp:236-251). */
int pbso_modes_project_face(const pbso_modes* md, int force_dim, const int* vids3,
                            const double* coords3, const double* vn3, double* out);
/* Batched sparse form: B vertex impulses at once; out is B x force_dim (row per impulse). */
int pbso_modes_project_vertices(const pbso_modes* md, int force_dim, int B, const int* vids,
                                const double* vn, double* out);
/* Dense form for B impulses with dense load vectors: Y[b][m] = sum_k U[m][k] F[b][k]; F is B x K (one load
 * vector per row), Y is B x force_dim (one ForceMessage::data per row).
 *   PBSO_PREC_F64     B == 1: HBM-bound FP64 GEMV (kernel K4); B > 1: FP64 tile GEMM -- exact-parity path
 *   PBSO_PREC_TF32X3  kernel K5: tcgen05 tensor cores, operands split hi+lo into TF32, three MMAs per product,
 *                     FP32 accumulation in TMEM (column rel-L2 error ~1e-6) */
int pbso_modes_project_dense(const pbso_modes* md, int force_dim, const double* F, int B, double* Y,
                             int precision);
/* CUDA-event time of the kernel(s) of the last pbso_modes_project_dense call (copies excluded). */
int pbso_modes_last_kernel_ms(const pbso_modes* md, float* ms);
/* K5 with device-resident FP32 inputs/outputs (d_F[B][K], d_Y[B][force_dim]); enqueue only. */
int pbso_modes_project_dense_device(const pbso_modes* md, int force_dim, const float* d_F, int B,
                                    float* d_Y, void* cuda_stream);

/* Contact storm, one audio buffer, projection -> load -> integrator without leaving the device:
 * B vertex impulses land on sample 0 of the buffer; their modal loads add up (superposition; the caller of the reference
 * sums them into the ONE ForceMessage a buffer can take, modal_solver.h:184, 206-221):
 *     space[m] = sum_b vn_b . U_m[3 vid_b .. 3 vid_b + 2]     (GetModalForceVertex, tools/real_time_modal_sound.cpp:268-280)
 *     y        = pbso_render_buffer(it, space, PointForce profile, T)
 * PBSO_PREC_TF32X3: the impulses are expanded to dense load vectors F[B][K] on the device and contracted with U by the
 * tensor-core GEMM (kernel K5) -- the batched projection of the north star; PBSO_PREC_F64: sparse gather (K4s), the
 * reference's arithmetic.  Everything runs on the integrator's stream; host pointers in and out (pinned staging inside);
 * y_out[L*T] listener-major, qnorm_out[N] or NULL.  force_dim must equal the integrator's N. */
int pbso_modes_storm_buffer(pbso_modes* md, pbso_integrator* it, int force_dim, int B, const int* vids,
                            const double* vn, int T, double* y_out, double* qnorm_out, int precision);

/* ---- offline batch renderer (many objects x long audio) ------------- */
/* n_obj independent sound objects with n_modes each; (a,b) are n_obj x n_modes as accepted by
 * ModalIntegrator(N,h,a,b).  Each object behaves exactly like its own ModalSolver::step loop
 * (modal_solver.h:181-276) fed with PointForce messages (forces.h:81-90). */
int pbso_batch_create(int n_obj, int n_modes, double h, const double* a, const double* b,
                      pbso_batch** out);
int pbso_batch_destroy(pbso_batch* bt);
/* One TransMessage per object, applied from buffer 0 (static listeners): trans[n_obj][n_modes]. */
int pbso_batch_set_transfer(pbso_batch* bt, const double* trans);
/* Impulse script: event e = a ForceMessage{data = space[e][n_modes], PointForce} enqueued so that
 * object obj[e] dequeues it in step #buf[e] (impulse lands on sample 0 of that buffer, forces.h:87).
 * step() dequeues at most one message per buffer (modal_solver.h:184), so two events of one object
 * in the same buffer are rejected with PBSO_ERR_INVALID. */
int pbso_batch_set_impulses(pbso_batch* bt, int n_events, const int* obj, const int* buf,
                            const double* space);
/* The same two calls without the final host synchronisation: the copies are only ENQUEUED on the handle's stream, so
 * `trans` / `space` must stay valid and unchanged until the stream has passed them (pbso_batch_sync, or a later call
 * that synchronises such as pbso_batch_render_mix).  With pinned buffers a whole step -- inputs in, render, audio
 * reduce, mix out -- then costs ONE host synchronisation. */
int pbso_batch_set_transfer_async(pbso_batch* bt, const double* trans);
int pbso_batch_set_impulses_async(pbso_batch* bt, int n_events, const int* obj, const int* buf,
                                  const double* space);
/* Render n_buffers x buf_size samples of every object from zero state and mix them down:
 * mix[i] = sum_obj y_obj[i]  (double, n_buffers*buf_size).  With PBSO_PREC_F32_TILED and buf_size 256 the
 * render is also parallel over n_chunks independent time chunks (0 = choose from the SM count) whose start
 * states come from closed-form pole powers of the impulses before them; FP64 runs one chunk. */
int pbso_batch_render_mix(pbso_batch* bt, int buf_size, int n_buffers, int precision,
                          int n_chunks, double* mix);
/* Device-resident variant: enqueue only; d_mix (double[n_buffers*buf_size]) is zeroed then
 * accumulated on the handle's stream -- e.g. straight into an NCCL send buffer. */
int pbso_batch_render_mix_device(pbso_batch* bt, int buf_size, int n_buffers, int precision,
                                 int n_chunks, double* d_mix);
/* Per-object stems: y[n_obj][n_buffers*buf_size] float32 (y as in SoundMessage, before /1e10). */
int pbso_batch_render_stems(pbso_batch* bt, int buf_size, int n_buffers, int precision,
                            float* stems);
/* Stateful range renders.  A render normally starts from the zero state of a fresh ModalSolver; pbso_batch_set_state
 * makes the next renders start from (q_km1, q_km2) = (q[k-1], q[k-2]) per (object, mode) -- the pair ModalIntegrator keeps
 * (modal_integrator.h:106-113) and pbso_integrator_get_state returns -- and pbso_batch_get_end_state returns that pair
 * after n_buffers x buf_size samples of the current state + impulse script, in closed form (FP64 pole powers, whatever
 * precision the audio was rendered in).  A long script is then rendered range by range: a TransMessage that arrives in
 * mid-render (modal_solver.h:249-252) is a range boundary with pbso_batch_set_transfer in between, and buffers in which a
 * Gaussian or autoregressive force is alive (forces.h:92-137) go through the per-buffer path (pbso_render_buffer on a
 * pbso_integrator holding the same state) between two batch ranges.  Both arrays NULL = back to the zero state. */
int pbso_batch_set_state(pbso_batch* bt, const double* q_km1, const double* q_km2);
int pbso_batch_get_end_state(pbso_batch* bt, int buf_size, int n_buffers, double* q_km1, double* q_km2);
int pbso_batch_sync(pbso_batch* bt);
/* Run this handle's work on a caller-owned CUDA stream (cudaStream_t as void*; NULL restores the
 * handle's own stream) so that renders order with the caller's copies / NCCL calls without host syncs. */
int pbso_batch_set_stream(pbso_batch* bt, void* cuda_stream);
/* Accumulate-truncation gain of the tensor-core path on the current device, measured by the one-time self-calibration
 * that the first PBSO_PREC_TC3X render runs (a synthetic batch against the FP64 kernel); 0 before that. */
int pbso_tc_gain(double* gain);
/* CUDA-event time of the last render kernel(s) on the handle's stream, and launches issued. */
int pbso_batch_last_kernel_ms(pbso_batch* bt, float* ms, int* launches);

/* ---- multi-GPU: object / mode-block sharding and the audio reduce -------------
 * Sound objects are independent, and so are the modes of one object (the reference's hot loop is a sum over modes,
 * modal_solver.h:261-272): every rank renders a contiguous block of (object, mode block) units into its own FP64 mix,
 * and one sum-reduce of the audio is the only exchange.  One rank per GPU; NCCL (libnccl.so.2) is loaded on first use.
 *   pbso_comm_unique_id   rank 0 creates the 128-byte id, the caller hands it to every rank (file, socket, MPI ...)
 *   pbso_comm_init        collective; binds the communicator to the calling thread's current device
 *   pbso_comm_shard       this rank's block [lo, hi) of n_units; the blocks tile [0, n_units)
 *   pbso_comm_reduce_audio  d_audio[n] (double, device) summed over the ranks onto `root` in place (root = -1: onto
 *                         every rank), enqueued on cuda_stream -- the stream the render was enqueued on, so no host
 *                         synchronisation sits between render and reduce */
int pbso_comm_unique_id(unsigned char* id128);
int pbso_comm_init(int nranks, int rank, const unsigned char* id128, pbso_comm** out);
int pbso_comm_destroy(pbso_comm* c);
int pbso_comm_info(const pbso_comm* c, int* nranks, int* rank, int* nccl_version);
int pbso_comm_shard(const pbso_comm* c, long long n_units, long long* lo, long long* hi);
int pbso_comm_reduce_audio(pbso_comm* c, double* d_audio, size_t n, int root, void* cuda_stream);
/* Same with the audio in host memory (staged through a device buffer inside; blocking) -- for callers that hold no
 * device pointers of their own, like tools/pbso_render -batch -gpus N.  Ranks other than root keep their input. */
int pbso_comm_reduce_audio_host(pbso_comm* c, double* audio, size_t n, int root);

/* ---- measurement helpers (device micro-benchmarks used by bench.py for roofline peaks) -- */
/* kind 0: FP32 FFMA (uniform operands); 1: packed fma.rn.f32x2 (uniform operands); 2: FP64 DFMA;
 * 3: FFMA with three distinct registers; 4: fma.rn.f32x2 with three distinct register pairs.
 * Returns measured TFLOP/s. */
int pbso_measure_fma_peak(int kind, double* tflops, double* sm_mhz_est);
/* Bare tcgen05.mma loop (A operand in TMEM, B in shared memory), one persistent CTA (cta_group 1) or CTA pair
 * (cta_group 2) per SM: the measured tensor-pipe peak that the tensor rooflines of the batch renderer and the
 * batched projection divide by.  kind 0: kind::tf32 (K = 8 per MMA), 1: kind::f16 (K = 16); n = MMA N (128 / 256);
 * stress != 0 also runs four warps of conflict-free shared-memory stores beside the MMAs and reports how many
 * wavefronts per cycle per SM they sustained.  Outputs: TFLOP/s of the whole device, median cycles per MMA. */
int pbso_measure_tc_peak(int kind, int cta_group, int n, int stress, double* tflops,
                         double* cycles_per_mma, double* stress_wavefronts_per_cycle);
/* The same loop launched back to back for at least min_ms of device time: the SUSTAINED tensor rate under the board's
 * power cap -- the roofline denominator of a kernel timed inside a long step (the burst figure is for a kernel timed alone). */
int pbso_measure_tc_peak_sustained(int kind, int cta_group, int n, double min_ms, double* tflops);
/* One [256 x 128] x K product on a CTA pair (tcgen05.mma cta_group::2, A from TMEM) with integer-valued operands:
 * max |D - A B^T| (0 when the operand placement assumed by the pair kernels is right). */
int pbso_tc_selftest(int kind, double* max_err);
/* STREAM-style device copy bandwidth in GB/s (read+write bytes). */
int pbso_measure_copy_bw(size_t bytes, double* gbs);
/* Writes `bytes` to a scratch buffer to evict L2 between timed iterations. */
int pbso_flush_l2(size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PBSO_B200_H */
