/*
 * pnpb200.h -- C ABI of the B200-native batched PnP solver (libpnpb200.so).
 *
 * The reference (Benson516/pnp_solver_test) is pure Python and has no FFI boundary: its
 * boundary is the class PNP_SOLVER in scripts/PNP_SOLVER_LIB.py.  Every entry point below
 * names the reference interface it replaces (file:line, relative to the reference root).
 * The Python mirror of that class (pnp_solver_test_b200/solver.py) binds these symbols with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes.  No torch types.
 *  - Pointers marked [device] are CUDA device pointers owned by the caller, 16-byte aligned;
 *    [host] are ordinary host pointers read during the call.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *    stream) unless stated otherwise.  The library allocates nothing persistent except inside
 *    a pnpb200_pipeline (host entry point), which the caller creates and destroys.
 *  - Return value: 0 on success, a negative PNPB200_E* code otherwise.  Like the reference,
 *    the numeric path never raises: a diverged solve returns garbage R/t and a large res_norm.
 *  - dtype selects BOTH the I/O element type of the float arrays and the arithmetic type.
 *    FP64 is the parity mode (the reference is float64 only).
 */
#ifndef PNPB200_H
#define PNPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNPB200_VERSION 100

/* method: which single-pattern solver of the reference to run */
#define PNPB200_METHOD_QEIF       0  /* solve_pnp_QEIF_single_pattern,          PNP_SOLVER_LIB.py:2771-3025 (what solve_pnp() dispatches to, :179) */
#define PNPB200_METHOD_LM         1  /* solve_pnp_LM_single_pattern,            PNP_SOLVER_LIB.py:2567-2769 */
#define PNPB200_METHOD_LINEAR_F2  2  /* solve_pnp_formulation_2_single_pattern, PNP_SOLVER_LIB.py:693-953   */
#define PNPB200_METHOD_LINEAR_F1  3  /* solve_pnp_single_pattern,               PNP_SOLVER_LIB.py:205-430   */
#define PNPB200_METHOD_LM_PLUS    4  /* NOT in the reference (non-parity extra): linear F2 initial pose, then LM's 12-state
                                        damped Gauss-Newton with the true constraint gradients and a convergence test
                                        (max |dx| <= 1e-10); res_norm at the returned state; moment mapping only */
#define PNPB200_METHOD_EIF2       5  /* solve_pnp_EIF2_single_pattern,          PNP_SOLVER_LIB.py:2001-2276: iterated information filter
                                        on LM's 12-state model with EKF2_get_process_covariance_R (:3668-3716) and QEIF's early exit */

#define PNPB200_DTYPE_F64 0
#define PNPB200_DTYPE_F32 1

/* execution shape (0 = let the library choose from n) */
#define PNPB200_MAP_AUTO    0
#define PNPB200_MAP_THREAD  1   /* one problem per thread, correspondences staged in shared memory */
#define PNPB200_MAP_MOMENT  2   /* moments -> O(1) iterations -> point-wise residual, as three      */
                                /* streaming kernels: LM, LM+, linear F2 and, in FP64, EIF2 and     */
                                /* QEIF from 12 landmarks on (their exit decisions certified, the   */
                                /* rest re-solved point-wise by a fix-up pass).  The default there; */
                                /* several patterns: one such solve per pattern, arg-min in between */
#define PNPB200_MAP_WARP    32  /* one problem per warp, shuffle-reduced normal equations          */

#define PNPB200_OK          0
#define PNPB200_EINVAL     -1   /* bad argument (null pointer, n < 1, unknown method/dtype, ...)   */
#define PNPB200_ECUDA      -2   /* a CUDA runtime call failed; see pnpb200_last_error()            */
#define PNPB200_ENODEVICE  -3   /* no usable CUDA device                                           */
#define PNPB200_ETOOLARGE  -4   /* n exceeds what the selected mapping can hold                    */

#define PNPB200_FLAG_PROFILE 1  /* record CUDA events around each kernel of the call (pnpb200_profile_read) */
#define PNPB200_FLAG_QEIF_DIRECT 2 /* QEIF in the direct mappings: accumulate H^T H point by point even for n >= 12 (default there: from
                                      the moments); with PNPB200_MAP_AUTO it also keeps QEIF out of the moment mapping */
#define PNPB200_FLAG_LM_TRUE_JACOBIAN 4 /* method LM only, NOT a parity mode: the reference's loop (identity start, max_it iterations,
                                           constant lambda) with the true gradients of its nine constraint rows (2u for the quadratic rows,
                                           u/|u| for the norm rows) in place of the halved ones it uses (PNP_SOLVER_LIB.py:3787-3823);
                                           moment mapping only */

#define PNPB200_MAX_PATTERNS 8
#define PNPB200_REPORT_WIDTH 16

/* Solver constants.  The reference hard-codes these inline; defaults reproduce it. */
typedef struct pnpb200_params {
    int32_t max_it;        /* 14      PNP_SOLVER_LIB.py:2635 (LM), :2857 (QEIF)                   */
    int32_t linear_it;     /* 3       :223 (F1), :762 (F2)                                         */
    double  lm_lambda;     /* 1e-5    :2631  constant damping                                      */
    double  exit_tol;      /* 1e-2    :2952  QEIF early exit on |d res / res|                      */
    double  f_weight;      /* 225.68  :2845  focal length used in the measurement weight (NOT K)   */
    double  meas_sigma_px; /* 3.0     :2846                                                        */
    double  proc_q;        /* 1e-1    :2792  process noise, quaternion states                      */
    double  proc_d;        /* 1e-2    :2793  process noise, delta states                           */
    double  omega0;        /* 1e-5    :2836  initial information                                   */
    double  res_old0;      /* 1e-7    :2863                                                        */
    int32_t mapping;       /* PNPB200_MAP_*                                                        */
    int32_t flags;         /* PNPB200_FLAG_*                                                       */
    void*   workspace;     /* [device] optional scratch of workspace_bytes (pnpb200_workspace_bytes);  */
    int64_t workspace_bytes; /* NULL/0: the call takes it from the stream-ordered CUDA memory pool   */
} pnpb200_params;

/* Synthetic workload description (random_stress_test.py:246-258, LM_noise_test.py:141-189). */
typedef struct pnpb200_synth {
    uint64_t seed;            /* Philox4x32-10 key; the counter is the GLOBAL problem index        */
    double   angle_range_deg; /* 45    roll, pitch, yaw ~ U(-a, a)                                 */
    double   depth_min_m;     /* 0.20                                                              */
    double   depth_max_m;     /* 2.25                                                              */
    double   fov_max_deg;     /* 45    t = depth * [tan U(-f,f), tan U(-f,f), 1]                   */
    int32_t  is_quantized;    /* round pixels to multiples of quantize_q (np.around)               */
    int32_t  reserved;
    double   quantize_q;      /* 1.0                                                               */
    double   noise_sigma_px;  /* 0 = none; Gaussian pixel noise added after quantisation           */
    double   roll_center_deg; /* 0    angles are centre + U(-angle_range, angle_range); a fixed    */
    double   pitch_center_deg;/* 0    pose (LM_noise_test.py:180-182, :262) is range 0 + centres   */
    double   yaw_center_deg;  /* 0                                                                 */
} pnpb200_synth;

int pnpb200_version(void);
int64_t pnpb200_launch_count(void);             /* kernels launched by this library since it was loaded (all threads) */
const char* pnpb200_last_error(void);          /* text of the last CUDA error seen by this thread */
int pnpb200_default_params(pnpb200_params* p);
int pnpb200_default_synth(pnpb200_synth* s);
int pnpb200_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes);
/* scratch bytes pnpb200_solve_batch needs for this call shape (0 for the direct mappings; QEIF: sized for the
   moment mapping, which it uses from 12 landmarks on) */
int64_t pnpb200_workspace_bytes(int method, int dtype, int64_t B, int n_patterns, int mapping);

/*
 * The hot path.  Replaces the per-problem loop around PNP_SOLVER.solve_pnp
 * (PNP_SOLVER_LIB.py:144-203; callers random_stress_test.py:322, LM_noise_test.py:312,
 * face_variation_test.py:431) and the solve_pnp_*_single_pattern methods it dispatches to.
 * One launch does, per problem: correspondence packing and K^-1 normalisation
 * (f2_get_P :3260, f2_get_B_xy :3291), the selected solver for each of the n_patterns patterns,
 * the arg-min over patterns on res_norm (strict <, first wins; :185), (R, t) reconstruction
 * (:3500 / :3542 / :4158) and Euler extraction (:4474).
 *
 *   uv        [device] [B, n_total, 2] pixels (u, v); the homogeneous coordinate is 1
 *   pattern   [device] [n_patterns, n_total, 3] metres, shared by all problems
 *   point_index [host] n indices into 0..n_total-1 selecting (and ordering) the landmarks the
 *             solver sees -- the LM_key_list of PNP_SOLVER_LIB.py:156 -- or NULL for all n_total
 *   K         [host] [3,3] row-major camera matrix (np_K_camera_est)
 *   params    [host] or NULL for defaults
 *   R [B,9]  t [B,3]  euler_deg [B,3] = (roll, yaw, pitch)  res_norm [B]      [device], dtype
 *   iters [B]  best_pattern [B]                                           [device], int32
 *             (any output pointer may be NULL to skip it)
 */
int pnpb200_solve_batch(int method, int dtype, int64_t B, int n_total, int n,
                        const void* uv, const void* pattern, int n_patterns,
                        const int32_t* point_index, const double* K,
                        const pnpb200_params* params,
                        void* R, void* t, void* euler_deg, void* res_norm,
                        int32_t* iters, int32_t* best_pattern, void* stream);

/*
 * Image points whose homogeneous coordinate is not 1.  The reference's packing stage multiplies the (3,1) vector it is
 * given by K^-1 without looking at its third entry (f2_get_B_xy, PNP_SOLVER_LIB.py:3305-3307), and its own projection
 * emits w = -1 for points behind the camera (perspective_projection :4548).  This entry computes that product for
 * n_points points, uvw [n_points, 3] -> uv_normalised [n_points, 2] = rows 0, 1 of K^-1 [u, v, w]^T (both device, dtype);
 * pnpb200_solve_batch is then called on uv_normalised with K = identity.  (Pixels with w = 1 need none of this.)
 */
int pnpb200_normalise_uvw(int dtype, int64_t n_points, const void* uvw, const double* K, void* uv_normalised, void* stream);

/*
 * Per-kernel device times of the pnpb200_solve_batch calls made by this thread with
 * PNPB200_FLAG_PROFILE since the last reset, measured with CUDA events on the call's stream.
 * ms[3] = average duration of (moments | direct solve kernel, iterate, residual); kernels a
 * mapping does not launch report 0.  pnpb200_profile_read synchronises with the recorded events.
 */
int pnpb200_profile_reset(void);
int pnpb200_profile_read(float* ms, int* n_calls);

/*
 * Same, from HOST buffers (pageable or pinned), chunked and double-buffered through a
 * caller-created pipeline so that H2D copies, the solve kernel and D2H copies overlap.
 * This is the call the drop-in PNP_SOLVER.solve_pnp_batch_host() makes for NumPy inputs and what
 * bench.py times as `e2e`.  Blocks until the results are in the host buffers.
 */
typedef struct pnpb200_pipeline pnpb200_pipeline;
int pnpb200_pipeline_create(pnpb200_pipeline** out, int dtype, int64_t chunk_problems, int n_total,
                            int n_patterns, int n_streams);
int pnpb200_pipeline_destroy(pnpb200_pipeline* p);

/*
 * Packed pixel transfer for pnpb200_solve_batch_host.  Detected landmarks are whole pixels
 * (random_stress_test.py projects with is_quantized=True, PNP_SOLVER_LIB.py:4549-4552) that the
 * reference carries as float64.  With n_threads > 0 a second host thread takes chunks from the far end
 * of the batch, converts each to int16 with n_threads workers (pnpb200_pack_i16) and, if that was exact
 * for every value of the chunk, ships the int16 copy (a quarter of the PCIe bytes) and widens it on the
 * device -- the solver sees the same values; chunks that are not whole numbers travel as they are.
 * The calling thread keeps sending chunks unchanged from the near end (in eighths, two in flight, so that a
 * packed copy never queues behind a whole plain chunk), so PCIe and the CPU work at the same time and share
 * the batch by their speeds.  n_threads = 0 switches it off (the default).
 * pnpb200_pipeline_last_packed: how many chunks of the last call travelled packed, of how many.
 */
int pnpb200_pipeline_set_packing(pnpb200_pipeline* p, int n_threads);
int pnpb200_pipeline_last_packed(const pnpb200_pipeline* p, int64_t* packed_chunks, int64_t* chunks);

/*
 * Host helper of the packed transfer: dst[i] = (int16) src[i] for n_values FP64 / FP32 values (dtype),
 * on n_threads threads (the calling thread and n_threads - 1 workers that stay alive, asleep, between calls;
 * concurrent calls take turns).  Returns 1 if every value was a whole number in [-32768, 32767] (the packing is
 * lossless; -0.0 counts as 0), 0 if not (dst is then not to be used), < 0 on a bad argument.
 */
int pnpb200_pack_i16(int dtype, const void* src, int64_t n_values, int16_t* dst, int n_threads);
/* params->workspace must be NULL here (EINVAL otherwise): the chunks run concurrently and the pipeline owns one scratch per stream */
int pnpb200_solve_batch_host(pnpb200_pipeline* p, int method, int64_t B, int n,
                             const void* uv_host, const void* pattern_host,
                             const int32_t* point_index, const double* K,
                             const pnpb200_params* params,
                             void* R, void* t, void* euler_deg, void* res_norm,
                             int32_t* iters, int32_t* best_pattern);

/*
 * The same for callers that hold their detections in a narrow type.  Landmark detectors deliver whole pixels
 * (the reference rounds them itself: is_quantized=True, PNP_SOLVER_LIB.py:4549-4552, random_stress_test.py:290)
 * and only the reference's dict-of-float64 convention makes them 16 bytes per landmark.  uv_host is
 * [B, n_total, 2] of pixel_type; a half (float32) or a quarter (int16 / uint16) of the bytes cross PCIe and the
 * values are widened on the device to the pipeline's arithmetic type -- exactly, so the results are bit-identical
 * to those of the same values handed over in the pipeline's dtype.  PNPB200_PIXEL_NATIVE = pnpb200_solve_batch_host.
 * No host pass, no packing thread: nothing on this path touches the pixels on the CPU.
 */
#define PNPB200_PIXEL_NATIVE 0   /* the pipeline's dtype */
#define PNPB200_PIXEL_I16    1
#define PNPB200_PIXEL_U16    2
#define PNPB200_PIXEL_F32    3   /* float32 pixels, arithmetic in the pipeline's dtype */
int pnpb200_solve_batch_host_px(pnpb200_pipeline* p, int method, int64_t B, int n, int pixel_type,
                                const void* uv_host, const void* pattern_host,
                                const int32_t* point_index, const double* K,
                                const pnpb200_params* params,
                                void* R, void* t, void* euler_deg, void* res_norm,
                                int32_t* iters, int32_t* best_pattern);

/*
 * Page-locked host memory for the buffers of the two calls above (cudaHostAlloc): copies from / to pinned memory
 * run asynchronously at full PCIe rate, pageable buffers are staged by the driver and block the submitting thread.
 * write_combined = 1 for buffers the CPU only WRITES (the pixels): not snooped, faster for the device to read,
 * very slow for the CPU to read back -- do not use it for the result buffers or with the packed transfer.
 */
int pnpb200_host_alloc(void** out, int64_t bytes, int write_combined);
int pnpb200_host_free(void* ptr);

/*
 * Euler <-> R, batched on the device.
 * Replaces get_rotation_matrix_from_Euler (PNP_SOLVER_LIB.py:4442-4472) and
 * get_Euler_from_rotation_matrix (:4474-4517).  euler = (roll, yaw, pitch).
 */
int pnpb200_R_from_euler(int dtype, int64_t B, const void* euler, int is_degree, void* R, void* stream);
int pnpb200_euler_from_R(int dtype, int64_t B, const void* R, int is_degree, void* euler, void* stream);

/*
 * Pinhole projection of a pattern, batched.  Replaces perspective_projection (:4532-4557) and
 * perspective_projection_golden_landmarks (:4586-4621): K (R theta + t) / |z|, optional
 * rounding to a grid (np.around = half to even).  uvw: [B, n, 3] homogeneous pixels.
 */
int pnpb200_project(int dtype, int64_t B, int n, const void* pattern, const double* K,
                    const void* R, const void* t, int is_quantized, double quantize_q,
                    void* uvw, void* stream);

/*
 * Synthetic workload generator, FP64 internally.  Problems b0 .. b0+B-1 of a global
 * counter-based stream, so a shard reproduces exactly the slice of the unsharded run.
 *   uv [B,n,2] (dtype)   gt [B,4] = (distance m, roll, pitch, yaw deg) (f64)
 *   R_gt [B,9], t_gt [B,3] (f64, may be NULL)
 * Workload definition: random_stress_test.py:246-290 (+ LM_noise_test.py noise).
 */
int pnpb200_synth_batch(int dtype, int64_t b0, int64_t B, int n, const void* pattern_f64,
                        const double* K, const pnpb200_synth* cfg,
                        void* uv, double* gt, double* R_gt, double* t_gt, void* stream);

/*
 * The workload of face_variation_test.py (:296-356): the same pose stream, but the pixels are the
 * projection of a PERTURBED pattern -- a random direction of the 3 (n - 1) coordinates of all
 * landmarks but `fixed_index` ("eye_c_51", :318; -1 = none fixed), scaled to perturb_radius_m
 * (0.02, :319) -- while the solver keeps the golden pattern.  perturb [B,n,3] (f64, may be NULL)
 * receives np_pattern_perturbation_dict of each problem, the input of pnpb200_fragility_accumulate.
 */
int pnpb200_synth_face_variation(int dtype, int64_t b0, int64_t B, int n, const void* pattern_f64,
                                 const double* K, const pnpb200_synth* cfg, double perturb_radius_m,
                                 int fixed_index, void* uv, double* gt, double* R_gt, double* t_gt,
                                 double* perturb, void* stream);

/*
 * Error reporting per problem.  Replaces check_if_the_sample_passed (TEST_TOOLBOX.py:55-62),
 * cal_LM_error_distances (:252-286), the error fields of compare_result_and_generate_result_dict
 * (:396-463) and the glue at random_stress_test.py:353-377.  Always FP64 outputs.
 *   report [B,16]: 0 depth_err(m) 1 roll_err 2 pitch_err 3 yaw_err (deg)
 *                  4,5 LM_GT avg,max  6,7 predict_LM avg,max  8,9 predict_GT avg,max (x distance_GT)
 *                  10 t3_est 11 distance_GT 12 roll_est 13 pitch_est 14 yaw_est 15 reserved
 *   flags [B,4] = depth, roll, pitch, yaw passed;  max_idx [B,3] = landmark index of each max
 *   bounds [host] 4 doubles (10 cm, 10, 10, 10 deg in the reference)
 */
int pnpb200_report_batch(int dtype, int64_t B, int n, const void* pattern, const void* uv,
                         const double* K, const void* R, const void* t, const void* euler_deg,
                         const double* gt, const double* bounds,
                         double* report, int32_t* flags, int32_t* max_idx, void* stream);

/*
 * The solve and the error report of the same problems in one call: pnpb200_solve_batch (one pattern) followed by
 * pnpb200_report_batch_strided over all n_total landmarks -- the body of the loop of random_stress_test.py:322-377
 * (solve_pnp, then compare_result_and_generate_result_dict on its result).  When the solve runs as the moment
 * mapping (LM / linear F2 / LM+, all landmarks) its last pass -- res_norm at the stored state, point by point -- is
 * folded into the report kernel, so the pixel rows are read twice per problem (moments; residual + report) instead
 * of three times; the results are those of the two separate calls.  Every other method / mapping runs the two calls
 * back to back.  R, t, euler_deg, gt and report are required; res_norm, iters, flags, max_idx may be NULL.
 */
int pnpb200_solve_report_batch(int method, int dtype, int64_t B, int n_total, int n,
                               const void* uv, const void* pattern, const int32_t* point_index, const double* K,
                               const pnpb200_params* params,
                               void* R, void* t, void* euler_deg, void* res_norm, int32_t* iters,
                               const double* gt, const double* bounds,
                               double* report, int64_t report_stride_problem, int64_t report_stride_column,
                               int32_t* flags, int32_t* max_idx, void* stream);

/*
 * The same report into a strided array: value k of problem b goes to
 * report[b * report_stride_problem + k * report_stride_column] (elements).  (16, 1) is the row layout
 * of pnpb200_report_batch; (1, B) is the column layout the Python wrapper uses: the kernels' writes
 * and the statistics' reads of one quantity are then contiguous.
 */
int pnpb200_report_batch_strided(int dtype, int64_t B, int n, const void* pattern, const void* uv,
                                 const double* K, const void* R, const void* t, const void* euler_deg,
                                 const double* gt, const double* bounds,
                                 double* report, int64_t report_stride_problem, int64_t report_stride_column,
                                 int32_t* flags, int32_t* max_idx, void* stream);

/*
 * Error statistics, two passes so that shards can be combined with two small all-reduce phases
 * (TEST_TOOLBOX.get_statistic_of_result, TEST_TOOLBOX.py:892-937), for nq <= 4 quantities at once
 * (depth, roll, pitch, yaw in TEST_TOOLBOX.data_analysis_and_saving, :1070-1112), per class and,
 * in the extra LAST row, over all problems.
 *   est[q], gt[q]: FP64 device vectors with element strides est_stride[q], gt_stride[q]
 *                  (gt or gt[q] may be NULL: statistics of the value itself);  arrays of nq [host]
 *   class_id: [B] int32 device or NULL (every problem in class 0); n_class <= 64 classes
 *   pass 1 -> sums1 [nq, n_class+1, 4] = n, sum(est/gt), sum(e), 0                  (SUM-reducible)
 *   pass 2, given the (all-reduced) sums1 from which it derives the means itself
 *          -> sums2 [nq, n_class+1, 4] = sum((e-m)^2), sum|e|, sum|e-m|, 0           (SUM-reducible)
 *             max2  [nq, n_class+1]    = max|e-m|                                    (MAX-reducible)
 */
int pnpb200_stats_pass1(int64_t B, int nq, const double* const* est, const int64_t* est_stride,
                        const double* const* gt, const int64_t* gt_stride,
                        const int32_t* class_id, int n_class, double* sums1, void* stream);
int pnpb200_stats_pass2(int64_t B, int nq, const double* const* est, const int64_t* est_stride,
                        const double* const* gt, const int64_t* gt_stride,
                        const int32_t* class_id, int n_class, const double* sums1, double* sums2,
                        double* max2, void* stream);

/*
 * Ground-truth classification, np.digitize(value, bins) (TEST_TOOLBOX.classify_drpy,
 * TEST_TOOLBOX.py:239-247): class = number of bins b with b <= value.  values: FP64 device,
 * element stride in doubles; bins [host], ascending, n_bins <= 32; class_id [B] int32 device.
 */
int pnpb200_classify(int64_t B, const double* values, int64_t stride, double scale,
                     const double* bins, int n_bins, int32_t* class_id, void* stream);

/*
 * The analysis stage of face_variation_test.py (:631-759) for up to 4 error quantities at once.
 *
 * Selection of the top k problems by |value| -- what the script does by popping k items off heaps of
 * (-|err|, idx) tuples (:631-653), so ties go to the SMALLER index -- as an exact radix select on the
 * 96-bit key (bits of |value|, then ~global index), 12 digits of 8 bits, most significant first.
 * pnpb200_topk_histogram counts, per quantity, the problems whose key starts with the
 * n_digits_decided digits of (prefix_hi, prefix_lo), by their next digit: hist [nq, 256] (device,
 * uint64).  The caller (workload.fragility_analysis) walks the histogram from 255 down to pick the
 * next digit; across GPUs the histograms are summed, nothing else is exchanged.  idx0 = global index
 * of this shard's first problem.  values[q]: device doubles with element stride[q].
 */
int pnpb200_topk_histogram(int64_t B, int64_t idx0, int nq, const double* const* values, const int64_t* stride,
                           const uint64_t* prefix_hi, const uint32_t* prefix_lo, int n_digits_decided,
                           uint64_t* hist, void* stream);

/*
 * get_most_fragile_point_and_perturbation_direction (face_variation_test.py:658-728) over the
 * problems whose key is >= (threshold_hi, threshold_lo): per quantity, n_selected, the sum and the
 * maximum of |value| (top_value_mean, value_max), count [nq, n] = how often each landmark carried
 * the largest perturbation norm (strict >, first landmark wins), and gram [nq, 3n, 3n] = sum of
 * v v^T over the selected perturbation vectors (upper 16 x 16 tiles only; mirror it), whose
 * eigen-decomposition is the script's SVD of the m x 3n perturbation matrix (singular values =
 * sqrt of the eigenvalues, right singular vectors = eigenvectors).  perturb [B, n, 3] from
 * pnpb200_synth_face_variation; list [nq, list_capacity] receives the selected local indices.
 * All outputs are device memory and are zeroed by the call.
 */
int pnpb200_fragility_accumulate(int64_t B, int64_t idx0, int nq, const double* const* values, const int64_t* stride,
                                 const uint64_t* threshold_hi, const uint32_t* threshold_lo, const double* perturb,
                                 int n, int64_t list_capacity, int64_t* list, uint64_t* n_selected, uint64_t* count,
                                 double* value_sum, double* value_max, double* gram, void* stream);

/*
 * Host-side result table (no GPU): the CSV that TEST_TOOLBOX.write_result_to_csv (TEST_TOOLBOX.py:693-708)
 * writes from the list of result dicts of compare_result_and_generate_result_dict (:396-463), straight
 * from the arrays pnpb200_report_batch returns, copied to the host (pinned or not).  Same column names
 * and order for every scalar / string / tuple / dict field (the six ndarray fields are left out), same
 * csv dialect (QUOTE_MINIMAL, "\r\n"), floats as Python's repr(float), bools as True / False.
 *   report [B,16], flags [B,4], max_idx [B,3], res_norm [B], gt [B,4]        (host)
 *   key_names [n_keys]: landmark names in pattern order (the *_error_max_key columns)
 *   bins[q] / n_bins[q] / labels[q], q = depth (cm), roll, pitch, yaw: classify_drpy (:239-247),
 *   n_bins[q] + 1 labels each; they give the `class` and `file_name` columns (random_stress_test.py:300-306)
 *   append = 0: truncate and write the header; 1: append rows (shards / chunks in order); idx0 = first idx
 * pnpb200_format_repr exposes the float formatter for the tests (out_size >= 32).
 */
int pnpb200_write_result_csv(const char* path, int append, int64_t B, int64_t idx0, const double* report,
                             const int32_t* flags, const int32_t* max_idx, const double* res_norm, const double* gt,
                             const char* const* key_names, int n_keys, const double* const* bins, const int32_t* n_bins,
                             const char* const* const* labels, int n_threads);
int pnpb200_format_repr(double v, char* out, int out_size);

/*
 * Combined class of the four ground-truth quantities (TEST_TOOLBOX.get_all_class_seperated_result,
 * :975-1030): id = ((cd * nr + cr) * np + cp) * ny + cy, each c = np.digitize(gt[:, q] * scale[q], bins[q])
 * in the order distance, roll, pitch, yaw; gt [B,4] device; bins / n_bins / scale host arrays of 4.
 * pnpb200_stats_pass1/2 accept the resulting n_class = nd*nr*np*ny (up to 2^20): beyond 64 classes
 * they add straight into the global table.
 */
int pnpb200_classify_drpy(int64_t B, const double* gt, const double* const* bins, const int32_t* n_bins,
                          const double* scale, int32_t* class_id, void* stream);

/*
 * FMA-pipe microbenchmark used by bench.py for the roofline denominator (MEASURED_PEAKS.json
 * has no FP64/FP32 FMA figure).  Runs `iters` dependent-chain-free FMAs per thread on a full
 * grid and returns the device-timed rate in FLOP/s (2 per FMA).  Synchronous.
 */
int pnpb200_fma_peak(int dtype, int iters, double* flops_per_s);

/*
 * Test hook: the branch-free FP64 reciprocal, reciprocal square root and square root that the
 * solvers use for pivots and norms (csrc/pnpb200_math.cuh), element-wise on n device doubles
 * (positive, normal).  No reference counterpart; tests/test_gpu_parity.py bounds their error.
 */
int pnpb200_selftest_math(int64_t n, const double* in, double* rcp, double* rsqrt, double* sqrt_out, void* stream);

/*
 * Test hook: the branch-free sine / cosine of bounded angles (radians, |x| < 1e4) that the report
 * kernels build the ground-truth rotation with (csrc/pnpb200_math.cuh, sincos_bounded).
 */
int pnpb200_selftest_sincos(int64_t n, const double* in, double* sin_out, double* cos_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNPB200_H */
