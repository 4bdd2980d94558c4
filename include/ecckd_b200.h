/* ecckd_b200.h - C ABI of the B200-native (sm_100a) implementation of ecCKD's
 * data-parallel hot path.
 *
 * The reference (ecmwf-ifs/ecckd v1.6) has no FFI/plugin layer: its tool
 * mains call plain C++ free functions taking adept::Vector/Matrix arguments.
 * Each entry point below replaces one such function (or one inline hot loop of
 * a tool main) and cites it as path:line relative to the reference root.
 * A maintainer of the reference binds them as shown in INTEGRATION.md.
 *
 * Conventions
 *  - Plain pointers and sizes only; row-major arrays, last index fastest,
 *    explicit row strides (in elements) where the reference passes slices.
 *  - Names ending in _dev take DEVICE pointers for every array marked d_ and
 *    HOST pointers for the small per-layer vectors marked h_; they enqueue work
 *    on the context's stream and do not synchronise unless stated.
 *    Names without _dev take HOST pointers only (the reference-facing call):
 *    they copy in, run the same kernels, copy out and synchronise.
 *  - Every function returns an int status: 0 = OK, otherwise the reference's
 *    ESA exit codes (src/include/EsaExitCodes.h:16-51): 147 PARAMETER_ERROR,
 *    148 PROCESSING_ERROR; CUDA failures give 131 UNEXPECTED_EXCEPTION.
 *    ecckd_last_error() returns the message of the calling thread's last error.
 *    Nothing throws across this ABI.
 *  - There is no CPU fallback: every compute entry point fails with 131 when no
 *    sm_100 device is usable.
 */
#ifndef ECCKD_B200_H
#define ECCKD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECCKD_B200_VERSION 100 /* 0.1.0 */

typedef struct ecckd_ctx ecckd_ctx;
typedef struct ecckd_gpoint_plan ecckd_gpoint_plan;
typedef struct ecckd_ckd_model ecckd_ckd_model;

/* src/include/EsaExitCodes.h:16-51 */
enum ecckd_status {
  ECCKD_OK = 0,
  ECCKD_UNEXPECTED_EXCEPTION = 131,
  ECCKD_PARAMETER_ERROR = 147,
  ECCKD_PROCESSING_ERROR = 148
};

/* averaging_method strings of src/ecckd/average_optical_depth.cpp:43-129 */
enum ecckd_averaging_method {
  ECCKD_AVG_LINEAR = 0,             /* "linear" */
  ECCKD_AVG_TRANSMISSION = 1,       /* "transmission" */
  ECCKD_AVG_TRANSMISSION_2 = 2,     /* "transmission-2" */
  ECCKD_AVG_TRANSMISSION_3 = 3,     /* "transmission-3" */
  ECCKD_AVG_TRANSMISSION_10 = 4,    /* "transmission-10" */
  ECCKD_AVG_SQUARE_ROOT = 5,        /* "square-root" */
  ECCKD_AVG_LOGARITHMIC = 6,        /* "logarithmic" */
  ECCKD_AVG_HYBRID_LOG_TRANS3 = 7,  /* "hybrid-logarithmic-transmission-3" */
  ECCKD_AVG_TOTAL_TRANSMISSION = 8  /* "total-transmission": find_g_points only (find_g_points.cpp:146,334) */
};

/* Returns the enum for a reference averaging_method string, or -1
 * (the reference throws PARAMETER_ERROR, average_optical_depth.cpp:130-133). */
int ecckd_averaging_method_from_string(const char* name);

/* How the averaging weights are supplied (create_look_up_table.cpp:323-339). */
enum ecckd_weight_mode {
  ECCKD_WEIGHT_MATRIX = 0, /* d_weight[nwrow][nwav], e.g. planck_fl             */
  ECCKD_WEIGHT_VECTOR = 1, /* d_weight[nwav] for all layers: spread<0>(ssi,nlay) */
  ECCKD_WEIGHT_PLANCK = 2  /* Planck function of h_temperature[nwrow] computed in
                              the kernel (fuses planck_function.cpp:47-53)       */
};

/* per-g-point warning bits, the WARNINGs of average_optical_depth.cpp:136,154,159; NONFINITE has no
 * reference counterpart: a NaN optical depth in the g point's wavenumbers (the averaged value is NaN,
 * as the reference's sum() makes it, instead of being absorbed by the min/max clamps) */
enum ecckd_warning {
  ECCKD_WARN_EMPTY_G = 1,
  ECCKD_WARN_MIN_GT_FIT = 2,
  ECCKD_WARN_MIN_GE_MAX = 4,
  ECCKD_WARN_NONFINITE = 8
};

/* conc_dependence codes written by src/ecckd/ckd_model.cpp:409-413 */
enum ecckd_conc_dependence {
  ECCKD_CONC_NONE = 0,
  ECCKD_CONC_LINEAR = 1,
  ECCKD_CONC_LUT = 2,
  ECCKD_CONC_RELATIVE_LINEAR = 3
};

/* ------------------------------------------------------------------ context */
int ecckd_version(void);
const char* ecckd_last_error(void);
/* Number of usable sm_100 devices (0 if none or no driver). */
int ecckd_device_count(void);
/* stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL
 * for a private non-blocking stream owned by the context.
 * Devices: every entry point that takes a context (or a plan / model created on one) makes the
 * context's device current for the duration of the call and restores the caller's device on
 * return, so one process may hold contexts on several GPUs.  Side effect on the process: the
 * release threshold of the device's DEFAULT stream-ordered memory pool is raised to "keep
 * everything" (plans and partial sums come from cudaMallocAsync); set ECCKD_KEEP_POOL=0 in the
 * environment to leave the pool untouched. */
int ecckd_ctx_create(int device, void* stream, ecckd_ctx** out);
int ecckd_ctx_destroy(ecckd_ctx* ctx);
int ecckd_ctx_set_stream(ecckd_ctx* ctx, void* stream);
int ecckd_synchronize(ecckd_ctx* ctx);
int ecckd_device_malloc(ecckd_ctx* ctx, size_t bytes, void** d_ptr);
int ecckd_device_free(ecckd_ctx* ctx, void* d_ptr);
int ecckd_host_malloc_pinned(size_t bytes, void** h_ptr);
int ecckd_host_free_pinned(void* h_ptr);
int ecckd_memcpy_h2d(ecckd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int ecckd_memcpy_d2h(ecckd_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* Number of kernels this library has launched on this context so far. */
int64_t ecckd_kernel_launch_count(ecckd_ctx* ctx);
/* Measurement hook: when on, CUDA events are recorded on the context stream
 * around every launch of the streaming reduction kernel;
 * ecckd_ctx_reduce_kernel_time synchronises, returns their summed duration and
 * launch count since the last call, and clears them. */
int ecckd_ctx_set_timing(ecckd_ctx* ctx, int on);
int ecckd_ctx_reduce_kernel_time(ecckd_ctx* ctx, double* total_ms, int64_t* launches);

/* -------------------------------------------------- stage 1: spectra -> LUT */

/* Test hook (no reference counterpart): the kernels' own 1 - exp(-kd*tau) of the transmission
 * averaging methods (average_optical_depth.cpp:49,53,57,61: `1.0 - exp(-D*k*optical_depth)`) and
 * their Planck weight (planck_function.cpp:47-53 with the folded constants of the fused path),
 * evaluated element-wise on host arrays so that tests can sweep their accuracy.
 * which = 0, 2, 3: out[i] = 1 - exp(-kd*x[i]) by the 32-entry lane table, the 256-entry and the
 * replicated 64-entry shared-memory table of trans_core.cuh;  which = 1: out[i] = planck(T = kd, nu = x[i], dnu = 1). */
int ecckd_debug_eval(ecckd_ctx* ctx, int which, double kd, int64_t n, const double* x, double* out);

/* planck_function: src/ecckd/planck_function.cpp:22-54.
 * d_ans[nt][nwav] with row stride ans_stride. */
int ecckd_planck_function_dev(ecckd_ctx* ctx, int nt, const double* h_temperature, int64_t nwav,
                              const double* d_wavenumber_cm_1, const double* d_d_wavenumber_cm_1,
                              double* d_ans, int64_t ans_stride);
int ecckd_planck_function(ecckd_ctx* ctx, int nt, const double* temperature, int64_t nwav,
                          const double* wavenumber_cm_1, const double* d_wavenumber_cm_1, double* ans,
                          int64_t ans_stride);

/* The g-point map is the same for every (gas, concentration, temperature)
 * block of a LUT build (create_look_up_table.cpp:89-107 reads it once), so the
 * per-chunk g-sorted visiting order that the reduction kernel uses is computed
 * once and reused.  d_g_point[nwav] int32; entries outside [0,ng) belong to no
 * g point (they never match `g_point == ig`, average_optical_depth.cpp:40). */
int ecckd_gpoint_plan_create(ecckd_ctx* ctx, const int32_t* d_g_point, int64_t nwav, int ng,
                             ecckd_gpoint_plan** out);
int ecckd_gpoint_plan_destroy(ecckd_gpoint_plan* plan);
/* Test / diagnostic hook (no reference counterpart): the plan's rounds as the kernels read them.
 * h_rounds receives up to capacity_rounds records of 1168 bytes each {uint16 off[16][32] (byte
 * offset of (slot, lane) into the chunk), uint32 meta[32] (bits 0-11 g point, 0xFFF = lane unused;
 * bits 12-27 valid-slot mask in balanced plans, bits 12-16 valid-slot count otherwise),
 * uint32 steps, chunk, flags, win}; info[4] = {rounds, chunk length, windowed, balanced}. */
int ecckd_gpoint_plan_export(ecckd_gpoint_plan* plan, void* h_rounds, int64_t capacity_rounds, int64_t* info);
/* Number of wavenumbers in each g point (h_count[ng], host). */
int ecckd_gpoint_plan_counts(ecckd_gpoint_plan* plan, int64_t* h_count);
/* Device pointer to the same counts as int64[ng] (for a cross-GPU sum when the
 * wavenumber axis is sharded). */
int ecckd_gpoint_plan_counts_dev(ecckd_gpoint_plan* plan, int64_t** d_count);

/* average_optical_depth_to_g_point: src/ecckd/average_optical_depth.cpp:22-197
 * (declaration average_optical_depth.h:19-30).
 *   d_od[nlay][nwav] row stride od_stride;
 *   weights per weight_mode: MATRIX d_weight[nlay][nwav] row stride w_stride;
 *     VECTOR d_weight[nwav]; PLANCK h_temperature[nlay] with d_wavenumber_cm_1,
 *     d_d_wavenumber_cm_1[nwav] (d_weight ignored);
 *   outputs d_molar_abs, d_min_molar_abs, d_max_molar_abs [nlay][ng] with row
 *     stride out_stride (min/max may both be NULL), d_warn[ng] int32 bit flags
 *     (may be NULL; OR-ed into, caller zeroes).
 *   h_pressure_fl[nlay], h_pressure_hl[nlay+1] host vectors.
 * reference_surface_vmr <= 0 returns mean optical depth instead of molar
 * absorption (:186-193). */
int ecckd_average_optical_depth_to_g_point_dev(
    ecckd_ctx* ctx, ecckd_gpoint_plan* plan, double reference_surface_vmr, int nlay,
    const double* h_pressure_fl, const double* h_pressure_hl, const double* d_od, int64_t od_stride,
    int weight_mode, const double* d_weight, int64_t w_stride, const double* h_temperature,
    const double* d_wavenumber_cm_1, const double* d_d_wavenumber_cm_1, int averaging_method,
    double* d_molar_abs, double* d_min_molar_abs, double* d_max_molar_abs, int64_t out_stride,
    int32_t* d_warn);

/* Same reduction for nspec spectra that share g points, weights and pressure
 * grid (all the (gas, concentration) blocks of one temperature index of the
 * create_look_up_table.cpp:249-505 loops): one pass streams every spectrum
 * once and the weights once.  h_d_od[nspec] are device pointers to
 * [nlay][nwav] arrays (row stride od_stride); h_reference_surface_vmr[nspec];
 * h_d_molar_abs / h_d_min / h_d_max [nspec] device pointers to [nlay][ng]
 * outputs (row stride out_stride; min/max arrays may be NULL); h_d_warn[nspec]
 * device pointers to int32[ng] (array or entries may be NULL). */
int ecckd_average_optical_depth_batch_dev(
    ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nspec, const double* h_reference_surface_vmr, int nlay,
    const double* h_pressure_fl, const double* h_pressure_hl, const double* const* h_d_od,
    int64_t od_stride, int weight_mode, const double* d_weight, int64_t w_stride,
    const double* h_temperature, const double* d_wavenumber_cm_1, const double* d_d_wavenumber_cm_1,
    int averaging_method, double* const* h_d_molar_abs, double* const* h_d_min_molar_abs,
    double* const* h_d_max_molar_abs, int64_t out_stride, int32_t* const* h_d_warn);

/* The same reduction split in two for wavenumber sharding across GPUs
 * (BASELINE north_star: only the small per-g-point partial accumulators cross
 * NVLink).  Each rank runs _partial_dev on its own wavenumber range (its own
 * plan), the ranks all-reduce
 *   d_sums  [nlay][nspec][ng][3] = {sum f(tau)*w, sum w, count(tau <= 0)}  with SUM,
 *   d_minmax[nlay][nspec][ng][2] = {min tau, -max tau}                      with MIN,
 *   the plan counts int64[ng]                                               with SUM,
 * and every rank runs _finalize_dev (average_optical_depth.cpp:43-193 after the
 * sums) on the reduced arrays.  _batch_dev above is exactly _partial_dev
 * followed by _finalize_dev on one device. */
int ecckd_average_optical_depth_partial_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nspec, int nlay,
                                            const double* h_pressure_fl, const double* const* h_d_od,
                                            int64_t od_stride, int weight_mode, const double* d_weight,
                                            int64_t w_stride, const double* h_temperature,
                                            const double* d_wavenumber_cm_1, const double* d_d_wavenumber_cm_1,
                                            int averaging_method, double* d_sums, double* d_minmax);
int ecckd_average_optical_depth_finalize_dev(ecckd_ctx* ctx, int ng, const int64_t* d_count, int nspec,
                                             const double* h_reference_surface_vmr, int nlay,
                                             const double* h_pressure_fl, const double* h_pressure_hl,
                                             int averaging_method, const double* d_sums, const double* d_minmax,
                                             double* const* h_d_molar_abs, double* const* h_d_min_molar_abs,
                                             double* const* h_d_max_molar_abs, int64_t out_stride,
                                             int32_t* const* h_d_warn);

/* Host-buffer form with exactly the reference's argument list
 * (average_optical_depth.h:19-30): g_point, optical_depth and planck_fl are
 * host arrays; planck_fl row stride 0 broadcasts one row over layers
 * (spread<0>(ssi,nlay), create_look_up_table.cpp:335).  warn[ng] may be NULL. */
int ecckd_average_optical_depth_to_g_point(
    ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav,
    const double* pressure_fl, const double* pressure_hl, const int32_t* g_point,
    const double* optical_depth, int64_t od_stride, const double* planck_fl, int64_t planck_stride,
    const char* averaging_method, double* molar_abs, double* min_molar_abs, double* max_molar_abs,
    int64_t out_stride, int32_t* warn);

/* The pair planck_function + average_optical_depth_to_g_point that
 * create_look_up_table calls for every longwave block
 * (src/ecckd/create_look_up_table.cpp:320-331, 394-405, 474-485) in one host
 * call: the Planck weights of temperature_fl[nlay] are computed inside the
 * reduction kernel, so planck_fl[nlay][nwav] is never formed or transferred. */
int ecckd_planck_average_optical_depth_to_g_point(
    ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav,
    const double* pressure_fl, const double* pressure_hl, const int32_t* g_point,
    const double* optical_depth, int64_t od_stride, const double* temperature_fl,
    const double* wavenumber_cm_1, const double* d_wavenumber_cm_1, const char* averaging_method,
    double* molar_abs, double* min_molar_abs, double* max_molar_abs, int64_t out_stride, int32_t* warn);

/* One (gas, concentration, temperature) block of the create_look_up_table driver loop:
 * src/ecckd/create_look_up_table.cpp:293-344 (:368-415, :448-495).  Host arrays.  pressure_hl,
 * temperature_hl[nlay+1] are the block's profile; pressure_fl (:305) and the pressure-weighted
 * temperature_fl (:316-317, returned in temperature_fl[nlay] if not NULL) are formed here; weights
 * are the Planck function at temperature_fl (ssi == NULL, :320-331) or ssi[nwav] replicated over
 * the layers (:333-341).  The temperature loop itself is `for icol < (ncol + stride - 1) / stride`
 * over profiles icol * stride (:293-300) in the caller, writing molar_abs(icol,__,__) slices. */
int ecckd_create_lut_block(ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav,
                           const double* pressure_hl, const double* temperature_hl, const int32_t* g_point,
                           const double* optical_depth, int64_t od_stride, const double* wavenumber_cm_1,
                           const double* d_wavenumber_cm_1, const double* ssi, const char* averaging_method,
                           double* temperature_fl, double* molar_abs, double* min_molar_abs,
                           double* max_molar_abs, int64_t out_stride, int32_t* warn);

/* The same with the optical depths as the float32 values the CKDMIP spectra files hold
 * (merge_spectra.cpp:108; the reference widens them to Real on read,
 * src/tools/DataFileEngineNetcdf.cpp:393-397).  The widening is done on the device: results are
 * bit-identical to widening on the host, and the PCIe transfer is halved. */
int ecckd_planck_average_optical_depth_to_g_point_f32(
    ecckd_ctx* ctx, int ng, double reference_surface_vmr, int nlay, int64_t nwav, const double* pressure_fl,
    const double* pressure_hl, const int32_t* g_point, const float* optical_depth, int64_t od_stride,
    const double* temperature_fl, const double* wavenumber_cm_1, const double* d_wavenumber_cm_1,
    const char* averaging_method, double* molar_abs, double* min_molar_abs, double* max_molar_abs,
    int64_t out_stride, int32_t* warn);

/* The create_look_up_table block loop's state on the device (create_look_up_table.cpp:249-591): the
 * tool reads the g-point map once (:84-148), the wavenumber grid (or the solar spectrum, :74-83) with
 * the first block, and then averages every (gas, concentration, temperature) block with the SAME map
 * and grid.  A builder uploads map and grid once and keeps the map's plan; each
 * ecckd_lut_builder_block then moves only its own optical depths over PCIe.  Longwave: wavenumber
 * grid given, ssi NULL (Planck weights at the block's temperature_fl, :320-331); shortwave: ssi[nwav]
 * given (:333-341).  Arguments of _block are those of ecckd_create_lut_block minus what the builder
 * holds; exactly one of optical_depth (float64) / optical_depth_f32 (the float32 values the spectra
 * files store, merge_spectra.cpp:108, widened on the device: bit-identical results, half the bytes)
 * is non-NULL.  This is the documented host path for a sequence of blocks. */
typedef struct ecckd_lut_builder ecckd_lut_builder;
int ecckd_lut_builder_create(ecckd_ctx* ctx, int ng, int64_t nwav, const int32_t* g_point,
                             const double* wavenumber_cm_1, const double* d_wavenumber_cm_1, const double* ssi,
                             ecckd_lut_builder** out);
int ecckd_lut_builder_block(ecckd_lut_builder* builder, double reference_surface_vmr, int nlay,
                            const double* pressure_hl, const double* temperature_hl, const double* optical_depth,
                            const float* optical_depth_f32, int64_t od_stride, const char* averaging_method,
                            double* temperature_fl, double* molar_abs, double* min_molar_abs, double* max_molar_abs,
                            int64_t out_stride, int32_t* warn);
int ecckd_lut_builder_destroy(ecckd_lut_builder* builder);

/* The host-array entry points above rebuild nothing they already have: the plan of the last g-point
 * map a context saw is kept and reused when the next call passes the same map (compared element for
 * element on the device).  hits / misses since the context was created; ECCKD_PLAN_CACHE=0 disables. */
int ecckd_ctx_plan_cache_stats(ecckd_ctx* ctx, int64_t* hits, int64_t* misses);

/* ---- the cross-GPU form of the reduction -------------------------------------
 * The wavenumber axis is cut into nseg_total segments whose boundaries do not depend on the number
 * of ranks.  A rank owns the contiguous run of segments [h_seg_begin[rank], h_seg_begin[rank+1]) and
 * reduces each with its own plan, _partial_dev writing into that segment's slot of
 *   d_seg_sums  [nseg_total][n_sums]    (n_sums   = nlay*nspec*ng*3)
 *   d_seg_minmax[nseg_total][n_minmax]  (n_minmax = nlay*nspec*ng*2)
 *   d_seg_count [nseg_total][n_count]   (n_count  = ng: the plans' per-g wavenumber counts; n_count may be 0)
 * _exchange_dev makes every slot visible on every rank (one grouped in-place NCCL broadcast per
 * owner over NVLink) and combines the slots IN SEGMENT ORDER into d_sums / d_minmax / d_count, which
 * _finalize_dev then takes.  Every slot depends only on its segment's data and the combination order
 * is fixed: the result is bitwise identical on every rank and for EVERY number of ranks (an all-reduce
 * sums in an order that follows the communicator).  comm == NULL: single process, all slots are local.
 * NCCL is bound at run time (libnccl.so.2, or ECCKD_NCCL_LIB); nothing links against it. */
typedef struct ecckd_comm ecckd_comm;
/* id128: 128 bytes (ncclUniqueId); created on one rank, handed to the others by the host program */
int ecckd_comm_unique_id(void* id128);
int ecckd_comm_create(ecckd_ctx* ctx, int nranks, int rank, const void* id128, ecckd_comm** out);
/* adopt a communicator the host program already has (ncclComm_t); not destroyed by _destroy */
int ecckd_comm_from_nccl(ecckd_ctx* ctx, void* nccl_comm, int nranks, int rank, ecckd_comm** out);
int ecckd_comm_destroy(ecckd_comm* comm);
int ecckd_average_optical_depth_exchange_dev(ecckd_ctx* ctx, ecckd_comm* comm, int nseg_total,
                                             const int32_t* h_seg_begin, int64_t n_sums, int64_t n_minmax,
                                             int64_t n_count, double* d_seg_sums, double* d_seg_minmax,
                                             int64_t* d_seg_count, double* d_sums, double* d_minmax,
                                             int64_t* d_count);

/* Merge of several spectra: src/ecckd/read_merged_spectrum.cpp:149-168.
 * d_optical_depth += d_od * scaling, or * h_scaling_profile[layer] if non-NULL. */
int ecckd_merge_axpy_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, double* d_optical_depth, int64_t stride,
                         const double* d_od, int64_t od_stride, double scaling,
                         const double* h_scaling_profile);
/* The same loop for nsrc (<= 8) sources with scalar scalings in one pass: d_optical_depth = 0, then
 * += h_d_od[k] * h_scaling[k] (plain += when the scaling is 1; < 0 means 1, :146-148) in source order -
 * bitwise what the sequence of ecckd_merge_axpy_dev calls on a zeroed array gives. */
int ecckd_merge_spectra_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, int nsrc, const double* const* h_d_od,
                            const int64_t* h_od_stride, const double* h_scaling, double* d_optical_depth,
                            int64_t stride);

/* gpoint_fraction: src/ecckd/create_look_up_table.cpp:512-548.
 * Bin i is (h_wavenumber1[i], h_wavenumber2[i]]; bins must be the contiguous
 * uniform grid the reference builds (:526-535).  d_out[ng][nbin].
 * The sums are accumulated exactly (128-bit fixed point, rounded to FP64 once): the result is
 * bitwise reproducible and independent of the launch geometry. */
int ecckd_gpoint_fraction_dev(ecckd_ctx* ctx, int ng, int64_t nwav, const int32_t* d_g_point,
                              const double* d_wavenumber_cm_1, const double* d_d_wavenumber_cm_1, int nbin,
                              const double* h_wavenumber1, const double* h_wavenumber2, double* d_out);

/* Solar irradiance per bin: src/ecckd/create_look_up_table.cpp:555-561. d_out[nbin].
 * Exact sums like ecckd_gpoint_fraction_dev. */
int ecckd_ssi_intervals_dev(ecckd_ctx* ctx, int64_t nwav, const double* d_wavenumber_cm_1,
                            const double* d_ssi, int nbin, const double* h_wavenumber1,
                            const double* h_wavenumber2, double* d_out);

/* Sums of nrow per-wavenumber rows over the wavenumbers of each g point:
 * sum(x(row, find(g_point == ig))) of src/ecckd/lw_spectra.cpp:250-255 and src/ecckd/scale_lut.cpp:119-123.
 * d_values[nrow][nwav] (row stride `stride`), d_out[nrow][ng]; fixed summation order. */
int ecckd_gpoint_sum_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nrow, const double* d_values, int64_t stride,
                         double* d_out);
/* sum(x, 1): every row over all its n elements (src/ecckd/lw_spectra.cpp:230-231), fixed order. d_out[nrow]. */
int ecckd_row_sum_dev(ecckd_ctx* ctx, int nrow, int64_t n, const double* d_values, int64_t stride, double* d_out);
/* The per-profile body of the lw_spectra tool, src/ecckd/lw_spectra.cpp:208-255: Planck
 * function at the half levels, radiative_transfer_lw over every wavenumber with unit surface emissivity,
 * broadband fluxes d_flux_dn_lw / d_flux_up_lw[nlay+1], and for the plan's g points the transmission-averaged
 * optical depth d_spectral_od[nlay][ng] (raw optical depth, Planck weights at the pressure-weighted layer
 * temperatures) and the fluxes summed over each g point, d_spectral_flux_dn / _up[nlay+1][ng].
 * d_optical_depth[nlay][nwav] contiguous (od_stride == nwav); any output may be NULL. */
int ecckd_lw_spectra_profile_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, int nlay, const double* d_wavenumber_cm_1,
                                 const double* d_d_wavenumber_cm_1, const double* h_pressure_hl,
                                 const double* h_temperature_hl, const double* d_optical_depth, int64_t od_stride,
                                 double* d_flux_dn_lw, double* d_flux_up_lw, double* d_spectral_od,
                                 double* d_spectral_flux_dn, double* d_spectral_flux_up);
/* scale_lut, src/ecckd/scale_lut.cpp:117-133, 180-181, on the per-g-point sums of the direct solar flux
 * (ecckd_gpoint_sum_dev of spectral_flux_dn_direct_sw[nz+1][nwav]): h_od_best[nz][ng] = -mu0 log(flux(iz+1) /
 * flux(iz)), -1 where the flux at the layer base is not positive; h_scaling[nz][ng] (may be NULL) =
 * od_best / h_od_total, 1 where od_best <= 0.  Host arrays (they are [nz][ng] small). */
int ecckd_scale_lut_scaling(int nz, int ng, double mu0, const double* h_flux_dn_gpoint, const double* h_od_total,
                            double* h_od_best, double* h_scaling);

/* Planck look-up table per g point: src/ecckd/create_look_up_table.cpp:579-591.
 * d_out[nlut][ng] = sum over the g point's wavenumbers of planck(T_lut, nu). */
int ecckd_planck_lut_dev(ecckd_ctx* ctx, ecckd_gpoint_plan* plan, const double* d_wavenumber_cm_1,
                         const double* d_d_wavenumber_cm_1, int nlut, const double* h_temperature_lut,
                         double* d_out);

/* Remove g points that occupy none of the spectrum, renumbering in place:
 * src/ecckd/create_look_up_table.cpp:111-160.  h_g_point_map[>=ng] (host)
 * receives the old index of each kept g point; *new_ng the new count.
 * Returns 1 (the reference's THROW(1), :148) if a point is left unassigned. */
int ecckd_compact_g_points_dev(ecckd_ctx* ctx, int ng, int64_t nwav, int32_t* d_g_point,
                               int32_t* h_g_point_map, int* new_ng);

/* ------------------------------------------ ranking / g-point assignment */

/* Sorting variable of reorder_spectrum: LW pseudo-height of peak cooling
 * (src/ecckd/reorder_spectrum.cpp:111-190) and SW height where the optical
 * depth from TOA reaches the threshold (:197-228).  d_od[nlay][nwav];
 * outputs d_key[nwav], d_column_od[nwav] (may be NULL).  The SW form returns
 * PROCESSING_ERROR where the reference throws (key > 30, :214-216). */
int ecckd_sort_key_lw_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, const double* h_pressure_hl,
                          const double* d_wavenumber_cm_1, const double* d_d_wavenumber_cm_1,
                          const double* d_od, int64_t od_stride, double threshold_optical_depth, double* d_key,
                          double* d_column_od);
int ecckd_sort_key_sw_dev(ecckd_ctx* ctx, int nlay, int64_t nwav, const double* h_pressure_hl,
                          const double* d_od, int64_t od_stride, double threshold_optical_depth, double* d_key,
                          double* d_column_od);

/* Band membership and per-band std::stable_sort of the sorting variable:
 * src/ecckd/reorder_spectrum.cpp:30-34,262-300.  Outputs (any may be NULL)
 * d_iband[nwav], d_rank[nwav], d_ordered_index[nwav], all int32.  Bit-exact
 * with std::stable_sort(key[i] < key[j]) for NaN-free keys. */
int ecckd_rank_bands_dev(ecckd_ctx* ctx, int64_t nwav, const double* d_wavenumber_cm_1, const double* d_key,
                         int nband, const double* h_band_bound1, const double* h_band_bound2,
                         int32_t* d_iband, int32_t* d_rank, int32_t* d_ordered_index);

/* SingleGasData::store_g_points: src/ecckd/single_gas_data.h:56-62. */
int ecckd_store_g_points_dev(ecckd_ctx* ctx, int64_t nwav, const int32_t* d_rank, int ng,
                             const int32_t* h_rank1, const int32_t* h_rank2, int32_t* d_g_point);

/* overlap_g_points (hypercube partition): src/ecckd/single_gas_data.cpp:23-124.
 * Host integer logic over a few dozen entries; all pointers are HOST.
 * h_n_g_points[ngas][nband]; h_sorting_variable is the concatenation of each
 * gas's per-g array with h_gas_offset[ngas] the start of each slice;
 * h_g_min/h_g_max[ngas][cap], h_band_number[cap].  *ng_out = merged count. */
int ecckd_overlap_g_points(int ngas, int nband, const int32_t* h_n_g_points, const double* h_sorting_variable,
                           const int32_t* h_gas_offset, int cap, int32_t* h_g_min, int32_t* h_g_max,
                           int32_t* h_band_number, int* ng_out);

/* Final multi-gas assignment: src/ecckd/find_g_points.cpp:1459-1475.
 * d_gas_g_point[ngas][nwav]; h_g_min/h_g_max[ngas][cap] (host);
 * d_g_point[nwav]; *h_n_unassigned (host, may be NULL) counts -1 entries. */
int ecckd_assign_g_points_dev(ecckd_ctx* ctx, int ngas, int64_t nwav, const int32_t* d_gas_g_point, int ng,
                              int cap, const int32_t* h_g_min, const int32_t* h_g_max, int32_t* d_g_point,
                              int64_t* h_n_unassigned);

/* ------------------------------------------------ stage 2: CKD flux solver */

/* One gas of a CKD model as stored by src/ecckd/ckd_model.cpp:288-641.
 * All pointers HOST; copied to the device by ecckd_ckd_model_create. */
typedef struct ecckd_gas_table {
  int32_t conc_dependence;  /* enum ecckd_conc_dependence */
  int32_t nconc;            /* LUT only: number of mole fractions */
  const double* vmr_lut;    /* LUT only: [nconc] */
  double reference_vmr;     /* relative-linear only */
  const double* molar_abs;  /* [nt][np][ng], or [nconc][nt][np][ng] for LUT */
} ecckd_gas_table;

/* LW: planck_lut[nlut][ng] with temperature_planck[nlut]; solar_irradiance and
 * rayleigh_molar_scat NULL.  SW: solar_irradiance[ng], rayleigh_molar_scat[ng];
 * planck tables NULL/0.  log_pressure[np] (evenly spaced), temperature[nt][np]. */
int ecckd_ckd_model_create(ecckd_ctx* ctx, int ngas, const ecckd_gas_table* gases, int nt, int np, int ng,
                           const double* log_pressure, const double* temperature, int nlut,
                           const double* temperature_planck, const double* planck_lut,
                           const double* solar_irradiance, const double* rayleigh_molar_scat,
                           ecckd_ckd_model** out);
int ecckd_ckd_model_destroy(ecckd_ckd_model* model);

/* CkdModel::calc_optical_depth: src/ecckd/ckd_model.cpp:923-1102.
 * d_pressure_hl[ncol][nlay+1], d_temperature_fl[ncol][nlay], d_vmr_fl[ncol][nlay]
 * (NULL = "not provided"); d_od[ncol][nlay][ng] is overwritten
 * (accumulate = 0) or added to (accumulate != 0, run_ckd.cpp:312). */
int ecckd_ckd_calc_optical_depth_dev(ecckd_ctx* ctx, ecckd_ckd_model* model, int igas, int ncol, int nlay,
                                     const double* d_pressure_hl, const double* d_temperature_fl,
                                     const double* d_vmr_fl, double* d_od, int accumulate);

/* CkdModel::calc_planck_function: src/ecckd/ckd_model.cpp:1105-1145.
 * d_temperature[n] -> d_planck[n][ng]. */
int ecckd_ckd_calc_planck_function_dev(ecckd_ctx* ctx, ecckd_ckd_model* model, int64_t n,
                                       const double* d_temperature, double* d_planck);

/* CkdModel::calc_rayleigh_optical_depth: src/ecckd/ckd_model.h:242-252. */
int ecckd_ckd_calc_rayleigh_optical_depth_dev(ecckd_ctx* ctx, ecckd_ckd_model* model, int ncol, int nlay,
                                              const double* d_pressure_hl, double* d_od);

/* radiative_transfer_lw<false>: src/ecckd/radiative_transfer_lw.cpp:25-60,
 * batched over ncol independent columns (ncol = 1 is the reference call).
 * d_planck[ncol][nlay+1][n], d_od[ncol][nlay][n], d_surf_emissivity[n] (shared
 * by all columns, run_ckd.cpp:344-345), d_surf_planck[ncol][n],
 * d_flux_dn/up[ncol][nlay+1][n]. */
int ecckd_radiative_transfer_lw_dev(ecckd_ctx* ctx, int ncol, int nlay, int64_t n, const double* d_planck,
                                    const double* d_od, const double* d_surf_emissivity,
                                    const double* d_surf_planck, double* d_flux_dn, double* d_flux_up);

/* radiative_transfer_direct_sw<false> (d_albedo, d_flux_up NULL):
 * src/ecckd/radiative_transfer_sw.cpp:24-43; radiative_transfer_norayleigh_sw
 * <false> (both non-NULL): :47-77.  d_ssi[n], d_albedo[n] shared by columns;
 * h_cos_sza[ncol] per column (host). */
int ecckd_radiative_transfer_sw_dev(ecckd_ctx* ctx, int ncol, int nlay, int64_t n, const double* h_cos_sza,
                                    const double* d_ssi, const double* d_od, const double* d_albedo,
                                    double* d_flux_dn, double* d_flux_up);

/* heating_rate<false>: src/ecckd/heating_rate.h:25-49, batched over columns.
 * d_pressure_hl[ncol][nlay+1]; d_flux_up may be NULL; d_hr[ncol][nlay][n]. */
int ecckd_heating_rate_dev(ecckd_ctx* ctx, int ncol, int nlay, int64_t n, const double* d_pressure_hl,
                           const double* d_flux_dn, const double* d_flux_up, double* d_hr);

/* Numerical core of run_ckd: src/ecckd/run_ckd.cpp:119-132,261-367 in one
 * fused pass per column x g point: full-level temperature, LUT interpolation of
 * every gas, sum, clamp at zero, Planck interpolation (LW) or Rayleigh +
 * incoming solar (SW, mu0 = 0.5), flux sweep, sums over g.
 * h_d_vmr_fl[ngas]: device pointers [ncol][nlay] per gas, NULL entry = variable
 * absent from the input file (:277-281).  Outputs (device, any may be NULL):
 * d_od[ncol][nlay][ng], d_planck_hl[ncol][nlay+1][ng], d_planck_surf[ncol][ng],
 * d_flux_dn, d_flux_up[ncol][nlay+1][ng] (SW: d_flux_dn = direct flux, d_flux_up
 * ignored), d_bb_flux_dn, d_bb_flux_up[ncol][nlay+1]. */
int ecckd_run_ckd_dev(ecckd_ctx* ctx, ecckd_ckd_model* model, const double* const* h_d_vmr_fl, double tsi,
                      int ncol, int nlay, const double* d_pressure_hl, const double* d_temperature_hl,
                      double* d_od, double* d_planck_hl, double* d_planck_surf, double* d_flux_dn,
                      double* d_flux_up, double* d_bb_flux_dn, double* d_bb_flux_up);
/* Host-buffer form: same arguments, all arrays on the host. */
int ecckd_run_ckd(ecckd_ctx* ctx, ecckd_ckd_model* model, const double* const* vmr_fl, double tsi, int ncol,
                  int nlay, const double* pressure_hl, const double* temperature_hl, double* od,
                  double* planck_hl, double* planck_surf, double* flux_dn, double* flux_up,
                  double* bb_flux_dn, double* bb_flux_up);

/* ---------------------------------------------------------------------------
 * Stage-1 extras
 * ------------------------------------------------------------------------- */

/* Base g-point split: src/ecckd/create_look_up_table.cpp:162-224.  For every band that contains
 * base_wavenumber_boundary values strictly inside (band_wn1, band_wn2), the first g point of
 * the band is cut at those wavenumbers; higher g points are renumbered.  In/out: *ng_io,
 * h_band_number[cap] (host), d_g_point[nwav] (device), h_solar_irradiance[cap] (host; NULL
 * together with d_ssi in the longwave).  PROCESSING_ERROR if cap < the new number of g points. */
int ecckd_split_base_g_points_dev(ecckd_ctx* ctx, int nband, const double* h_band_wn1, const double* h_band_wn2,
                                  int nbound, const double* h_base_wavenumber_boundary, int64_t nwav,
                                  const double* d_wavenumber_cm_1, const double* d_ssi, int cap, int* ng_io,
                                  int32_t* h_band_number, int32_t* d_g_point, double* h_solar_irradiance);

/* rayleigh_molar_scattering_coeff: src/ecckd/rayleigh_scattering.h:23-43 (m2 mol-1). */
int ecckd_rayleigh_molar_scattering_coeff_dev(ecckd_ctx* ctx, int64_t nwav, const double* d_wavenumber_cm_1,
                                              double* d_scat_coeff);

/* CkdModel::calc_rayleigh_molar_scat: src/ecckd/ckd_model.h:368-385.
 * d_wavenumber1/2[nbin], d_gpoint_fraction[ng][nbin], d_ssi_intervals[nbin] -> [ng]. */
int ecckd_rayleigh_molar_scat_gpoints_dev(ecckd_ctx* ctx, int ng, int nbin, const double* d_wavenumber1,
                                          const double* d_wavenumber2, const double* d_gpoint_fraction,
                                          const double* d_ssi_intervals, double* d_rayleigh_molar_scat);

/* ---------------------------------------------------------------------------
 * Broadband solvers and the candidate-interval error of find_g_points (rows a13, a15)
 * ------------------------------------------------------------------------- */

/* radiative_transfer_lw_bb<false>: src/ecckd/radiative_transfer_lw.cpp:85-142.
 * d_planck[nlay+1][n], d_spectral_od[nlay][n] (row stride `stride`), d_grey_od[nlay],
 * d_surf_emissivity[n], d_surf_planck[n] -> d_flux_dn, d_flux_up[nlay+1] (sums over n). */
int ecckd_radiative_transfer_lw_bb_dev(ecckd_ctx* ctx, int nlay, int64_t n, int64_t stride, const double* d_planck,
                                       const double* d_spectral_od, const double* d_grey_od,
                                       const double* d_surf_emissivity, const double* d_surf_planck,
                                       double* d_flux_dn, double* d_flux_up);

/* radiative_transfer_direct_sw_bb<false> (d_flux_up NULL): src/ecckd/radiative_transfer_sw.cpp:116-141;
 * radiative_transfer_norayleigh_sw_bb<false>: :145-184 (scalar albedo). */
int ecckd_radiative_transfer_sw_bb_dev(ecckd_ctx* ctx, int nlay, int64_t n, int64_t stride, double cos_sza,
                                       const double* d_ssi, const double* d_spectral_od, const double* d_grey_od,
                                       double albedo, double* d_flux_dn, double* d_flux_up);

/* CkdEquipartition::lower_index / upper_index and the bound checks of calc_error:
 * src/ecckd/find_g_points.cpp:284-315.  bound1/2[nint] are fractions of the band's rank axis;
 * returns PROCESSING_ERROR (148) exactly where the reference throws it. */
int ecckd_interval_bounds_to_indices(int64_t npoints, int nint, const double* bound1, const double* bound2,
                                     int64_t* i1, int64_t* i2);

/* CkdEquipartition::calc_error for a BATCH of candidate intervals (the reference calls it from an
 * OpenMP loop, equipartition.h:100-105): fit_optical_depth_lw (find_g_points.cpp:54-116, methods
 * linear / transmission / transmission-2 / square-root / logarithmic) followed by
 * calc_cost_function_lw (calc_cost_function_lw.cpp:23-110).  Spectral arrays are the band's
 * slices in rank order: d_planck_hl[nlay+1][n], d_bg_optical_depth, d_metric, d_hr[nlay][n]
 * (row stride `stride`), d_surf_emissivity, d_surf_planck, d_flux_dn_surf, d_flux_up_toa[n],
 * d_pressure_hl[nlay+1], d_layer_weight[nlay].  h_i1/h_i2[nint]: inclusive index bounds (host).
 * Outputs: d_error[nint]; d_optical_depth_fit[nint][nlay] (may be NULL). */
int ecckd_interval_error_batch_lw_dev(ecckd_ctx* ctx, int averaging_method, int nlay, int64_t n, int64_t stride,
                                      const double* d_pressure_hl, const double* d_planck_hl,
                                      const double* d_surf_emissivity, const double* d_surf_planck,
                                      const double* d_bg_optical_depth, const double* d_metric,
                                      const double* d_flux_dn_surf, const double* d_flux_up_toa, const double* d_hr,
                                      double flux_weight, const double* d_layer_weight, int nint,
                                      const int64_t* h_i1, const int64_t* h_i2, double* d_error,
                                      double* d_optical_depth_fit);

/* Shortwave form: fit_optical_depth_sw (find_g_points.cpp:119-173) or, for
 * ECCKD_AVG_TOTAL_TRANSMISSION, fit_optical_depth_sw_total_trans (:176-204) with the error taken as
 * the mean of calc_cost_function_sw (calc_cost_function_sw.cpp:20-110) against the *_low set at
 * min_scaling and the *_high set at max_scaling (find_g_points.cpp:334-384); the plain reference
 * set may then be NULL, otherwise the low/high sets may be NULL. */
int ecckd_interval_error_batch_sw_dev(
    ecckd_ctx* ctx, int averaging_method, int nlay, int64_t n, int64_t stride, double cos_sza,
    const double* d_pressure_hl, const double* d_ssi, double albedo, const double* d_bg_optical_depth,
    const double* d_metric, const double* d_flux_dn_surf, const double* d_flux_up_toa, const double* d_hr,
    double flux_weight, const double* d_layer_weight, const double* d_flux_dn_surf_low,
    const double* d_flux_up_toa_low, const double* d_hr_low, const double* d_flux_dn_surf_high,
    const double* d_flux_up_toa_high, const double* d_hr_high, double min_scaling, double max_scaling, int nint,
    const int64_t* h_i1, const int64_t* h_i2, double* d_error, double* d_optical_depth_fit);

/* ---------------------------------------------------------------------------
 * Partition search of find_g_points: the Equipartition class of the reference
 * (src/ecckd/equipartition.h:64-209, equipartition.cpp:160-805) as a host driver over a BATCHED interval
 * error evaluator.  Every calc_error_all of the reference (equipartition.h:97-113, one calc_error per
 * OpenMP thread) is one evaluator call with all its intervals; the candidates of a line search
 * (equipartition.cpp:160-196: the proposal moved 0, 1, 2, ... times halfway back) do not depend on the
 * errors, so `speculative_line_search` of them go into one batch.  Given an evaluator returning the
 * same errors, bounds / errors / number of intervals / status are those of the reference, bit for bit.
 * ------------------------------------------------------------------------- */
enum ecckd_ep_status { /* EpStatus, equipartition.h:31-39 */
  ECCKD_EP_SUCCESS = 0,
  ECCKD_EP_MAX_ITERATIONS_REACHED = 1,
  ECCKD_EP_FAILED_TO_CONVERGE = 2,
  ECCKD_EP_RESOLUTION_LIMIT_REACHED = 3,
  ECCKD_EP_NO_PROGRESS = 4,
  ECCKD_EP_FAILURE = 5,
  ECCKD_EP_INPUT_ERROR = 6
};
/* error[i] of the interval [bound1[i], bound2[i]], i < n; returns ECCKD_OK or the code to abort with */
typedef int (*ecckd_interval_error_fn)(void* user, int n, const double* bound1, const double* bound2, double* error);
typedef struct ecckd_partition_options { /* defaults: equipartition.h:187-202 */
  double partition_tolerance;        /* 0.05; find_g_points sets it (find_g_points.cpp:1180) */
  double next_bound_error_tolerance; /* 0.05 */
  double resolution;                 /* 0; find_g_points: 1 / npoints (:230, :258) */
  int partition_max_iterations;      /* 20 (:1179) */
  int line_search_max_iterations;    /* 10 */
  int next_bound_max_iterations;     /* 20 */
  int minimize_frac_range;           /* 1: (max - min) / mean; 0: fractional standard deviation */
  int cubic_interpolation;           /* 0 */
  int speculative_line_search;       /* line-search candidates per evaluator batch (>= 1; 1 = one at a time) */
} ecckd_partition_options;
void ecckd_partition_options_default(ecckd_partition_options* options);
typedef struct ecckd_partition_search ecckd_partition_search;
int ecckd_partition_search_create(const ecckd_partition_options* options /* NULL: defaults */, ecckd_interval_error_fn fn,
                                  void* user, ecckd_partition_search** out);
int ecckd_partition_search_destroy(ecckd_partition_search* search);
/* Equipartition::equipartition_n (equipartition.cpp:353-562): bounds[ni+1] in/out (ends fixed), error[ni] out
 * (in as well when errors_given != 0, as after equipartition_e); *ep_status one of ecckd_ep_status.
 * Returns ECCKD_OK, or what the evaluator returned when it failed. */
int ecckd_partition_search_n(ecckd_partition_search* search, int ni, double* bounds, double* error, int errors_given,
                             int* ep_status);
/* Equipartition::equipartition_e (:569-623): as many intervals as a target error needs, then equal errors.
 * bounds[capacity+1], error[capacity]; *ni out. */
int ecckd_partition_search_e(ecckd_partition_search* search, double target_error, double bound0, double boundn,
                             int capacity, int* ni, double* bounds, double* error, int* ep_status);
/* One band of find_g_points without sub-bands (find_g_points.cpp:1229-1256, 1396-1400): equipartition_e on
 * [0, 1], re-partitioned into min/max_g_points intervals from the bounds sqrt(i / ng) when the count falls
 * outside that range; rank1/rank2[ng] (may be NULL) = the band-relative index range of every g point. */
int ecckd_partition_search_band(ecckd_partition_search* search, int64_t npoints, double target_error, int min_g_points,
                                int max_g_points, int capacity, int* ng, double* bounds, double* error, int64_t* rank1,
                                int64_t* rank2, int* ep_status);
/* evaluator calls, intervals evaluated, intervals evaluated speculatively and not needed */
int ecckd_partition_search_stats(ecckd_partition_search* search, int64_t* batches, int64_t* intervals,
                                 int64_t* speculative_unused);

/* The evaluators of find_g_points: CkdEquipartition::calc_error (find_g_points.cpp:291-405) on a batch, i.e.
 * ecckd_interval_bounds_to_indices + ecckd_interval_error_batch_{lw,sw}_dev + the read-back of the errors.
 * Pass the function as `fn` and a filled argument block as `user` (fields: the arguments of the batch calls). */
typedef struct ecckd_interval_error_lw_args {
  ecckd_ctx* ctx;
  int averaging_method, nlay;
  int64_t n, stride;
  const double *d_pressure_hl, *d_planck_hl, *d_surf_emissivity, *d_surf_planck, *d_bg_optical_depth, *d_metric,
      *d_flux_dn_surf, *d_flux_up_toa, *d_hr;
  double flux_weight;
  const double* d_layer_weight;
} ecckd_interval_error_lw_args;
typedef struct ecckd_interval_error_sw_args {
  ecckd_ctx* ctx;
  int averaging_method, nlay;
  int64_t n, stride;
  double cos_sza;
  const double *d_pressure_hl, *d_ssi;
  double albedo;
  const double *d_bg_optical_depth, *d_metric, *d_flux_dn_surf, *d_flux_up_toa, *d_hr;
  double flux_weight;
  const double *d_layer_weight, *d_flux_dn_surf_low, *d_flux_up_toa_low, *d_hr_low, *d_flux_dn_surf_high,
      *d_flux_up_toa_high, *d_hr_high;
  double min_scaling, max_scaling;
} ecckd_interval_error_sw_args;
int ecckd_interval_error_lw_eval(void* lw_args, int n, const double* bound1, const double* bound2, double* error);
int ecckd_interval_error_sw_eval(void* sw_args, int n, const double* bound1, const double* bound2, double* error);

/* ---------------------------------------------------------------------------
 * Forward cost of optimize_lut (rows a21-a23), batched over the profiles of a scene
 * ------------------------------------------------------------------------- */

/* calc_total_optical_depth: src/ecckd/solve_adept.cpp:23-69.  d_od[ncol][nlay][ng] = Rayleigh
 * optical depth (shortwave models) + every gas of the model; h_d_vmr_fl[ngas] holds device
 * pointers [ncol][nlay] or NULL for a gas the training scene does not hold: such a gas is
 * included without concentration if it has no concentration dependence and skipped otherwise. */
int ecckd_calc_total_optical_depth_dev(ecckd_ctx* ctx, ecckd_ckd_model* model, const double* const* h_d_vmr_fl,
                                       int ncol, int nlay, const double* d_pressure_hl,
                                       const double* d_temperature_hl, double* d_od);

/* Adjoint of the above: the optical depth is linear in the look-up tables, and
 * h_d_grad_molar_abs[igas] (device, laid out like the gas's table: [nconc][nt][np][ng] with a concentration axis,
 * [nt][np][ng] without; NULL: not wanted) receives sum over (column, layer) of d_grad_od[ncol][nlay][ng] times the
 * interpolation coefficient of that table entry (ckd_model.cpp:962-1088 transposed) - with d_grad_od from
 * ecckd_cost_function_ckd_lw_grad_dev this is dJ/d(molar absorption), what the Adept tape of solve_adept.cpp:91,
 * 201-203 yields for the state vector of optimize_lut.  A gather per table node in a fixed order: bitwise
 * reproducible, no floating-point atomics.  accumulate != 0 adds to what the arrays hold (the loop over training
 * scenes, solve_adept.cpp:96).  Gases the scene does not hold are skipped as in the forward call. */
int ecckd_calc_total_optical_depth_adjoint_dev(ecckd_ctx* ctx, ecckd_ckd_model* model, const double* const* h_d_vmr_fl,
                                               int ncol, int nlay, const double* d_pressure_hl,
                                               const double* d_temperature_hl, const double* d_grad_od,
                                               double* const* h_d_grad_molar_abs, int accumulate);

/* calc_cost_function_ckd_lw (forward value): src/ecckd/calc_cost_function_lw.cpp:115-232, for
 * every profile of a training scene as solve_adept.cpp:96-184 drives it, including the layer
 * weights (d_layer_weight[nprof][nlay], or NULL: computed from pressure_weight_power and
 * normalised, solve_adept.cpp:133-145) and the negative-optical-depth penalty and clamp
 * (solve_adept.cpp:107-116; the input array is not modified).
 * d_planck_hl[nprof][nlay+1][ng], d_surf_emissivity[nprof][nband] (or [ng] without band_mapping),
 * d_surf_planck[nprof][ng], d_optical_depth[nprof][nlay][ng], truth d_flux_dn/up[nprof][nlay+1][nband],
 * d_hr[nprof][nlay][nband], optional d_spectral_flux_*[nprof][ng], d_relative_ckd_flux_*[nprof][nlay+1][ng],
 * d_band_mapping[ng] (NULL: nband == ng).  Outputs (device, any may be NULL): d_cost_profile[nprof],
 * d_cost_total[1] = penalty + sum of the profile costs in profile order, d_penalty_total[1]. */
int ecckd_cost_function_ckd_lw_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_pressure_hl, const double* d_planck_hl,
    const double* d_surf_emissivity, const double* d_surf_planck, const double* d_optical_depth,
    const double* d_flux_dn, const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf,
    const double* d_spectral_flux_up_toa, double flux_weight, double flux_profile_weight, double broadband_weight,
    double spectral_boundary_weight, const double* d_layer_weight, double pressure_weight_power,
    double negative_od_penalty, const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up,
    const int32_t* d_band_mapping, double* d_cost_profile, double* d_cost_total, double* d_penalty_total);

/* Reverse mode of the above (first part): d_grad_optical_depth[nprof][nlay][ng] = the
 * derivative of (sum over the profiles of calc_cost_function_ckd_lw + the negative-optical-depth penalty of
 * solve_adept.cpp:107-116) with respect to d_optical_depth - what the Adept tape of solve_adept.cpp:91, 201-203
 * yields up to the look-up-table interpolation, written by hand: adjoints of the quadratic cost terms, of the
 * heating-rate stencil (heating_rate.h:38-48) and of the two flux recurrences of radiative_transfer_lw.cpp:25-60.
 * Same arguments as the forward call.  Clamped (negative) optical depths carry only the penalty's derivative. */
int ecckd_cost_function_ckd_lw_grad_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_pressure_hl, const double* d_planck_hl,
    const double* d_surf_emissivity, const double* d_surf_planck, const double* d_optical_depth,
    const double* d_flux_dn, const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf,
    const double* d_spectral_flux_up_toa, double flux_weight, double flux_profile_weight, double broadband_weight,
    double spectral_boundary_weight, const double* d_layer_weight, double pressure_weight_power,
    double negative_od_penalty, const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up,
    const int32_t* d_band_mapping, double* d_grad_optical_depth);

/* calc_cost_function_ckd_sw (forward value): src/ecckd/calc_cost_function_sw.cpp:115-277.
 * d_cos_sza[nprof], d_ssi[ng] (already scaled to the scene's TSI, solve_adept.cpp:160-166),
 * h_albedo[nalbedo] indexed by band_mapping; d_cost_fn_per_band[nband] (may be NULL) is added to. */
int ecckd_cost_function_ckd_sw_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_cos_sza, const double* d_pressure_hl,
    const double* d_ssi, int nalbedo, const double* h_albedo, const double* d_optical_depth, const double* d_flux_dn,
    const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf,
    const double* d_spectral_flux_up_toa, double flux_weight, double flux_profile_weight, double broadband_weight,
    const double* d_spectral_boundary_weights, const double* d_layer_weight, double pressure_weight_power,
    double negative_od_penalty, const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up,
    const int32_t* d_band_mapping, double* d_cost_fn_per_band, double* d_cost_profile, double* d_cost_total,
    double* d_penalty_total);

/* Reverse mode of the above: d_grad_optical_depth[nprof][nlay][ng] = derivative of (sum over the profiles of
 * calc_cost_function_ckd_sw, src/ecckd/calc_cost_function_sw.cpp:115-277, + the negative-optical-depth penalty) with
 * respect to d_optical_depth: adjoints of the quadratic terms (the 20-fold TOA upwelling term, the broadband part only
 * when broadband_weight > 0, the upwelling in it only when every albedo is positive), of the direct-beam heating rate
 * and of the two exponential recurrences of radiative_transfer_sw.cpp:24-77. */
int ecckd_cost_function_ckd_sw_grad_dev(
    ecckd_ctx* ctx, int nprof, int nlay, int ng, int nband, const double* d_cos_sza, const double* d_pressure_hl,
    const double* d_ssi, int nalbedo, const double* h_albedo, const double* d_optical_depth, const double* d_flux_dn,
    const double* d_flux_up, const double* d_hr, const double* d_spectral_flux_dn_surf, double flux_weight,
    double flux_profile_weight, double broadband_weight, const double* d_spectral_boundary_weights,
    const double* d_layer_weight, double pressure_weight_power, double negative_od_penalty,
    const double* d_relative_ckd_flux_dn, const double* d_relative_ckd_flux_up, const int32_t* d_band_mapping,
    double* d_grad_optical_depth);

/* The forward cost of several scenes of optimize_lut's training set in ONE launch: the loop over
 * LblFluxes objects of src/ecckd/solve_adept.cpp:96-199 (each iteration = one call of the two entry
 * points above).  A scene's kernel has only nprof CTAs (50-150) and is latency-bound on its own; the
 * batch fills the GPU.  One struct per scene, its fields the per-scene arguments of
 * ecckd_cost_function_ckd_lw_dev / _sw_dev (device pointers except h_albedo; the fields of the other
 * spectral region are ignored); nlay, ng, nband, the weights and d_band_mapping are shared by the
 * scenes, as they are in solve_adept.  Outputs (device, any may be NULL): d_cost_scene[nscene] =
 * penalty + sum of the scene's profile costs in profile order, d_penalty_scene[nscene],
 * d_cost_total[1] = sum of d_cost_scene in scene order, d_cost_fn_per_band[nband] (shortwave) added to. */
typedef struct ecckd_cost_scene {
  int32_t nprof;
  int32_t nalbedo;                           /* shortwave */
  const double* d_pressure_hl;               /* [nprof][nlay+1] */
  const double* d_optical_depth;             /* [nprof][nlay][ng] */
  const double* d_flux_dn;                   /* [nprof][nlay+1][nband] */
  const double* d_flux_up;                   /* [nprof][nlay+1][nband] */
  const double* d_hr;                        /* [nprof][nlay][nband] */
  const double* d_spectral_flux_dn_surf;     /* [nprof][ng] or NULL */
  const double* d_spectral_flux_up_toa;      /* [nprof][ng] or NULL (longwave) */
  const double* d_layer_weight;              /* [nprof][nlay] or NULL */
  const double* d_relative_ckd_flux_dn;      /* [nprof][nlay+1][ng] or NULL */
  const double* d_relative_ckd_flux_up;
  const double* d_planck_hl;                 /* longwave [nprof][nlay+1][ng] */
  const double* d_surf_emissivity;           /* longwave [nprof][nband or ng] */
  const double* d_surf_planck;               /* longwave [nprof][ng] */
  const double* d_cos_sza;                   /* shortwave [nprof] */
  const double* d_ssi;                       /* shortwave [ng] */
  const double* h_albedo;                    /* shortwave [nalbedo], host */
  const double* d_spectral_boundary_weights; /* shortwave [ng] or NULL */
  double* d_cost_profile;                    /* [nprof] out, or NULL */
} ecckd_cost_scene;
int ecckd_cost_function_ckd_scenes_dev(ecckd_ctx* ctx, int is_sw, int nscene, const ecckd_cost_scene* scenes, int nlay,
                                       int ng, int nband, double flux_weight, double flux_profile_weight,
                                       double broadband_weight, double spectral_boundary_weight,
                                       double pressure_weight_power, double negative_od_penalty,
                                       const int32_t* d_band_mapping, double* d_cost_fn_per_band, double* d_cost_scene,
                                       double* d_penalty_scene, double* d_cost_total);

/* Background (prior) term of the optimize_lut cost function for one active gas:
 * CkdModel<true>::calc_background_cost_function, src/ecckd/ckd_model.cpp:838-866.  d_delta_x and d_gradient are the
 * gas's [nx][ng] slice of the state vector and of its gradient (g point fastest, :851-855), d_inv_background_shape
 * the gas's [nx][nx] inverse background shape matrix (row-major), d_background_error [ng].  Per g point:
 * gradient_local = (1 / background_error(g)^2) * (inv_background_shape ** delta_x_local) (:857-858);
 * d_gradient += gradient_local (:861; NULL = cost only; the caller zeroes the gradient once before the first gas as
 * :842 does); *h_cost = sum over g of 0.5 * dot_product(delta_x_local, gradient_local) (:859,864), fixed order. */
int ecckd_background_cost_gas_dev(ecckd_ctx* ctx, int nx, int ng, const double* d_inv_background_shape,
                                  const double* d_background_error, const double* d_delta_x, double* d_gradient,
                                  double* h_cost);

/* The Rayleigh part of the same function (ckd_model.cpp:868-873): *h_cost = 0.5 * sum_g delta_x(g) *
 * (rayleigh_inv_background(g) * delta_x(g)).  As in the reference, it contributes to the cost only: the reference
 * forms gradient_local and does not add it to the gradient. */
int ecckd_background_cost_rayleigh_dev(ecckd_ctx* ctx, int ng, const double* d_rayleigh_inv_background,
                                       const double* d_delta_x, double* h_cost);

#ifdef __cplusplus
}
#endif
#endif /* ECCKD_B200_H */
