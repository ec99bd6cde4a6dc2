/* ventb200 -- C ABI of the B200-native Vent_Analysis voxel hot path (N4 -> VDP -> CI).
 *
 * Plain C: pointers, sizes, status codes.  No torch / numpy / C++ types cross this
 * boundary.  The reference (thomenr/Vent_Analysis) is pure Python and has no FFI of its
 * own; each entry point below names the reference function whose arithmetic it replaces,
 * and INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   * A batch is `n` subjects.  Subject i is a C-ordered [rows][cols][slices] volume
 *     (exactly numpy's layout of the reference's arrays) with shape shapes[3*i .. 3*i+2];
 *     all per-voxel arrays are the subjects' volumes concatenated in batch order.
 *   * `_dev` entry points take DEVICE pointers and a cudaStream_t (as void*), enqueue work
 *     and return without synchronising; `shapes` / `table_handles` are HOST arrays.
 *   * `_host` entry points take HOST pointers, copy in, run, copy out and synchronise.
 *   * Every function returns VB_OK or a VB_ERR_* code; vb_last_error() gives the text
 *     (thread local).  Nothing falls back to the CPU: without a CUDA device vb_create fails.
 */
#ifndef VENTB200_H
#define VENTB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VB_OK 0
#define VB_ERR_CUDA 1        /* CUDA runtime / launch failure                              */
#define VB_ERR_ARG 2         /* bad argument                                               */
#define VB_ERR_UNSUPPORTED 3 /* size outside what the kernels were built for               */
#define VB_ERR_CI_RMAX 4     /* a defect voxel exhausted the sphere table: the reference
                                raises ValueError there (CI.py:101-103)                    */

#define VB_DO_SNR 1u         /* Vent_Analysis.calculate_SNR        Vent_Analysis.py:337-357 */
#define VB_DO_N4 2u          /* Vent_Analysis.N4_bias_correction   Vent_Analysis.py:316-334 */
#define VB_DO_VDP 4u         /* Vent_Analysis.calculate_VDP body   Vent_Analysis.py:244-257 */
#define VB_DO_CI 8u          /* CI.calculate_CI + percentile  CI.py:107-145, V_A.py:268-270 */
#define VB_DO_MASK_BORDER 16u/* Vent_Analysis.calculateBorder(mask) Vent_Analysis.py:225-231 */
#define VB_HOST_STATS_ONLY 32u /* vb_pipeline_batch_host: copy back only vb_subject_stats (cohort
                                screening: VDP / VDP_lb / CI / SNR per subject); output array
                                pointers of computed stages may then be NULL                   */

#define VB_MAX_LEVELS 8
#define VB_CI_OFF 0xFFFFu    /* ci_shell value of a voxel that is not a defect             */

typedef struct vb_ctx vb_ctx;

/* Parameters of sitk.N4BiasFieldCorrectionImageFilter (Vent_Analysis.py:330 uses defaults). */
typedef struct vb_n4_params {
    int32_t n_levels;                 /* len(MaximumNumberOfIterations)            4       */
    int32_t max_iter[VB_MAX_LEVELS];  /* MaximumNumberOfIterations                 50 x 4  */
    float conv_threshold;             /* ConvergenceThreshold                      0.001   */
    float fwhm;                       /* BiasFieldFullWidthAtHalfMaximum           0.15    */
    float wiener_noise;               /* WienerFilterNoise                         0.01    */
    int32_t n_bins;                   /* NumberOfHistogramBins                     200     */
    int32_t ncp[3];                   /* NumberOfControlPoints (row, col, slice)   4,4,4   */
    int32_t mask_label;               /* MaskLabel (UseMaskLabel = true)           1       */
    int32_t max_total_iters;          /* 0 = no cap; >0 stops after that many iterations
                                         overall (used by the stage-level parity tests)    */
} vb_n4_params;

/* Per-subject scalars.  Counts are exact integers; the Python shim forms the reference's
 * float64 expressions (100*sum(defect)/sum(mask), ...) from them. */
typedef struct vb_subject_stats {
    int32_t n4_iters[VB_MAX_LEVELS];  /* iterations run at each fitting level              */
    float n4_last_conv;               /* last convergence measurement                      */
    int32_t n4_n_mask;                /* #(mask == mask_label)                             */
    int32_t n_mask;                   /* #(mask != 0)            np.sum(mask) for 0/1 masks */
    float mean;                       /* float32 mean of N4 over the mask   V_A.py:246     */
    float p99;                        /* sorted(N4[mask>0])[int(M*.99)]     V_A.py:255     */
    int32_t n_defect;                 /* np.sum(defectArray)                V_A.py:251     */
    int32_t n_lb12;                   /* #(defectArrayLB in {1,2})          V_A.py:257     */
    int32_t n_noise;                  /* #noise voxels of calculate_SNR                    */
    double signal_mean, noise_mean, noise_std, snr; /* V_A.py:356 (float64 accumulation)  */
    int32_t ci_n;                     /* defect voxels given to CI                         */
    int32_t ci_t95;                   /* shell ordinal of sort(CV)[int(.95*D)], -1 if D==0 */
    int32_t ci_error;                 /* 1 = a voxel exhausted the table (VB_ERR_CI_RMAX)  */
    int32_t reserved;
} vb_subject_stats;

/* Output (and in/out) arrays of the pipeline; NULL = not wanted (where optional). */
typedef struct vb_buffers {
    const float* image;       /* in   HPvent.astype(float32)                               */
    const uint8_t* mask;      /* in   mask.astype(float32) -> uint8 (V_A.py:323,327)       */
    float* n4;                /* out  N4HPvent; INPUT when VB_DO_N4 is off and VDP is on    */
    float* logbias;           /* out  optional: log bias field (every voxel)               */
    uint8_t* defect;          /* out  defectArray as 0/1                                   */
    uint8_t* defect_lb;       /* out  defectArrayLB as 0..6                                */
    uint8_t* defect_border;   /* out  defectBorder as 0/1                                  */
    uint8_t* mask_border;     /* out  optional: mask_border as 0/1                         */
    uint16_t* ci_shell;       /* out  shell ordinal t per voxel, VB_CI_OFF off-defect;
                                      CIarray = radii[t] * min(vox) (vb_ci_table_radii)    */
    vb_subject_stats* stats;  /* out  [n]                                                  */
} vb_buffers;

int vb_version(void);
const char* vb_last_error(void);
int vb_device_count(void);
int vb_create(int device, vb_ctx** ctx);
void vb_destroy(vb_ctx* ctx);
/* kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t vb_launch_count(const vb_ctx* ctx);
void vb_n4_default_params(vb_n4_params* p);

/* Sphere-growing table (CI.getSpherePix, CI.py:33-63): `table` is the reference's N x 4
 * float64 [radius, dx, dy, dz] array (host).  Prepares, for one volume shape, the
 * de-duplicated flat offsets and the shell boundaries (CI.py:79-85, :93) on the device. */
int vb_ci_table_create(vb_ctx* ctx, const double* table, int64_t nrows, int32_t n0, int32_t n1,
                       int32_t n2, int32_t* handle);
/* radii[t] = table[ii_t - 1][0] for every shell boundary t (CI.py:105); returns the count */
int vb_ci_table_radii(vb_ctx* ctx, int32_t handle, double* radii, int32_t capacity, int32_t* n_bounds);

/* ---- stage entry points, device buffers ------------------------------------------------ */
/* Vent_Analysis.calculate_SNR (V_A.py:337-357): fills stats[i].snr & co. */
int vb_snr_batch_dev(vb_ctx* ctx, const float* image, const uint8_t* mask, int32_t n,
                     const int32_t* shapes, int32_t fov_buffer, vb_subject_stats* stats, void* stream);
/* Vent_Analysis.N4_bias_correction (V_A.py:316-334) */
int vb_n4_batch_dev(vb_ctx* ctx, const float* image, const uint8_t* mask, int32_t n,
                    const int32_t* shapes, const vb_n4_params* params, float* n4, float* logbias,
                    vb_subject_stats* stats, void* stream);
/* Vent_Analysis.calculate_VDP after N4 (V_A.py:244-257) */
int vb_vdp_batch_dev(vb_ctx* ctx, const float* n4, const uint8_t* mask, int32_t n,
                     const int32_t* shapes, float thresh, uint8_t* defect, uint8_t* defect_lb,
                     uint8_t* defect_border, vb_subject_stats* stats, void* stream);
/* Vent_Analysis.calculateBorder (V_A.py:225-231) on any 0/1 volume */
int vb_border_batch_dev(vb_ctx* ctx, const uint8_t* a, int32_t n, const int32_t* shapes,
                        uint8_t* border, void* stream);
/* CI.calculate_CI (CI.py:107-145) + the percentile of V_A.py:268-270 */
int vb_ci_batch_dev(vb_ctx* ctx, const uint8_t* defect, int32_t n, const int32_t* shapes,
                    const int32_t* table_handles, uint16_t* ci_shell, vb_subject_stats* stats,
                    void* stream);

/* Vent_Analysis.build4DdataArray (V_A.py:292-313), the step after the path: per
   subject float32 [rows][cols][slices][6] = proton[0] HPvent[1] mask[2] N4HPvent[3] defectArray[4]
   CIarray[5] (mm = float32(radius[shell] * vox_min[i]); 0 off the defect).  A null input leaves its
   channel 0, like a missing attribute in the reference.  `out` holds 6 floats per voxel. */
int vb_pack6_batch_dev(vb_ctx* ctx, int32_t n, const int32_t* shapes, const int32_t* table_handles,
                       const double* vox_min, const float* proton, const float* image, const uint8_t* mask,
                       const float* n4, const uint8_t* defect, const uint16_t* ci_shell, float* out,
                       void* stream);
int vb_pack6_batch_host(vb_ctx* ctx, int32_t n, const int32_t* shapes, const int32_t* table_handles,
                        const double* vox_min, const float* proton, const float* image, const uint8_t* mask,
                        const float* n4, const uint8_t* defect, const uint16_t* ci_shell, float* out);

/* ---- whole path --------------------------------------------------------------------------- */
int vb_pipeline_batch_dev(vb_ctx* ctx, int32_t n, const int32_t* shapes, const int32_t* table_handles,
                          const vb_n4_params* params, float thresh, uint32_t flags,
                          const vb_buffers* dev, void* stream);
/* Same with HOST buffers: H2D of image+mask, the kernels, D2H of every non-NULL output,
 * then a stream synchronise.  Returns VB_ERR_CI_RMAX if any subject hit it. */
int vb_pipeline_batch_host(vb_ctx* ctx, int32_t n, const int32_t* shapes, const int32_t* table_handles,
                           const vb_n4_params* params, float thresh, uint32_t flags,
                           const vb_buffers* host);

/* ---- whole path, COMPACT host I/O ------------------------------------------------------------
 * The same path with the fewest bytes over PCIe: the image as the reference reads it from DICOM
 * (uint16 pixel data, Vent_Analysis.py:169-181) or float32, the 0/1 mask bit-packed, and on the way
 * back one byte of codes per voxel plus the CI map as a sparse list (2.125 B/voxel up and about
 * 5.2 B/voxel down instead of 5 and 9).  Arrays are whole-batch, subjects concatenated.
 *   mask_bits   subject i owns ceil(V_i / 32) 32-bit words; bit b of word w = voxel 32 w + b of the
 *               C-ordered volume (numpy: packbits(mask.ravel() != 0, bitorder='little'), padded to a
 *               multiple of 4 bytes).  The mask is 0/1 here; multi-label masks use mask_u8.
 *   codes       bits 0-2 defectArrayLB (0..6) | bit 3 defectArray | bit 4 defectBorder |
 *               bit 5 mask_border (with VB_DO_MASK_BORDER)
 *   ci_index / ci_value   for subject i the entries [ci_offset[i], ci_offset[i+1]): C-order flat
 *               voxel index inside the subject and shell ordinal t (CIarray = radii[t]*min(vox)) of
 *               every defect voxel, in no particular order.  VB_ERR_ARG if ci_capacity is too small
 *               (a batch never needs more than its voxel count). */
#define VB_CODE_LB_MASK 7u
#define VB_CODE_DEFECT 8u
#define VB_CODE_DEFECT_BORDER 16u
#define VB_CODE_MASK_BORDER 32u
typedef struct vb_host_io {
    const uint16_t* image_u16; /* in   exactly one of image_u16 / image_f32                      */
    const float* image_f32;    /* in                                                            */
    const uint32_t* mask_bits; /* in   exactly one of mask_bits / mask_u8                        */
    const uint8_t* mask_u8;    /* in                                                            */
    float* n4;                 /* out  optional N4HPvent                                         */
    float* logbias;            /* out  optional                                                  */
    uint8_t* codes;            /* out  optional (needs VB_DO_VDP)                                */
    uint32_t* ci_index;        /* out  with VB_DO_CI (optional: NULL = CI scalars only)          */
    uint16_t* ci_value;        /* out                                                            */
    int64_t ci_capacity;       /* entries available in ci_index / ci_value                       */
    int64_t* ci_offset;        /* out  [n + 1]                                                   */
    vb_subject_stats* stats;   /* out  [n]                                                       */
} vb_host_io;
int vb_pipeline_batch_host_packed(vb_ctx* ctx, int32_t n, const int32_t* shapes, const int32_t* table_handles,
                                  const vb_n4_params* params, float thresh, uint32_t flags,
                                  const vb_host_io* io);

/* ---- overlay compositions: the consumers right behind the path -------------------
 * vb_overlay_rgb_*: uint8 [rows][cols][slices][3] per subject, a 0/1 array `b` painted on the
 * min-max normalised |a| (normalisation over the whole volume).
 *   gui = 0  Vent_Analysis.exportDICOM (V_A.py:387-391): BW = uint8(normalize(|N4HPvent|) * 255) in
 *            float32, R = 255 on defects else BW, G = B = BW off the defects;
 *   gui = 1  the GUI's colorBinary + arrayToImage (V_A.py:621,628-634): the same in float64,
 *            R = uint8((A*(B==0) + B) * 255).
 * mode: VB_NORM_F32 when the reference would normalise a float32 array, VB_NORM_F64 for integer /
 * float64 arrays (their values are passed as float32 and must be exact in it, e.g. uint16 DICOM data). */
#define VB_NORM_F32 0
#define VB_NORM_F64 1
#define VB_ERR_EMPTY 5       /* cropToData: an axis has no index > 0 with signal: the reference raises
                                IndexError (V_A.py:443-451)                                            */
#define VB_ERR_CI_COLOR 6    /* screenShot: a CI value reached 40 mm, parula[int(CI*64/40)] is out of
                                range: the reference raises IndexError (V_A.py:482-484)                */
int vb_overlay_rgb_batch_dev(vb_ctx* ctx, int32_t n, const int32_t* shapes, const float* a, const uint8_t* b,
                             int32_t gui, int32_t mode, uint8_t* rgb, void* stream);
int vb_overlay_rgb_batch_host(vb_ctx* ctx, int32_t n, const int32_t* shapes, const float* a, const uint8_t* b,
                              int32_t gui, int32_t mode, uint8_t* rgb);
/* Vent_Analysis.cropToData (V_A.py:430-456) on the mask: boxes[6 i ..] = r0, r1, c0, c1, s0, s1
 * (half-open).  VB_ERR_EMPTY if any subject has no box (its entries are -1). */
int vb_crop_boxes_batch_host(vb_ctx* ctx, int32_t n, const int32_t* shapes, const uint8_t* mask, int32_t border,
                             int32_t border_slices, int32_t* boxes);
/* Vent_Analysis.screenShot (V_A.py:466-494) up to the array handed to PIL: per subject uint8
 * [7 R][S C][3] (R, C, S = the box of cropToData(mask, border=5)); bands: blank, blank, proton, HPvent,
 * N4 + mask border, N4 + defects, N4 + CI in parula colours.  Subject i's image starts at byte
 * out_offset[i] of `out` (out_offset[n] = bytes used; VB_ERR_ARG if out_capacity is too small;
 * 21 bytes per voxel always suffice). */
typedef struct vb_montage_inputs {
    const float* proton;        /* optional (blank band)                                           */
    const float* image;         /* HPvent                                                          */
    const float* n4;            /* N4HPvent                                                        */
    const uint8_t* mask;        /* for the crop box                                                */
    const uint8_t* mask_border;
    const uint8_t* defect;
    const uint16_t* ci_shell;   /* optional: with table_handles and vox_min                        */
    int32_t image_mode, proton_mode;   /* VB_NORM_F32 / VB_NORM_F64                                */
} vb_montage_inputs;
int vb_montage_batch_host(vb_ctx* ctx, int32_t n, const int32_t* shapes, const int32_t* table_handles,
                          const double* vox_min, const vb_montage_inputs* host, const double* parula64x3,
                          int32_t* boxes, int64_t* out_offset, uint8_t* out, int64_t out_capacity);

/* ---- k-space reconstruction: the step before the path for TWIX users ------------
 * Vent_Analysis.process_RAW (V_A.py:536-540): per slice fftshift(fft2(fftshift(K[:, :, k]))), then
 * transpose(., (1, 0, 2))[:, ::-1, :].  `kspace`: complex64 [rows][cols][slices] per subject as
 * interleaved (re, im) float32; `image`: complex64 [cols][rows][slices] likewise (the reference
 * keeps complex128; single precision holds it to about 1e-6 of the largest magnitude). */
int vb_recon_batch_dev(vb_ctx* ctx, int32_t n, const int32_t* shapes, const float* kspace, float* image, void* stream);
int vb_recon_batch_host(vb_ctx* ctx, int32_t n, const int32_t* shapes, const float* kspace, float* image);

/* Page-locked host memory for the *_host entry points (asynchronous copies need it to overlap with
 * compute).  The pages are touched by the calling thread, so a caller that has bound itself to the
 * CPUs next to the context's GPU gets NUMA-local buffers. */
int vb_host_alloc(vb_ctx* ctx, size_t bytes, void** ptr);
int vb_host_free(vb_ctx* ctx, void* ptr);

/* Per-stage device timing of vb_pipeline_batch_dev: with profiling on, every call records a
 * CUDA event on the launching stream before the first stage and after each of the VB_N_STAGES
 * stages (SNR, N4, VDP, mask border, CI; a skipped stage reads ~0).  vb_profile_fetch waits for
 * the recorded calls, returns the SUM of each stage's elapsed milliseconds over them and their
 * count, and resets the record.  (bench.py's roofline numbers come from here.) */
#define VB_N_STAGES 5
int vb_set_profiling(vb_ctx* ctx, int32_t on);
int vb_profile_fetch(vb_ctx* ctx, double* stage_ms, int32_t* n_calls);

/* N4 on small batches (the reference's use: ONE subject per Vent_Analysis object,
 * Vent_Analysis.py:316-334).  A batch that would leave most SMs idle runs with a thread-block
 * cluster of K CTAs (K SMs) per subject; results are bit-identical to one CTA per subject.
 * vb_set_n4_cluster(max_k):
 *   0  automatic (default): the largest K in {8, 4, 2, 1} with n * K <= 7/8 of the SM count, for
 *      volumes of >= 200 000 voxels that are not a chunk of a multi-stream schedule (the measured
 *      break-even points: profiles/r02_n4_cluster_timings.jsonl);
 *   1  always one CTA per subject;
 *   2 / 4 / 8  use up to that many CTAs per subject whenever n * K <= SM count (any volume size).
 * vb_n4_cluster_size: K of the most recent N4 launch of this context (1 before any launch). */
int vb_set_n4_cluster(vb_ctx* ctx, int32_t max_k);
int32_t vb_n4_cluster_size(const vb_ctx* ctx);

/* debug: last Parzen histogram / E(u|v) table / lattice of subject `i` of the most recent
 * vb_n4_batch_* call (needs ctx debug capture; see vb_set_debug). */
int vb_set_debug(vb_ctx* ctx, int32_t on);
int vb_n4_debug_fetch(vb_ctx* ctx, int32_t subject, float* hist, float* emap, float* lattice,
                      int32_t lattice_capacity, int32_t* ncp, float* conv_trace, int32_t trace_capacity);

#ifdef __cplusplus
}
#endif
#endif
