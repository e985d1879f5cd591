/*
 * amico_b200 -- C ABI of the B200-native per-voxel microstructure fit.
 *
 * This is the drop-in boundary for the ONE hot path of daducci/AMICO: what `model.fit(evaluation)`
 * does for NODDI / FreeWater / CylinderZeppelinBall / SANDI (amico/models.pyx:795-991, 1147-1286,
 * 526-652, 1489-1627), i.e. for every mask voxel: direction -> LUT index (amico/lut.pyx:314-356),
 * dictionary look-up, NNLS / non-negative elastic-net (the `nnls` / `lasso` entry points the
 * reference cimports from spams-cython, amico/models.pyx:18; call sites :615, :911, :926, :940,
 * :1238, :1569), scalar maps and optional fit errors.
 *
 * Plain C: pointers, sizes, ints.  No torch / numpy types.  All functions return 0 on success and a
 * negative AMX_E_* code on failure; amx_last_error() gives the message of the calling thread's
 * last failure.  Nothing allocated by the library crosses the ABI except the opaque plan.
 *
 * Ownership: the caller owns every buffer it passes; the library owns the plan's device tables
 * (uploaded KERNELS, per-direction Gram tables) and its internal workspace until amx_plan_destroy.
 * Threading: a plan may be used by one host thread at a time; different plans are independent.
 * There is NO CPU fallback: every entry point below fails with AMX_E_CUDA when no usable GPU exists.
 */
#ifndef AMICO_B200_H
#define AMICO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMX_VERSION 100

/* error codes */
#define AMX_OK 0
#define AMX_E_INVALID (-1)   /* bad argument */
#define AMX_E_CUDA (-2)      /* CUDA runtime / driver failure, or no device */
#define AMX_E_LUT_RANGE (-3) /* a direction fell outside the LUT angle grid: the reference raises
                                RuntimeError('"amico.lut.dir_to_lut_idx" index out of bounds ...'),
                                amico/lut.pyx:352-354 */
#define AMX_E_CAPACITY (-4)  /* an active set outgrew the solver workspace */

/* models */
#define AMX_MODEL_NODDI 0
#define AMX_MODEL_FREEWATER 1
#define AMX_MODEL_CZB 2 /* CylinderZeppelinBall */
#define AMX_MODEL_SANDI 3

/* fit flags (BaseModel.fit reads them from the evaluation config, amico/models.pyx:213-217, :797, :1149) */
#define AMX_FLAG_RMSE 1u  /* doComputeRMSE   -> rmse  */
#define AMX_FLAG_NRMSE 2u /* doComputeNRMSE  -> nrmse */
#define AMX_FLAG_EXTRA 4u /* NODDI: doSaveModulatedMaps -> extra (n_vox x 2);
                             FreeWater: doSaveCorrectedDWI -> extra (n_vox x m) */
#define AMX_FLAG_EXACT 8u /* bit-reproducible mode: follow the reference's CPU arithmetic operation for operation (un-fused, same
                             summation order).  FreeWater / CylinderZeppelinBall / SANDI: the TMA-staged kernel with the SPAMS path
                             restated step by step instead of the DMMA-batched throughput kernel (identical maps to ~1e-12);
                             NODDI: every voxel through the A-space Lawson-Hanson path (amico/models.pyx:911, 940), not only
                             the exact-fit ones.  Results equal the CPU oracle bit for bit; throughput is not the point. */

/* element type of y */
#define AMX_F32 0
#define AMX_F64 1

/* address space of the per-voxel buffers handed to amx_fit */
#define AMX_SPACE_HOST 0   /* host pointers: the library stages H2D / D2H itself */
#define AMX_SPACE_DEVICE 1 /* device pointers on the plan's GPU (e.g. torch tensors' data_ptr) */

typedef struct amx_plan amx_plan;

const char *amx_last_error(void);
int amx_version(void);
/* Number of CUDA devices visible (>= 0), or AMX_E_CUDA. */
int amx_device_count(void);

/*
 * Plans: upload one model's KERNELS (the dict `<Model>.resample` returns) to `device`, re-lay the
 * rotated LUT as one contiguous (m x n) fp32 slab per direction and precompute the per-direction
 * Gram tables.  All array arguments are HOST pointers in the reference's own layouts.
 * `htable` is the int16[181*181] table of amico/lut.pyx:71-91.
 */

/* NODDI (KERNELS: amico/models.pyx:762-768).  wm: float32 [n_wm][ndirs][m] C order; iso: float32 [m];
 * norms: float64 [dwi_count][n_wm]; icvf, kappa: float32 [n_wm]; dwi_idx: int64 [dwi_count]
 * (scheme.dwi_idx; when m == 1 + dwi_count the rows 1..m-1 are used instead, :916-918). */
int amx_plan_create_noddi(int device, int m, int ndirs, int n_wm, const float *wm, const float *iso,
                          const double *norms, const float *icvf, const float *kappa,
                          const int64_t *dwi_idx, int dwi_count, int is_exvivo, const int16_t *htable,
                          amx_plan **out);

/* FreeWater (KERNELS: amico/models.pyx:1122-1123).  D: float32 [n_perp][ndirs][m]; CSF: float32 [n_iso][m]. */
int amx_plan_create_freewater(int device, int m, int ndirs, int n_perp, const float *D, int n_iso,
                              const float *CSF, int is_mouse, const int16_t *htable, amx_plan **out);

/* CylinderZeppelinBall (KERNELS: amico/models.pyx:491-493).  wmr: float32 [n_rs][ndirs][m];
 * wmh: float32 [n_perp][ndirs][m]; iso: float32 [n_iso][m]; Rs: float64 [n_rs] (metres). */
int amx_plan_create_czb(int device, int m, int ndirs, int n_rs, const float *wmr, int n_perp,
                        const float *wmh, int n_iso, const float *iso, const double *Rs,
                        const int16_t *htable, amx_plan **out);

/* SANDI (KERNELS: amico/models.pyx:1456-1457).  signal: float64 [m][n] column-major (Fortran order,
 * columns L2-normalised), n = n_rs + n_in + n_iso; norms: float64 [n]; Rs, d_in, d_isos: the
 * model's physical grids (float64). */
int amx_plan_create_sandi(int device, int m, int n_rs, int n_in, int n_iso, const double *signal,
                          const double *norms, const double *Rs, const double *d_in,
                          const double *d_isos, amx_plan **out);

int amx_plan_destroy(amx_plan *plan);

/* Shape information of a plan: model id, m, n atoms, number of maps in `estimates`, ndirs, device. */
int amx_plan_info(const amx_plan *plan, int *model, int *m, int *n_atoms, int *n_maps, int *ndirs,
                  int *device);

typedef struct amx_fit_args {
    int space;       /* AMX_SPACE_HOST | AMX_SPACE_DEVICE: where every pointer below lives */
    int y_dtype;     /* AMX_F32 | AMX_F64 */
    const void *y;   /* [n_vox][m], C order, >= 0 (evaluation.y, amico/core.py:451-452) */
    int64_t n_vox;
    double *dirs;    /* [n_vox][3] principal directions (evaluation.DIRs); FLIPPED IN PLACE to the
                        y >= 0 hemisphere exactly like the reference (amico/lut.pyx:335-338);
                        NULL for SANDI */
    double lambda1;  /* solver_params['lambda1'] */
    double lambda2;  /* solver_params['lambda2'] */
    uint32_t flags;  /* AMX_FLAG_* */
    double *estimates; /* out [n_vox][n_maps] */
    double *rmse;      /* out [n_vox] when AMX_FLAG_RMSE, else may be NULL */
    double *nrmse;     /* out [n_vox] when AMX_FLAG_NRMSE */
    double *extra;     /* out when AMX_FLAG_EXTRA (see flag) */
    int32_t *lut_out;     /* optional out [n_vox]: LUT index of each voxel (diagnostics / tests) */
    int32_t *support_out; /* optional out [n_vox]: NODDI stage-2 support size incl. iso(/dot);
                             other models: number of non-zero coefficients */
    double *coeff_out;    /* optional out [n_vox][n_atoms]: the fitted coefficients x (diagnostics / tests) */
    void *stream;    /* cudaStream_t to run on (AMX_SPACE_DEVICE); NULL = the legacy default stream (stream 0) */
} amx_fit_args;

/* Fit every voxel.  Synchronous for AMX_SPACE_HOST; for AMX_SPACE_DEVICE the work is enqueued on
 * `stream` and the call returns after the final status word has been read back (one 16-byte D2H).
 * On AMX_E_LUT_RANGE `*err_voxel` (may be NULL) receives the first offending voxel index. */
int amx_fit(amx_plan *plan, const amx_fit_args *args, int64_t *err_voxel);

/* LUT index of n directions (dirs flipped in place), bit-for-bit amico/lut.pyx:314-356; out-of-range
 * directions give -1 and the call returns AMX_E_LUT_RANGE after filling `idx`. */
int amx_lut_indices(amx_plan *plan, int space, double *dirs, int64_t n, int32_t *idx);

/* Device-side durations (milliseconds, CUDA events on the launching stream) of the last amx_fit on
 * this plan: out[0] = LUT index + voxel binning, out[1] = the fused per-voxel fit kernel,
 * out[2] = whole enqueue-to-done span including any staging copies.  n <= 8 values are written. */
int amx_plan_last_timing(amx_plan *plan, double *out_ms, int n);

/* Counters of the last amx_fit: out[0] = kernels launched, out[1] = voxel tiles (upper bound),
 * out[2] = voxels a kernel without slow path could not finish (-> AMX_E_CAPACITY; 0 in a healthy
 * run), out[3] = bytes of dynamic shared memory per CTA, out[4] = warps per CTA, out[5] = 1 if the
 * slab was staged through TMA, out[6] = voxel fits redone by the scalar slow path (active sets
 * larger than a warp), out[7] = grid size. */
int amx_plan_last_counters(amx_plan *plan, int64_t *out, int n);

/* =====================================================================================================
 * The callers either side of model.fit(): what
 * amico.Evaluation.load_data() / fit() do around the per-voxel solve, on the GPU, so that a raw 4-D
 * volume can stay in HBM from load to maps.  These entry points do not need a plan.
 * ===================================================================================================== */

#define AMX_E_NONFINITE (-5) /* NaN / Inf in the signal and no replacement value given: the reference stops with
                                ERROR('Nan or Inf values in the raw signal ...'), amico/core.py:151-156, 273-278 */

#define AMX_PRE_NORMALIZE 1u   /* doNormalizeSignal: divide by the voxel's mean b0 (amico/core.py:209-222) */
#define AMX_PRE_MERGE_B0 2u    /* doMergeB0: [mean of the b0 volumes | dwi volumes] (amico/core.py:224-227) */
#define AMX_PRE_DIR_AVG 4u     /* doDirectionalAverage: [mean b0 | mean of each shell, ascending b] (amico/core.py:231-266) */
#define AMX_PRE_REPLACE_BAD 8u /* replace_bad_voxels given: NaN / +-Inf -> replace_bad (np.nan_to_num, core.py:153, :275) */

typedef struct amx_pre_args {
    int space;          /* AMX_SPACE_HOST | AMX_SPACE_DEVICE for dwi, mask, y, vox_idx, mean_b0s */
    int device;
    const float *dwi;   /* [n_total][nS] niiDWI_img (float32, amico/core.py:136), voxel-major C order */
    int64_t n_total;    /* voxels in the volume */
    int nS;             /* volumes = scheme.nS */
    const uint8_t *mask; /* [n_total] niiMASK_img, or NULL (= all ones); voxels with mask == 1 are kept, in C-order
                            scan order, exactly like `niiDWI_img[niiMASK_img==1, :]` (amico/core.py:451) */
    /* scheme index lists: HOST pointers in either space (they are tiny) */
    const int32_t *b0_idx;  int b0_count;   /* scheme.b0_idx */
    const int32_t *dwi_idx; int dwi_count;  /* scheme.dwi_idx */
    const int32_t *shell_idx;               /* AMX_PRE_DIR_AVG: the shells' 'idx' lists concatenated in ascending-b order */
    const int32_t *shell_off;               /* [n_shells + 1] offsets into shell_idx */
    int n_shells;
    uint32_t flags;       /* AMX_PRE_* */
    float b0_threshold;   /* voxels whose mean b0 <= this get norm factor 0; the reference's value is
                             b0_min_signal * mean(mean_b0s[mean_b0s > 0]) (amico/core.py:216), 0 by default */
    float replace_bad;    /* AMX_PRE_REPLACE_BAD */
    /* outputs */
    float *y;             /* [n_kept][m_out] pre-processed signal of the kept voxels, negative values set to 0
                             (amico/core.py:452); float32 holds it exactly: the reference's volume is float32 */
    int64_t y_capacity;   /* rows y can hold (>= number of mask==1 voxels) */
    int32_t *vox_idx;     /* [n_kept] flat index of each kept voxel (ascending) */
    float *mean_b0s;      /* optional [n_total]: mean of the raw b0 volumes (evaluation.mean_b0s, core.py:212) */
    void *stream;         /* AMX_SPACE_DEVICE: cudaStream_t to enqueue on (NULL = legacy default stream) */
} amx_pre_args;

/* Pre-process a raw volume and compact the mask voxels.  m_out = nS, 1 + dwi_count (MERGE_B0) or 1 + n_shells
 * (DIR_AVG).  Arithmetic is the reference's float32 arithmetic operation for operation (numpy sums the indexed
 * volumes sequentially in index order, then divides by the count), so y is bit-identical to
 * `niiDWI_img[niiMASK_img==1, :]` after load_data().  Returns after the kept-voxel count has been read back. */
int amx_preprocess(const amx_pre_args *args, int64_t *n_kept, int *m_out);

/* Mean of the raw b0 volumes for every voxel (first half of doNormalizeSignal): lets the caller form
 * b0_threshold when b0_min_signal != 0.  mean_b0s: [n_total] float32 in `space`. */
int amx_mean_b0(int space, int device, const float *dwi, int64_t n_total, int nS, const int32_t *b0_idx, int b0_count,
                float *mean_b0s, void *stream);

/* Principal diffusion direction of every voxel = what `dipy.reconst.dti.TensorModel(gtab, fit_method='OLS')
 * .fit(y).directions` yields at amico/core.py:436, 458: D = W log(max(y, min_signal)), eigen-decomposition of the
 * symmetric 3x3 tensor, eigenvector of the largest eigenvalue (unit length; its SIGN is arbitrary, as it is with
 * LAPACK, and irrelevant downstream: amico/lut.pyx:335-338 flips to the y >= 0 hemisphere).
 * W: HOST float64 [6][m], the rows Dxx, Dxy, Dyy, Dxz, Dyz, Dzz of pinv(design matrix); dirs: out [n_vox][3] float64. */
int amx_dti_directions(int device, int space, const void *y, int y_dtype, int64_t n_vox, int m, const double *W,
                       double min_signal, double *dirs, void *stream);

/* The same with `fit_method='WLS'` (amico/core.py:95, 419, 436 -> dipy wls_fit_tensor): weights w = exp(X beta_ols), then the
 * weighted least-squares tensor min || diag(w) (X beta - log max(y, min_signal)) ||.
 * W7: HOST float64 [7][m] = pinv(design matrix) (all seven rows); X: HOST float64 [m][7] = the design matrix itself. */
int amx_dti_directions_wls(int device, int space, const void *y, int y_dtype, int64_t n_vox, int m, const double *W7,
                           const double *X, double min_signal, double *dirs, void *stream);

/* RESULTS['MAPs'][mask==1, :] = estimates (amico/core.py:472-498): zero-fill volume [n_total][k] (float32) and
 * scatter the float64 rows values[i][0..k) to voxel vox_idx[i]. */
int amx_scatter_maps(int device, int space, const double *values, int64_t n_vox, int k, const int32_t *vox_idx,
                     float *volume, int64_t n_total, void *stream);

/* Row f-3: `amico.lut.resample_kernel` (amico/lut.pyx:274-311) for a stack of atoms: project the rotated response
 * functions from SH space to the subject's acquisition scheme.  KRlm: [n_rows][n_coef] float32, n_rows = atoms x LUT
 * directions of the `A_###.npy` files `generate` wrote (n_rows = atoms for isotropic atoms); Ylm_out: [dwi_count][n_coef]
 * float32 and idx_out: int32 [dwi_count] from `aux_structures_resample` (lut.pyx:199-224); merge_idx: int32 [nS_out], the
 * scheme row each output column takes (`[b0_idx[0], dwi_idx...]` with doMergeB0, else 0..nS-1; amico/models.pyx:756-761).
 * out: [n_rows][nS_out] float32 = KR[:, merge_idx]; columns that are not dwi rows are 1 (lut.pyx:297).  idx_out and
 * merge_idx are HOST pointers in either space.  Accumulates in fp64 and rounds once. */
int amx_resample_kernels(int device, int space, const float *KRlm, int64_t n_rows, int n_coef, const float *Ylm_out,
                         const int32_t *idx_out, int dwi_count, const int32_t *merge_idx, int nS_out, int nS, float *out,
                         void *stream);

/* On-disk image -> the layout the fit wants.  src: the NIfTI data block as stored, memory order [nS][n_total] (x fastest,
 * volume index slowest), element type `nifti_datatype` (NIfTI-1 codes: 2 uint8, 4 int16, 8 int32, 16 float32, 64 float64,
 * 256 int8, 512 uint16, 768 uint32); dst: float32 [n_total][nS] (voxels in the file's own x-fastest order).  Applies
 * `raw * scl_slope + scl_inter` in float64 when the header's scaling is meaningful, then rounds to float32 -- what
 * `nibabel.load(...).get_fdata().astype(np.float32)` yields at amico/core.py:135-136. */
int amx_volume_to_voxel_major(int device, int space, const void *src, int nifti_datatype, int64_t n_total, int nS,
                              double scl_slope, double scl_inter, float *dst, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AMICO_B200_H */
