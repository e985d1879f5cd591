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
    void *stream;    /* cudaStream_t to run on (AMX_SPACE_DEVICE); NULL = the plan's own stream */
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

#ifdef __cplusplus
}
#endif
#endif /* AMICO_B200_H */
