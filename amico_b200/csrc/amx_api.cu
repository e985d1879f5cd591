// C ABI of the B200-native per-voxel fit (see include/amico_b200.h).
#include "../../include/amico_b200.h"
#include "amx_kernels.cuh"
#include "amx_lean.cuh"
#include "amx_slow.cuh"
#include "amx_exact.cuh"
#include "amx_small.cuh"
#include "amx_err.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace amx;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(AMX_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                              \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

int amx::set_error(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

namespace {

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// ---- host-side staging of the signal --------------------------------------------------------------------------------------
// The reference hands a model `evaluation.y` as a PAGEABLE float64 array (amico/core.py:451-452) whose values are float32 (the
// volume is float32, core.py:136).  A cudaMemcpyAsync from such memory is staged by the driver on one thread at a few GB/s.
// Instead, host threads move each voxel chunk into pinned memory -- narrowing float64 to float32 on the way when every value
// survives the round trip, which halves the PCIe bytes -- while the GPU is busy with the previous chunks.
int host_threads()
{
    const int hw = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(env_int("AMX_HOST_THREADS", std::min(hw > 0 ? hw : 1, 16)), 64));
}

template <typename F>
void parallel_blocks(size_t n_blocks, int threads, F f)
{
    threads = (int)std::min<size_t>((size_t)threads, n_blocks);
    if (threads <= 1) { for (size_t b = 0; b < n_blocks; ++b) f(b); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    auto work = [&]() { for (size_t b = next.fetch_add(1); b < n_blocks; b = next.fetch_add(1)) f(b); };
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

// Copy `count` signal values into pinned `dst`.  float64 input: written as float32 when lossless (returns AMX_F32), else as is.
int stage_signal(const void *src, int src_dtype, void *dst, size_t count, int threads)
{
    constexpr size_t BLK = 1 << 16;
    const size_t nb = (count + BLK - 1) / BLK;
    if (src_dtype == AMX_F64) {
        std::atomic<int> lossy{0};
        const double *s = (const double *)src;
        float *d = (float *)dst;
        parallel_blocks(nb, threads, [&](size_t b) {
            const size_t i0 = b * BLK, i1 = std::min(count, i0 + BLK);
            int bad = 0;
            for (size_t i = i0; i < i1; ++i) {
                const float f = (float)s[i];
                bad |= ((double)f != s[i]);
                d[i] = f;
            }
            if (bad) lossy.store(1, std::memory_order_relaxed);
        });
        if (!lossy.load()) return AMX_F32;
        parallel_blocks(nb, threads, [&](size_t b) {
            const size_t i0 = b * BLK, i1 = std::min(count, i0 + BLK);
            memcpy((double *)dst + i0, s + i0, (i1 - i0) * sizeof(double));
        });
        return AMX_F64;
    }
    parallel_blocks(nb, threads, [&](size_t b) {
        const size_t i0 = b * BLK, i1 = std::min(count, i0 + BLK);
        memcpy((float *)dst + i0, (const float *)src + i0, (i1 - i0) * sizeof(float));
    });
    return AMX_F32;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

}  // namespace

struct amx_plan {
    int device = 0, model = 0, m = 0, n = 0, n_pad = 0, ndirs = 1, n_maps = 0, npl = 1;
    int n_rot = 0, n_wm = 0, exvivo = 0, mouse = 0, n_perp = 0, n_iso = 0, n_rs = 0, n_in = 0, dc = 0, norms_const = 0;
    bool slab_f64 = false;
    size_t slab_stride = 0;   // elements per direction
    unsigned slab_bytes = 0;  // bytes to stage per direction
    void *d_slab = nullptr;
    double *d_T1 = nullptr, *d_T2 = nullptr, *d_diag0 = nullptr, *d_W = nullptr;
    int ldW = 0; size_t W_stride = 0;
    double ridge_baked = -1.0;  // ridge currently added to the diagonal of d_T2 (< 0: none)
    int ldT1 = 0, ldT2 = 0, K2 = 0;
    size_t T1_stride = 0, T2_stride = 0;
    int16_t *d_htable = nullptr;
    int *d_dwi_rows = nullptr;
    double *d_norms = nullptr;
    float *d_icvf = nullptr, *d_kappa = nullptr;
    double *d_Rs = nullptr, *d_sandi_norms = nullptr, *d_d_in = nullptr, *d_d_isos = nullptr;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_lut[2] = {nullptr, nullptr};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [0] pre-LUT [1] post-binning [2] post-fit [3] start [4] end
    // workspace
    // per-launch workspace; two sets so that consecutive voxel chunks can be in flight on two compute streams
    struct Work { DevBuf lut, order, bins, tiles, status, scratch, xiso, supmask, ovf_list, slow_ws, exact_list, exact_a, c1_all, redo; } work[2];
    cudaStream_t cs[2] = {nullptr, nullptr};          // [0] == stream
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    struct Stage { DevBuf y, dirs, est, rmse, nrmse, extra, sup, coef, lut; } stg[2];  // host-path staging, double buffered
    void *hpin[2] = {nullptr, nullptr};  // pinned host staging of pageable / float64 signals (host path)
    size_t hpin_cap[2] = {0, 0};
    void *hout[4] = {nullptr, nullptr, nullptr, nullptr};  // pinned landing buffers of pageable outputs: estimates, dirs, rmse, nrmse
    size_t hout_cap[4] = {0, 0, 0, 0};
    int max_smem = 0, sm_count = 0;
    // last-call records
    double last_ms[8] = {0};
    int64_t last_cnt[16] = {0};
    bool timing_valid = false;
};

namespace {

template <typename T>
int upload(T **dst, const T *src, size_t count)
{
    CK(cudaMalloc((void **)dst, std::max<size_t>(count, 1) * sizeof(T)));
    if (count) CK(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return AMX_OK;
}

int plan_common_init(amx_plan *pl, int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return fail(AMX_E_CUDA, "no CUDA device available (%s): amico_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(AMX_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    CK(cudaSetDevice(device));
    pl->device = device;
    CK(cudaDeviceGetAttribute(&pl->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CK(cudaDeviceGetAttribute(&pl->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    pl->cs[0] = pl->stream;
    CK(cudaStreamCreateWithFlags(&pl->cs[1], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
    for (auto &e : pl->ev_join) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking));
    for (auto &ev : pl->ev) CK(cudaEventCreate(&ev));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&pl->ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pl->ev_comp[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pl->ev_out[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pl->ev_lut[i], cudaEventDisableTiming));
    }
    return AMX_OK;
}

// slab geometry shared by every model
void slab_geometry(amx_plan *pl, size_t elem)
{
    pl->n_pad = pl->n | 1;  // odd row stride: conflict-free for both lane-per-atom and lane-per-row access
    // NODDI feeds DMMA B-fragments from the slab (lane -> row lane%4, column lane/4): stride = 8 (mod 16) words
    // makes those 32 addresses hit 32 distinct banks
    if (pl->model == AMX_MODEL_NODDI) pl->n_pad = ((pl->n + 7) & ~15) + 8 >= pl->n ? ((pl->n + 7) & ~15) + 8 : ((pl->n + 23) & ~15) + 8;
    size_t bytes = (size_t)pl->m * pl->n_pad * elem;
    bytes = (bytes + 127) & ~(size_t)127;
    pl->slab_bytes = (unsigned)bytes;
    pl->slab_stride = bytes / elem;
    pl->npl = (pl->n + 31) / 32;
}

int build_gram(amx_plan *pl, double **G, int K, int *ld, size_t *stride, const int *d_rows, int nrows, const double *d_norms,
               int ldn, int norms_const)
{
    *ld = (K + 3) & ~3;
    *stride = (size_t)K * *ld;
    // + 256 doubles: the warp solvers read whole 32-lane column groups of a row without bounds predicates, i.e. up to
    // 32 * NPL - K entries past a row's end (the next row; past the very last row: this zero padding)
    CK(cudaMalloc((void **)G, ((size_t)pl->ndirs * *stride + 256) * sizeof(double)));
    CK(cudaMemsetAsync(*G, 0, ((size_t)pl->ndirs * *stride + 256) * sizeof(double), pl->stream));
    int by = std::max(1, std::min(64, (K * K + 8 * 256 - 1) / (8 * 256)));
    dim3 grid(pl->ndirs, by);
    if (pl->slab_f64)
        k_gram<double><<<grid, 256, 0, pl->stream>>>((const double *)pl->d_slab, pl->slab_stride, pl->n_pad, K, d_rows, nrows,
                                                      d_norms, ldn, norms_const, *G, *ld, *stride);
    else
        k_gram<float><<<grid, 256, 0, pl->stream>>>((const float *)pl->d_slab, pl->slab_stride, pl->n_pad, K, d_rows, nrows,
                                                     d_norms, ldn, norms_const, *G, *ld, *stride);
    CK(cudaGetLastError());
    return AMX_OK;
}

int build_rotated_plan(amx_plan *pl, const float *rot0, int n0, const float *rot1, int n1, int with_dot, const float *iso,
                       int n_iso, const int16_t *htable)
{
    const int m = pl->m, ndirs = pl->ndirs;
    pl->n = n0 + n1 + (with_dot ? 1 : 0) + n_iso;
    pl->n_rot = n0 + n1;
    slab_geometry(pl, sizeof(float));
    float *d_rot0 = nullptr, *d_rot1 = nullptr, *d_iso = nullptr;
    int rc;
    if ((rc = upload(&d_rot0, rot0, (size_t)n0 * ndirs * m))) return rc;
    if (n1 && (rc = upload(&d_rot1, rot1, (size_t)n1 * ndirs * m))) return rc;
    if ((rc = upload(&d_iso, iso, (size_t)n_iso * m))) return rc;
    if ((rc = upload(&pl->d_htable, htable, (size_t)181 * 181))) return rc;
    CK(cudaMalloc(&pl->d_slab, (size_t)ndirs * pl->slab_stride * sizeof(float) + 1024));
    CK(cudaMemsetAsync(pl->d_slab, 0, (size_t)ndirs * pl->slab_stride * sizeof(float) + 1024, pl->stream));
    k_build_slab<<<ndirs, 256, 0, pl->stream>>>(d_rot0, n0, d_rot1, n1, with_dot, d_iso, n_iso, ndirs, m, pl->n_pad,
                                                 pl->slab_stride, (float *)pl->d_slab);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(pl->stream));
    cudaFree(d_rot0);
    if (d_rot1) cudaFree(d_rot1);
    cudaFree(d_iso);
    return AMX_OK;
}

int check_common(int m, int ndirs, const void *a, const void *b, amx_plan **out)
{
    if (!out) return fail(AMX_E_INVALID, "out is NULL");
    *out = nullptr;
    if (m <= 0 || ndirs <= 0) return fail(AMX_E_INVALID, "m and ndirs must be positive (m=%d, ndirs=%d)", m, ndirs);
    if (!a || !b) return fail(AMX_E_INVALID, "NULL kernel table");
    return AMX_OK;
}

}  // namespace

extern "C" {

const char *amx_last_error(void) { return g_err.c_str(); }
int amx_version(void) { return AMX_VERSION; }

int amx_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(AMX_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int amx_plan_destroy(amx_plan *pl)
{
    if (!pl) return AMX_OK;
    cudaSetDevice(pl->device);
    void *ptrs[] = {pl->d_slab, pl->d_T1, pl->d_T2, pl->d_diag0, pl->d_W, pl->d_htable, pl->d_dwi_rows, pl->d_norms, pl->d_icvf, pl->d_kappa,
                    pl->d_Rs, pl->d_sandi_norms, pl->d_d_in, pl->d_d_isos};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (auto &wk : pl->work) {
        DevBuf *bufs[] = {&wk.lut, &wk.order, &wk.bins, &wk.tiles, &wk.status, &wk.scratch, &wk.xiso, &wk.supmask, &wk.ovf_list, &wk.slow_ws,
                          &wk.exact_list, &wk.exact_a, &wk.c1_all, &wk.redo};
        for (DevBuf *b : bufs) b->release();
    }
    for (void *h : pl->hpin) if (h) cudaFreeHost(h);
    for (void *h : pl->hout) if (h) cudaFreeHost(h);
    if (pl->cs[1]) cudaStreamDestroy(pl->cs[1]);
    if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
    for (auto &e : pl->ev_join) if (e) cudaEventDestroy(e);
    for (auto &sg : pl->stg) {
        DevBuf *sb[] = {&sg.y, &sg.dirs, &sg.est, &sg.rmse, &sg.nrmse, &sg.extra, &sg.sup, &sg.coef, &sg.lut};
        for (DevBuf *b : sb) b->release();
    }
    for (auto &ev : pl->ev) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        if (pl->ev_in[i]) cudaEventDestroy(pl->ev_in[i]);
        if (pl->ev_comp[i]) cudaEventDestroy(pl->ev_comp[i]);
        if (pl->ev_out[i]) cudaEventDestroy(pl->ev_out[i]);
        if (pl->ev_lut[i]) cudaEventDestroy(pl->ev_lut[i]);
    }
    if (pl->stream) cudaStreamDestroy(pl->stream);
    if (pl->s_in) cudaStreamDestroy(pl->s_in);
    if (pl->s_out) cudaStreamDestroy(pl->s_out);
    delete pl;
    return AMX_OK;
}

int amx_plan_create_noddi(int device, int m, int ndirs, int n_wm, const float *wm, const float *iso, const double *norms,
                          const float *icvf, const float *kappa, const int64_t *dwi_idx, int dwi_count, int is_exvivo,
                          const int16_t *htable, amx_plan **out)
{
    int rc = check_common(m, ndirs, wm, iso, out);
    if (rc) return rc;
    if (n_wm <= 0 || dwi_count <= 0 || dwi_count > m || !norms || !icvf || !kappa || !htable || (!dwi_idx && m != 1 + dwi_count))
        return fail(AMX_E_INVALID, "bad NODDI tables (n_wm=%d, dwi_count=%d, m=%d)", n_wm, dwi_count, m);
    if (n_wm + 2 > 32 * 8) return fail(AMX_E_INVALID, "n_wm=%d exceeds the supported 254 atoms", n_wm);
    amx_plan *pl = new amx_plan;
    pl->model = AMX_MODEL_NODDI; pl->m = m; pl->ndirs = ndirs; pl->n_wm = n_wm; pl->exvivo = is_exvivo ? 1 : 0;
    pl->n_maps = is_exvivo ? 4 : 3; pl->dc = dwi_count;
    if ((rc = plan_common_init(pl, device))) { amx_plan_destroy(pl); return rc; }
    if ((rc = build_rotated_plan(pl, wm, n_wm, nullptr, 0, pl->exvivo, iso, 1, htable))) { amx_plan_destroy(pl); return rc; }
    // DWI rows: scheme.dwi_idx, or rows 1..m-1 in the single-b0 case (amico/models.pyx:916-918)
    std::vector<int> rows(dwi_count);
    for (int j = 0; j < dwi_count; ++j) {
        long long r = (m == 1 + dwi_count) ? j + 1 : (long long)dwi_idx[j];
        if (r < 0 || r >= m) { amx_plan_destroy(pl); return fail(AMX_E_INVALID, "dwi_idx[%d]=%lld outside [0,%d)", j, r, m); }
        rows[j] = (int)r;
    }
    // norms rows are identical by construction (models.pyx:781-784); detect it to keep them in registers
    pl->norms_const = 1;
    for (int j = 1; j < dwi_count && pl->norms_const; ++j)
        if (memcmp(norms, norms + (size_t)j * n_wm, sizeof(double) * n_wm) != 0) pl->norms_const = 0;
    if ((rc = upload(&pl->d_dwi_rows, rows.data(), rows.size())) || (rc = upload(&pl->d_norms, norms, (size_t)dwi_count * n_wm)) ||
        (rc = upload(&pl->d_icvf, icvf, (size_t)n_wm)) || (rc = upload(&pl->d_kappa, kappa, (size_t)n_wm))) {
        amx_plan_destroy(pl);
        return rc;
    }
    if ((rc = build_gram(pl, &pl->d_T1, pl->n, &pl->ldT1, &pl->T1_stride, nullptr, m, nullptr, 0, 0)) ||
        (rc = build_gram(pl, &pl->d_T2, n_wm, &pl->ldT2, &pl->T2_stride, pl->d_dwi_rows, dwi_count, pl->d_norms, n_wm, pl->norms_const))) {
        amx_plan_destroy(pl);
        return rc;
    }
    pl->K2 = n_wm;
    cudaError_t e = cudaStreamSynchronize(pl->stream);
    if (e != cudaSuccess) { amx_plan_destroy(pl); return fail(AMX_E_CUDA, "table build failed: %s", cudaGetErrorString(e)); }
    *out = pl;
    return AMX_OK;
}

int amx_plan_create_freewater(int device, int m, int ndirs, int n_perp, const float *D, int n_iso, const float *CSF,
                              int is_mouse, const int16_t *htable, amx_plan **out)
{
    int rc = check_common(m, ndirs, D, CSF, out);
    if (rc) return rc;
    if (n_perp <= 0 || n_iso <= 0 || !htable || (is_mouse && n_iso < 2)) return fail(AMX_E_INVALID, "bad FreeWater tables (n_perp=%d, n_iso=%d)", n_perp, n_iso);
    if (n_perp + n_iso > 256) return fail(AMX_E_INVALID, "too many atoms (%d)", n_perp + n_iso);
    amx_plan *pl = new amx_plan;
    pl->model = AMX_MODEL_FREEWATER; pl->m = m; pl->ndirs = ndirs; pl->n_perp = n_perp; pl->n_iso = n_iso; pl->mouse = is_mouse ? 1 : 0;
    pl->n_maps = is_mouse ? 4 : 2;
    if ((rc = plan_common_init(pl, device)) || (rc = build_rotated_plan(pl, D, n_perp, nullptr, 0, 0, CSF, n_iso, htable)) ||
        (rc = build_gram(pl, &pl->d_T2, pl->n, &pl->ldT2, &pl->T2_stride, nullptr, m, nullptr, 0, 0))) {
        amx_plan_destroy(pl);
        return rc;
    }
    pl->K2 = pl->n;
    cudaError_t e = cudaStreamSynchronize(pl->stream);
    if (e != cudaSuccess) { amx_plan_destroy(pl); return fail(AMX_E_CUDA, "table build failed: %s", cudaGetErrorString(e)); }
    *out = pl;
    return AMX_OK;
}

int amx_plan_create_czb(int device, int m, int ndirs, int n_rs, const float *wmr, int n_perp, const float *wmh, int n_iso,
                        const float *iso, const double *Rs, const int16_t *htable, amx_plan **out)
{
    int rc = check_common(m, ndirs, wmr, iso, out);
    if (rc) return rc;
    if (n_rs <= 0 || n_perp < 0 || n_iso <= 0 || !Rs || !htable || (n_perp && !wmh))
        return fail(AMX_E_INVALID, "bad CylinderZeppelinBall tables (n_rs=%d, n_perp=%d, n_iso=%d)", n_rs, n_perp, n_iso);
    if (n_rs + n_perp + n_iso > 256) return fail(AMX_E_INVALID, "too many atoms (%d)", n_rs + n_perp + n_iso);
    amx_plan *pl = new amx_plan;
    pl->model = AMX_MODEL_CZB; pl->m = m; pl->ndirs = ndirs; pl->n_rs = n_rs; pl->n_perp = n_perp; pl->n_iso = n_iso; pl->n_maps = 3;
    if ((rc = plan_common_init(pl, device)) || (rc = build_rotated_plan(pl, wmr, n_rs, wmh, n_perp, 0, iso, n_iso, htable)) ||
        (rc = upload(&pl->d_Rs, Rs, (size_t)n_rs)) ||
        (rc = build_gram(pl, &pl->d_T2, pl->n, &pl->ldT2, &pl->T2_stride, nullptr, m, nullptr, 0, 0))) {
        amx_plan_destroy(pl);
        return rc;
    }
    pl->K2 = pl->n;
    cudaError_t e = cudaStreamSynchronize(pl->stream);
    if (e != cudaSuccess) { amx_plan_destroy(pl); return fail(AMX_E_CUDA, "table build failed: %s", cudaGetErrorString(e)); }
    *out = pl;
    return AMX_OK;
}

int amx_plan_create_sandi(int device, int m, int n_rs, int n_in, int n_iso, const double *signal, const double *norms,
                          const double *Rs, const double *d_in, const double *d_isos, amx_plan **out)
{
    int rc = check_common(m, 1, signal, norms, out);
    if (rc) return rc;
    if (n_rs < 0 || n_in < 0 || n_iso < 0 || n_rs + n_in + n_iso <= 0 || !Rs || !d_in || !d_isos)
        return fail(AMX_E_INVALID, "bad SANDI tables");
    if (n_rs + n_in + n_iso > 256) return fail(AMX_E_INVALID, "too many atoms (%d)", n_rs + n_in + n_iso);
    amx_plan *pl = new amx_plan;
    pl->model = AMX_MODEL_SANDI; pl->m = m; pl->ndirs = 1; pl->n_rs = n_rs; pl->n_in = n_in; pl->n_iso = n_iso; pl->n_maps = 6;
    pl->n = n_rs + n_in + n_iso; pl->slab_f64 = true;
    if ((rc = plan_common_init(pl, device))) { amx_plan_destroy(pl); return rc; }
    slab_geometry(pl, sizeof(double));
    double *d_A = nullptr;
    if ((rc = upload(&d_A, signal, (size_t)m * pl->n)) || (rc = upload(&pl->d_sandi_norms, norms, (size_t)pl->n)) ||
        (rc = upload(&pl->d_Rs, Rs, (size_t)std::max(n_rs, 1))) || (rc = upload(&pl->d_d_in, d_in, (size_t)std::max(n_in, 1))) ||
        (rc = upload(&pl->d_d_isos, d_isos, (size_t)std::max(n_iso, 1)))) {
        amx_plan_destroy(pl);
        return rc;
    }
    cudaError_t e = cudaMalloc(&pl->d_slab, pl->slab_stride * sizeof(double) + 1024);
    if (e == cudaSuccess) e = cudaMemsetAsync(pl->d_slab, 0, pl->slab_stride * sizeof(double) + 1024, pl->stream);
    if (e != cudaSuccess) { amx_plan_destroy(pl); return fail(AMX_E_CUDA, "slab alloc: %s", cudaGetErrorString(e)); }
    k_build_slab_f64<<<1, 256, 0, pl->stream>>>(d_A, m, pl->n, pl->n_pad, (double *)pl->d_slab);
    if ((rc = build_gram(pl, &pl->d_T2, pl->n, &pl->ldT2, &pl->T2_stride, nullptr, m, nullptr, 0, 0))) { amx_plan_destroy(pl); return rc; }
    pl->K2 = pl->n;
    e = cudaStreamSynchronize(pl->stream);
    cudaFree(d_A);
    if (e != cudaSuccess) { amx_plan_destroy(pl); return fail(AMX_E_CUDA, "table build failed: %s", cudaGetErrorString(e)); }
    *out = pl;
    return AMX_OK;
}

int amx_plan_info(const amx_plan *pl, int *model, int *m, int *n_atoms, int *n_maps, int *ndirs, int *device)
{
    if (!pl) return fail(AMX_E_INVALID, "plan is NULL");
    if (model) *model = pl->model;
    if (m) *m = pl->m;
    if (n_atoms) *n_atoms = pl->n;
    if (n_maps) *n_maps = pl->n_maps;
    if (ndirs) *ndirs = pl->ndirs;
    if (device) *device = pl->device;
    return AMX_OK;
}

}  // extern "C"

namespace {

template <int NPL, int MAXT>
int launch_noddi_split_t(const FitParams &p, int grid, int block, size_t smem, cudaStream_t st)
{
    auto k1 = k_noddi_stage<1, NPL, float, MAXT>;
    auto k2 = k_noddi_stage<2, NPL, float, MAXT>;
    auto k3 = k_noddi_stage<3, NPL, float, MAXT>;
    const size_t fixed = p.ws_smem_off;
    const size_t s1 = fixed + (size_t)p.ws_doubles_stage[0] * 8 * p.nwarps, s2 = fixed + (size_t)p.ws_doubles_stage[1] * 8 * p.nwarps,
                 s3 = fixed + (size_t)p.ws_doubles_stage[2] * 8 * p.nwarps;
    (void)smem;
    CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
    CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
    CK(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3));
    // carve-out hint: what is not shared memory is L1 (Gram rows)
    CK(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<size_t>(100, (s1 * 100 + 233471) / 233472 + 1)));
    // Stage 1's workspace is small (active sets <= 16): 32 warps fit next to ~90 KB of L1 when the kernel is compiled for
    // 1024 threads (64 registers, a few spilled words); more resident warps hide the L2 latency of the Gram rows.
    const int w1 = env_int("AMX_STAGE1_WARPS", 32);
    const size_t s1w = fixed + (size_t)p.ws_doubles_stage[0] * 8 * 32;
    const bool pair1 = env_int("AMX_PAIR1", 0) && MAXT == 768 && block == 768 && p.cap_stage[0] <= 16;
    const bool lean1 = env_int("AMX_LEAN1", 1) && MAXT == 768 && block == 768 && p.cap_stage[0] <= 16;
    // stage 1 with one voxel per thread and warp-cooperative dual passes (amx_lean.cuh); voxels it hands back (passive set > CAPT)
    // run through the warp-per-voxel lean kernel as one-voxel tiles
    if (p.tpv1) {
        const int capt = env_int("AMX_TPV1_CAP", 8);
        auto launch_tpv = [&](auto kern, int smem_t) -> int {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_t));
            const int ctas = std::max(1, std::min(env_int("AMX_TPV1_CTAS", 4), (int)(232448 / (smem_t + 1024))));
            kern<<<grid * ctas, TPV_THREADS, smem_t, st>>>(p, p.redo_tiles, p.redo_count);
            return AMX_OK;
        };
        int rc = capt <= 6   ? launch_tpv(k_noddi_stage1_tpv<NPL, 6>, tpv1_smem_bytes<NPL, 6>())
                 : capt == 7 ? launch_tpv(k_noddi_stage1_tpv<NPL, 7>, tpv1_smem_bytes<NPL, 7>())
                             : launch_tpv(k_noddi_stage1_tpv<NPL, 8>, tpv1_smem_bytes<NPL, 8>());
        if (rc) return rc;
        FitParams p1 = p;
        p1.tiles = p.redo_tiles;
        p1.n_tiles_ptr = p.redo_count;
        p1.tile_counter = p.redo_count + 1;
        const int warps = 32;
        const size_t sw = fixed + (size_t)(LeanWS<16>::SIZE + 32 * NPL) * 8 * warps;
        auto kern = k_noddi_stage1_lean<NPL, 1024, 16>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw));
        kern<<<grid, warps * 32, sw, st>>>(p1);
    } else
    if (pair1) {
        const int warps = (w1 == 24 || w1 == 28) ? w1 : 32;
        const size_t sw = fixed + (size_t)(2 * (PairWS<16>::CS + 16 * 2 * NPL) + BV) * 8 * warps;
        auto launch_pair = [&](auto kern) -> int {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw));
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<size_t>(100, (sw * 100 + 233471) / 233472 + 1)));
            kern<<<grid, warps * 32, sw, st>>>(p);
            return AMX_OK;
        };
        int rc;
        if (warps == 32) rc = launch_pair(k_noddi_stage1_pair<2 * NPL, 1024, 16>);
        else if (warps == 28) rc = launch_pair(k_noddi_stage1_pair<2 * NPL, 896, 16>);
        else rc = launch_pair(k_noddi_stage1_pair<2 * NPL, 768, 16>);
        if (rc) return rc;
    } else if (lean1) {  // one voxel per warp, lean solver (same arithmetic as k_noddi_stage<1>, fewer instructions)
        const int warps = (w1 == 24 || w1 == 28) ? w1 : 32;
        const size_t sw = fixed + (size_t)(LeanWS<16>::SIZE + 32 * NPL) * 8 * warps;
        auto launch_lean = [&](auto kern) -> int {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw));
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<size_t>(100, (sw * 100 + 233471) / 233472 + 1)));
            kern<<<grid, warps * 32, sw, st>>>(p);
            return AMX_OK;
        };
        int rc;
        if (warps == 32) rc = launch_lean(k_noddi_stage1_lean<NPL, 1024, 16>);
        else if (warps == 28) rc = launch_lean(k_noddi_stage1_lean<NPL, 896, 16>);
        else rc = launch_lean(k_noddi_stage1_lean<NPL, 768, 16>);
        if (rc) return rc;
    } else if (MAXT == 768 && block == 768 && w1 == 32 && s1w <= 227 * 1024) {
        auto k1w = k_noddi_stage<1, NPL, float, 1024>;
        CK(cudaFuncSetAttribute(k1w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1w));
        CK(cudaFuncSetAttribute(k1w, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<size_t>(100, (s1w * 100 + 233471) / 233472 + 1)));
        k1w<<<grid, 1024, s1w, st>>>(p);
    } else {
        k1<<<grid, block, s1, st>>>(p);
    }
    // Stages 2 and 3: more resident warps hide more of the L2 latency of the Gram rows as long as registers (64K / threads)
    // and shared memory (workspace x warps <= 227 KB) allow; the 1024-thread builds spill ~150 bytes.
    const int w2 = env_int("AMX_STAGE2_WARPS", 32), w3 = env_int("AMX_STAGE3_WARPS", 24);
    auto launch_wide = [&](auto kern, int warps, unsigned ws_doubles) -> int {
        const size_t sw = fixed + (size_t)ws_doubles * 8 * warps;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw));
        kern<<<grid, warps * 32, sw, st>>>(p);
        return AMX_OK;
    };
    const bool wide_ok = MAXT == 768 && block == 768;
    auto fits = [&](int warps, unsigned ws_doubles) { return fixed + (size_t)ws_doubles * 8 * warps <= (size_t)227 * 1024; };
    int rc = AMX_OK;
    if (env_int("AMX_LEAN2", 1) && p.fast_lars && wide_ok && p.NA == 32 * NPL) {  // inlined stage-2 solver (amx_lean.cuh), own workspace layout
        const unsigned wsd = Lars2WS<NPL>::SIZE;
        if (w2 == 24) rc = launch_wide(k_noddi_stage2_lean<NPL, 768>, 24, wsd);
        else if (w2 == 32 && fits(32, wsd)) rc = launch_wide(k_noddi_stage2_lean<NPL, 1024>, 32, wsd);
        else rc = launch_wide(k_noddi_stage2_lean<NPL, 896>, 28, wsd);
    } else if (wide_ok && w2 == 32 && fits(32, p.ws_doubles_stage[1])) rc = launch_wide(k_noddi_stage<2, NPL, float, 1024>, 32, p.ws_doubles_stage[1]);
    else k2<<<grid, block, s2, st>>>(p);
    if (rc) return rc;
    const bool tpv3 = p.tpv3 != 0;
    FitParams p3 = p;
    if (tpv3) {
        // one voxel per thread (amx_lean.cuh); what it hands back runs through the warp-per-voxel kernel as one-voxel tiles
        const long long blocks = std::max<long long>(1, std::min<long long>((p.n_vox + TPV_THREADS - 1) / TPV_THREADS, (long long)grid * 64));
        auto launch_t3 = [&](auto kt, int smem_t) -> int {
            CK(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_t));
            kt<<<(int)blocks, TPV_THREADS, smem_t, st>>>(p, p.redo_tiles, p.redo_count + 2);
            return AMX_OK;
        };
        // AMX_TPV3_CAP = 3: test hook, most voxels outgrow the per-thread capacity and take the hand-back path
        const int cap3 = env_int("AMX_TPV3_CAP", 5);
        if (int rc3 = cap3 <= 3   ? launch_t3(k_noddi_stage3_tpv<NPL, 3>, tpv3_smem_bytes<3>())
                      : cap3 == 5 ? launch_t3(k_noddi_stage3_tpv<NPL, 5>, tpv3_smem_bytes<5>())
                                  : launch_t3(k_noddi_stage3_tpv<NPL, 6>, tpv3_smem_bytes<6>()))
            return rc3;
        p3.tiles = p.redo_tiles;
        p3.n_tiles_ptr = p.redo_count + 2;
        p3.tile_counter = p.redo_count + 3 - 2;  // stage 3 pulls from tile_counter[2]
    }
    if (wide_ok && w3 == 32 && fits(32, p.ws_doubles_stage[2])) {
        const size_t sw = fixed + (size_t)p.ws_doubles_stage[2] * 8 * 32;
        auto kern = k_noddi_stage<3, NPL, float, 1024>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw));
        kern<<<grid, 1024, sw, st>>>(p3);
    } else k3<<<grid, block, s3, st>>>(p3);
    if (rc) return rc;
    CK(cudaGetLastError());
    return AMX_OK;
}

template <int NPL>
int launch_noddi_split(const FitParams &p, int grid, int block, size_t smem, cudaStream_t st)
{
    if (block > 512) return launch_noddi_split_t<NPL, 768>(p, grid, block, smem, st);
    return launch_noddi_split_t<NPL, 512>(p, grid, block, smem, st);
}

template <int NPL>
int launch_noddi_exact(const FitParams &p, int grid, size_t smem, long long *status, double *scratch_a, cudaStream_t st)
{
    auto kern = k_noddi_exact<NPL, float>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 32 * NPL, smem, st>>>(p, p.exact_list, status, scratch_a);
    CK(cudaGetLastError());
    return AMX_OK;
}

template <int MODEL, int NPL, typename TS>
int launch_fit(const FitParams &p, int grid, int block, size_t smem, cudaStream_t st)
{
    auto kern = k_fit<MODEL, NPL, TS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(p);
    CK(cudaGetLastError());
    return AMX_OK;
}

template <int MODEL, int NPL, typename TS, int MAXT>
int launch_lasso_batched(const FitParams &p, int grid, int block, size_t smem, cudaStream_t st)
{
    auto kern = k_lasso_batched<MODEL, NPL, TS, MAXT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(p);
    CK(cudaGetLastError());
    return AMX_OK;
}

template <int MODEL, typename TS>
int dispatch_npl(int npl, const FitParams &p, int grid, int block, size_t smem, cudaStream_t st)
{
    if constexpr (MODEL == MODEL_FREEWATER || MODEL == MODEL_SANDI) {
        if (p.batched == 4) {  // thread-per-voxel kernel for dictionaries of <= 16 atoms (amx_small.cuh); grid = n_vox / 128, no shared memory
            const int blocks = (int)std::max<long long>(1, std::min<long long>((p.n_vox + 127) / 128, (long long)grid * 16));
            if (std::min(p.m, p.n) <= 4) k_lasso_small<MODEL, TS, 16, 4><<<blocks, 128, 0, st>>>(p);
            else k_lasso_small<MODEL, TS, 16, 12><<<blocks, 128, 0, st>>>(p);
            CK(cudaGetLastError());
            return AMX_OK;
        }
    }
    if (MODEL != MODEL_NODDI && p.batched == 3) {
        switch (npl) {
        case 1: return block > 768 ? launch_lasso_batched<MODEL, 1, TS, 1024>(p, grid, block, smem, st)
                                   : launch_lasso_batched<MODEL, 1, TS, 768>(p, grid, block, smem, st);
        case 2: return launch_lasso_batched<MODEL, 2, TS, 768>(p, grid, block, smem, st);
        case 3: case 4: return launch_lasso_batched<MODEL, 4, TS, 768>(p, grid, block, smem, st);
        }
    }
    if (MODEL == MODEL_NODDI && p.batched == 2) {
        switch (npl) {
        case 1: return launch_noddi_split<1>(p, grid, block, smem, st);
        case 2: return launch_noddi_split<2>(p, grid, block, smem, st);
        case 3: case 4: return launch_noddi_split<4>(p, grid, block, smem, st);
        case 5: return launch_noddi_split<5>(p, grid, block, smem, st);
        }
    }
    if constexpr (MODEL == MODEL_NODDI) {
        // dictionaries of up to 160 atoms run as stage kernels; only larger ones (user grids, <= 254 atoms) take the per-voxel kernel
        if (npl >= 6 && npl <= 8) return launch_fit<MODEL, 8, TS>(p, grid, block, smem, st);
    } else {
        switch (npl) {
        case 1: return launch_fit<MODEL, 1, TS>(p, grid, block, smem, st);
        case 2: return launch_fit<MODEL, 2, TS>(p, grid, block, smem, st);
        case 3: case 4: return launch_fit<MODEL, 4, TS>(p, grid, block, smem, st);
        case 5: return launch_fit<MODEL, 5, TS>(p, grid, block, smem, st);
        case 6: case 7: case 8: return launch_fit<MODEL, 8, TS>(p, grid, block, smem, st);
        }
    }
    return fail(AMX_E_INVALID, "unsupported atom count (npl=%d)", npl);
}

// Per-call table state that every chunk shares: ridge baked into the LARS Gram diagonal, active-set cap constant.
// Runs on the caller's stream BEFORE the chunk streams fork.
int prepare_tables(amx_plan *pl, const amx_fit_args *a, cudaStream_t st, int *launches)
{
    {
        const double ridge = a->lambda2 > 1e-10 ? a->lambda2 : 1e-10;
        if (ridge != pl->ridge_baked) {
            if (!pl->d_diag0) {
                CK(cudaMalloc((void **)&pl->d_diag0, (size_t)pl->ndirs * pl->K2 * sizeof(double)));
                k_save_diag<<<pl->ndirs, 128, 0, st>>>(pl->d_T2, pl->K2, pl->ldT2, pl->T2_stride, pl->d_diag0);
            }
            k_set_ridge<<<pl->ndirs, 128, 0, st>>>(pl->d_T2, pl->K2, pl->ldT2, pl->T2_stride, pl->d_diag0, ridge);
            CK(cudaGetLastError());
            pl->ridge_baked = ridge;
            *launches += 1;
            if (pl->model == AMX_MODEL_CZB && pl->n <= 32) {  // inverse tables of the dense-start NNQP follow the ridge
                if (!pl->d_W) {
                    pl->ldW = pl->ldT2; pl->W_stride = pl->T2_stride;
                    CK(cudaMalloc((void **)&pl->d_W, ((size_t)pl->ndirs * pl->W_stride + 256) * sizeof(double)));
                    CK(cudaMemsetAsync(pl->d_W, 0, ((size_t)pl->ndirs * pl->W_stride + 256) * sizeof(double), st));
                }
                k_invert_spd<<<pl->ndirs, 32, 0, st>>>(pl->d_T2, pl->n, pl->ldT2, pl->T2_stride, pl->d_W, pl->ldW, pl->W_stride);
                CK(cudaGetLastError());
                *launches += 1;
            }
        }
    }
    {
        const int want = std::max(1, std::min(LC, env_int("AMX_LC_CAP", LC)));
        static int cap_on_device[64];  // value of the constant on each device (0 = the compiled default LC)
        int &cur = cap_on_device[pl->device & 63];
        if (want != (cur ? cur : LC)) {
            CK(cudaMemcpyToSymbolAsync(c_lc_cap, &want, sizeof(int), 0, cudaMemcpyHostToDevice, st));
            cur = want;
        }
    }
    return AMX_OK;
}

// Enqueue LUT index + binning + fit kernels for device-resident voxels; fully asynchronous (the tile count stays on the
// device, the error / overflow words are read by the caller at the end).  All pointers are device pointers.
int fit_device(amx_plan *pl, amx_plan::Work &wk, const amx_fit_args *a, cudaStream_t st, int *launches, long long vox_offset,
               bool record_events, cudaEvent_t after_lut = nullptr)
{
    const long long n_vox = a->n_vox;
    const bool batched = pl->model == AMX_MODEL_NODDI && pl->npl <= 5;  // larger dictionaries: the per-voxel kernel (k_fit)
    // single-fit models: DMMA-batched throughput kernel unless the caller asks for the bit-reproducible one
    const bool lasso_fast = pl->model != AMX_MODEL_NODDI && pl->npl <= 4 && !(a->flags & AMX_FLAG_EXACT) && env_int("AMX_LASSO_FAST", 1);
    // ... and one voxel per THREAD when the dictionary is tiny (FreeWater, SANDI)
    // (measured: SANDI, <= 4 active atoms, 860 M voxels/s against 317 M warp-per-voxel; FreeWater, <= 11 active atoms and 1.7 KB of
    //  thread-local state, 44 M against 95 M -- so only paths of <= AMX_LASSO_SMALL_L atoms take it)
    const bool lasso_small = lasso_fast && (pl->model == AMX_MODEL_FREEWATER || pl->model == AMX_MODEL_SANDI) && pl->n <= 16 &&
                             std::min(pl->m, pl->n) <= std::min(12, env_int("AMX_LASSO_SMALL_L", 4)) && env_int("AMX_LASSO_SMALL", 1);
    const int tile_v = (batched || lasso_fast) ? BV : std::max(1, env_int("AMX_TILE_VOX", 256));
    const bool rotated = pl->model != AMX_MODEL_SANDI;
    const long long max_tiles = n_vox / tile_v + pl->ndirs + 1;
    CK(wk.tiles.reserve((size_t)max_tiles * sizeof(int4)));
    CK(wk.bins.reserve(((size_t)4 * pl->ndirs + 8) * sizeof(int)));
    int *bins = (int *)wk.bins.p;
    int *hist = bins, *offs = bins + pl->ndirs, *cursor = bins + 2 * pl->ndirs, *tile_offs = bins + 3 * pl->ndirs,
        *totals = bins + 4 * pl->ndirs;  // totals[0]=n_tiles, [1]=n binned, [2..4]=tile counters
    long long *status = (long long *)wk.status.p;
    CK(cudaMemsetAsync(wk.bins.p, 0, ((size_t)4 * pl->ndirs + 8) * sizeof(int), st));
    if (record_events) CK(cudaEventRecord(pl->ev[0], st));
    long long n_tiles_bound = max_tiles;
    int *lut = nullptr;
    if (rotated) {
        if (!a->dirs) return fail(AMX_E_INVALID, "dirs is NULL");
        if (a->lut_out) lut = a->lut_out;
        else { CK(wk.lut.reserve((size_t)n_vox * sizeof(int))); lut = (int *)wk.lut.p; }
        CK(wk.order.reserve((size_t)n_vox * sizeof(int)));
        const int B = 256;
        const unsigned G = (unsigned)((n_vox + B - 1) / B);
        k_lut<<<G, B, 0, st>>>(a->dirs, n_vox, pl->d_htable, pl->ndirs, lut, hist, status, vox_offset);
        k_scan_bins<<<1, 1024, 0, st>>>(hist, pl->ndirs, tile_v, offs, cursor, tile_offs, totals);
        k_scatter<<<G, B, 0, st>>>(lut, n_vox, cursor, (int *)wk.order.p);
        k_tiles<<<(pl->ndirs + 127) / 128, 128, 0, st>>>(hist, offs, tile_offs, pl->ndirs, tile_v, (int4 *)wk.tiles.p, 0);
        CK(cudaGetLastError());
        *launches += 4;
    } else {
        const int n_tiles = (int)((n_vox + tile_v - 1) / tile_v);
        n_tiles_bound = n_tiles;
        k_tiles_linear<<<(n_tiles + 127) / 128, 128, 0, st>>>(n_vox, tile_v, (int4 *)wk.tiles.p, n_tiles, totals);
        CK(cudaGetLastError());
        *launches += 1;
        if (a->lut_out) CK(cudaMemsetAsync(a->lut_out, 0, (size_t)n_vox * sizeof(int), st));
    }
    if (record_events) CK(cudaEventRecord(pl->ev[1], st));
    if (after_lut) CK(cudaEventRecord(after_lut, st));  // dirs are final (flipped) from here on

    FitParams p;
    memset(&p, 0, sizeof p);
    p.model = pl->model; p.m = pl->m; p.n = pl->n; p.n_pad = pl->n_pad; p.ndirs = pl->ndirs; p.n_maps = pl->n_maps;
    p.NA = 32 * (pl->npl == 3 ? 4 : (pl->npl == 6 || pl->npl == 7) ? 8 : pl->npl);
    p.slab = pl->d_slab; p.slab_stride = pl->slab_stride;
    p.T1 = pl->d_T1; p.ldT1 = pl->ldT1; p.T1_stride = pl->T1_stride;
    p.T2 = pl->d_T2; p.ldT2 = pl->ldT2; p.T2_stride = pl->T2_stride; p.K2 = pl->K2;
    if (pl->d_W && a->lambda1 == 0.0 && pl->m >= pl->n && env_int("AMX_CZB_DENSE", 1)) { p.W = pl->d_W; p.ldW = pl->ldW; p.W_stride = pl->W_stride; }
    p.y = a->y; p.y_f64 = a->y_dtype == AMX_F64; p.n_vox = n_vox;
    p.order = rotated ? (const int *)wk.order.p : nullptr; p.tiles = (const int4 *)wk.tiles.p; p.n_tiles_ptr = totals;
    p.tile_counter = totals + 2;
    p.lambda1 = a->lambda1; p.lambda2 = a->lambda2; p.flags = a->flags;
    p.dwi_rows = pl->d_dwi_rows; p.dc = pl->dc; p.norms = pl->d_norms; p.norms_const = pl->norms_const;
    p.icvf = pl->d_icvf; p.kappa = pl->d_kappa; p.exvivo = pl->exvivo; p.n_wm = pl->n_wm;
    p.mouse = pl->mouse; p.n_perp = pl->n_perp; p.n_iso = pl->n_iso; p.n_rs = pl->n_rs; p.n_in = pl->n_in;
    p.Rs = pl->d_Rs; p.sandi_norms = pl->d_sandi_norms; p.d_in = pl->d_d_in; p.d_isos = pl->d_d_isos;
    p.lut = lut;
    p.est = a->estimates; p.rmse = a->rmse; p.nrmse = a->nrmse; p.extra = a->extra; p.support_out = a->support_out; p.coeff_out = a->coeff_out;
    p.status = status;
    p.batched = batched ? 2 : lasso_small ? 4 : lasso_fast ? 3 : 0;
    p.fast_lars = env_int("AMX_FAST_LARS", 1);
    p.compact3 = env_int("AMX_COMPACT3", 1);
    p.aspace = env_int("AMX_ASPACE", 1);
    p.m_pad = (pl->m + 1) & ~1; p.dc_pad = p.batched ? 0 : (pl->dc + 1) & ~1;
    if (p.batched && !(a->flags & (AMX_FLAG_RMSE | AMX_FLAG_NRMSE))) p.m_pad = 0;
    if (p.batched == 3 && (a->flags & AMX_FLAG_EXTRA)) p.m_pad = (pl->m + 1) & ~1;  // FreeWater corrected DWI reads the signal
    p.ws_doubles = ws_doubles_for(p.NA, p.m_pad, p.dc_pad, p.batched == 2 ? 1 : 0);

    // shared-memory budget: [header 128][slab (optional)][nwarps x workspace]
    if (p.batched == 3 || p.batched == 4) {
        p.cap_stage[1] = std::max(4, std::min(LC, std::min(pl->n, std::min(pl->m, pl->n)) + 1));  // the path never holds more than min(m, n) atoms
        if (p.W) p.cap_stage[1] = std::max(p.cap_stage[1], 18);  // the block-pivoting scratch (NNQP_KMAX x (NNQP_KMAX + 1)) lives in the same matrix area
        p.ws_doubles_stage[1] = ws_doubles_for(p.NA, p.m_pad, 0, 1, p.cap_stage[1]);
        p.ws_doubles = p.ws_doubles_stage[1];
    }
    if (p.batched == 2) {
        // stage 1 (NNLS on the full dictionary) never holds more than ~8 passive atoms on NODDI dictionaries (rank ~11):
        // a 16-atom workspace leaves ~85 KB more L1 for its Gram rows; rarer larger sets go to the slow path
        p.cap_stage[0] = std::max(4, std::min(LC, env_int("AMX_CAP_STAGE1", 16)));
        p.cap_stage[1] = std::max(4, std::min(LC, env_int("AMX_CAP_STAGE2", LC)));
        p.cap_stage[2] = std::max(4, std::min(LC, env_int("AMX_CAP_STAGE3", LC)));
        for (int k = 0; k < 3; ++k) p.ws_doubles_stage[k] = ws_doubles_for(p.NA, p.m_pad, p.dc_pad, (k == 1 && p.fast_lars) ? 2 : 1, p.cap_stage[k]);
    }
    const size_t ws_bytes = (size_t)p.ws_doubles * sizeof(double);
    const size_t budget = (size_t)pl->max_smem;
    const int max_warps = lasso_fast ? (pl->npl == 1 ? 32 : 24) : batched ? 24 : 16;
    const int want_warps = std::min(max_warps, std::max(1, env_int("AMX_WARPS", max_warps)));
    const int min_staged_warps = std::max(1, env_int("AMX_MIN_STAGED_WARPS", 8));
    bool staged = !batched && !lasso_fast && env_int("AMX_NO_TMA", 0) == 0 && 128 + (size_t)pl->slab_bytes + ws_bytes * min_staged_warps <= budget;
    size_t fixed = 128 + (staged ? pl->slab_bytes : 0);
    int nwarps = (int)std::min<size_t>(want_warps, (budget - fixed) / ws_bytes);
    if (nwarps < 1) return fail(AMX_E_INVALID, "per-warp workspace (%zu B) does not fit in shared memory (m=%d, n=%d)", ws_bytes, pl->m, pl->n);
    p.slab_bytes = staged ? pl->slab_bytes : 0;
    p.slab_smem_off = 128;
    p.ws_smem_off = (unsigned)fixed;
    p.nwarps = nwarps;
    const size_t smem = fixed + ws_bytes * nwarps;
    int ctas_per_sm = std::max(1, (int)std::min<size_t>(budget / smem, (size_t)std::max(1, 16 / nwarps)));
    if (staged) ctas_per_sm = std::max(1, std::min(ctas_per_sm, env_int("AMX_CTAS_PER_SM", 1)));
    int grid = (int)std::max<long long>(1, std::min<long long>(n_tiles_bound, (long long)pl->sm_count * ctas_per_sm));

    if (p.batched == 4) {
        // no per-warp scratch
    } else if (p.batched == 3) {
        CK(wk.scratch.reserve((size_t)grid * 32 * BV * p.NA * sizeof(double)));
        p.scratch = (double *)wk.scratch.p;
    } else if (p.batched) {
        CK(wk.scratch.reserve((size_t)grid * 32 * 2 * BV * p.NA * sizeof(double)));
        p.scratch = (double *)wk.scratch.p;
        CK(wk.xiso.reserve((size_t)n_vox * 2 * sizeof(double)));
        CK(wk.supmask.reserve((size_t)n_vox * 8 * sizeof(unsigned)));
        p.xiso = (double *)wk.xiso.p;
        p.supmask = (unsigned *)wk.supmask.p;
        // stage 3, one voxel per thread: voxels it hands back (passive set > 6 atoms, support > 32) as one-voxel tiles + 2 counters
        CK(wk.redo.reserve((size_t)n_vox * sizeof(int4) + 16));
        p.redo_count = (int *)wk.redo.p;
        p.redo_tiles = (int4 *)((char *)wk.redo.p + 16);
        CK(cudaMemsetAsync(p.redo_count, 0, 16, st));
        // stage 1's c1 = A^T y kept per voxel for stage 3 when it fits the budget (AMX_C1_STORE_MB, default 4096 MB)
        const size_t c1_bytes = (size_t)n_vox * p.NA * sizeof(double);
        if (c1_bytes <= (size_t)std::max(0, env_int("AMX_C1_STORE_MB", 4096)) * 1048576) {
            CK(wk.c1_all.reserve(c1_bytes));
            p.c1_all = (double *)wk.c1_all.p;
        }
        p.tpv3 = env_int("AMX_TPV3", 1) && p.c1_all && p.n <= 255;
        p.tpv1 = env_int("AMX_TPV1", 0) && p.c1_all && p.cap_stage[0] <= 16 && p.n <= 255;
        p.ovf_cap = 4 * n_vox;
        CK(wk.ovf_list.reserve((size_t)p.ovf_cap * sizeof(int)));
        p.ovf_list = (int *)wk.ovf_list.p;
        // per-chunk queue counters (the consumers add them to the call totals in status[6], status[7])
        CK(cudaMemsetAsync(status + 2, 0, sizeof(long long), st));
        CK(cudaMemsetAsync(status + 4, 0, sizeof(long long), st));
        p.exact_tol = 0.0;
        if (p.batched == 2) {
            const char *tol = getenv("AMX_EXACT_TOL");
            p.exact_tol = (tol && *tol) ? atof(tol) : 1e-6;
            if (a->flags & AMX_FLAG_EXACT) p.exact_tol = 1e300;  // bit-reproducible mode: every voxel is re-fitted by the A-space path
            p.exact_cap = n_vox;
            CK(wk.exact_list.reserve((size_t)n_vox * sizeof(int)));
            p.exact_list = (int *)wk.exact_list.p;
        }
    }
    if (pl->model == AMX_MODEL_NODDI && !p.batched) {  // per-voxel kernel (dictionaries of > 160 atoms): overflow queue of the slow path
        p.ovf_cap = n_vox;
        CK(wk.ovf_list.reserve((size_t)p.ovf_cap * sizeof(int)));
        p.ovf_list = (int *)wk.ovf_list.p;
        CK(cudaMemsetAsync(status + 2, 0, sizeof(long long), st));
    }
    int rc;
    switch (pl->model) {
    case AMX_MODEL_NODDI:
        rc = dispatch_npl<MODEL_NODDI, float>(pl->npl, p, grid, nwarps * 32, smem, st);
        break;
    case AMX_MODEL_FREEWATER: rc = dispatch_npl<MODEL_FREEWATER, float>(pl->npl, p, grid, nwarps * 32, smem, st); break;
    case AMX_MODEL_CZB: rc = dispatch_npl<MODEL_CZB, float>(pl->npl, p, grid, nwarps * 32, smem, st); break;
    default: rc = dispatch_npl<MODEL_SANDI, double>(pl->npl, p, grid, nwarps * 32, smem, st); break;
    }
    if (rc) return rc;
    *launches += (p.batched == 2) ? 3 + p.tpv3 + p.tpv1 : 1;
    if (p.batched == 2 && p.exact_tol > 0.0) {
        // exact-fit voxels queued by stage 1 are re-fitted from scratch by the reference's own algorithm (A-space Lawson-Hanson
        // with Householder QR, amx_exact.cuh); unconditional launch, returns at once when the queue is empty
        const int m_pad = (pl->m + 1) & ~1, dc_pad = (pl->dc + 1) & ~1;
        const size_t smem_x = (size_t)(ws_doubles_for(p.NA, m_pad, dc_pad, 0, LC) + exact_extra_doubles(pl->m, p.NA)) * sizeof(double);
        if (smem_x <= (size_t)pl->max_smem) {
            const int grid_x = (int)std::max<long long>(1, std::min<long long>(n_vox, (long long)pl->sm_count * 6));
            CK(wk.exact_a.reserve((size_t)grid_x * pl->m * pl->n * sizeof(double)));
            switch (pl->npl) {
            case 1: rc = launch_noddi_exact<1>(p, grid_x, smem_x, status, (double *)wk.exact_a.p, st); break;
            case 2: rc = launch_noddi_exact<2>(p, grid_x, smem_x, status, (double *)wk.exact_a.p, st); break;
            case 3: case 4: rc = launch_noddi_exact<4>(p, grid_x, smem_x, status, (double *)wk.exact_a.p, st); break;
            default: rc = launch_noddi_exact<5>(p, grid_x, smem_x, status, (double *)wk.exact_a.p, st); break;
            }
            if (rc) return rc;
            *launches += 1;
        }
    }
    if (p.batched == 2 || (pl->model == AMX_MODEL_NODDI && !p.batched)) {
        // voxels whose active set outgrew a warp (possible with a small lambda1) are re-fitted by the scalar slow path;
        // the launch is unconditional and returns at once when the queue is empty
        const int cap = std::min(pl->m, pl->n) + 2;
        const size_t wsb = slow_ws_bytes(cap, p.NA);
        const int slow_threads = pl->sm_count * std::max(1, env_int("AMX_SLOW_WARPS_PER_SM", 2)) * 32;  // one voxel per thread, ~100-200 KB of workspace each
        CK(wk.slow_ws.reserve(wsb * slow_threads));
        k_slow_noddi<float><<<slow_threads / 32, 32, 0, st>>>(p, p.ovf_list, status, (unsigned char *)wk.slow_ws.p, wsb, cap);
        CK(cudaGetLastError());
        *launches += 1;
    }
    if (record_events) CK(cudaEventRecord(pl->ev[2], st));
    pl->last_cnt[1] = n_tiles_bound;  // upper bound; the exact count stays on the device
    pl->last_cnt[3] = (int64_t)smem;
    pl->last_cnt[4] = nwarps;
    pl->last_cnt[5] = staged ? 1 : 0;
    pl->last_cnt[7] = grid;
    return AMX_OK;
}

}  // namespace

extern "C" {

int amx_fit(amx_plan *pl, const amx_fit_args *a, int64_t *err_voxel)
{
    if (!pl || !a) return fail(AMX_E_INVALID, "plan/args is NULL");
    if (a->n_vox < 0 || a->n_vox >= ((long long)1 << 31)) return fail(AMX_E_INVALID, "n_vox=%lld out of range", (long long)a->n_vox);
    if (!a->y || !a->estimates) return fail(AMX_E_INVALID, "y/estimates is NULL");
    if ((a->flags & AMX_FLAG_RMSE) && !a->rmse) return fail(AMX_E_INVALID, "AMX_FLAG_RMSE without rmse buffer");
    if ((a->flags & AMX_FLAG_NRMSE) && !a->nrmse) return fail(AMX_E_INVALID, "AMX_FLAG_NRMSE without nrmse buffer");
    const bool has_extra = (a->flags & AMX_FLAG_EXTRA) && (pl->model == AMX_MODEL_NODDI || pl->model == AMX_MODEL_FREEWATER);
    if (has_extra && !a->extra) return fail(AMX_E_INVALID, "AMX_FLAG_EXTRA without extra buffer");
    if (a->y_dtype != AMX_F32 && a->y_dtype != AMX_F64) return fail(AMX_E_INVALID, "bad y_dtype %d", a->y_dtype);
    if (pl->model != AMX_MODEL_SANDI && !a->dirs) return fail(AMX_E_INVALID, "dirs is NULL");
    if (a->space != AMX_SPACE_DEVICE && a->space != AMX_SPACE_HOST) return fail(AMX_E_INVALID, "bad space %d", a->space);
    CK(cudaSetDevice(pl->device));
    pl->timing_valid = false;
    memset(pl->last_cnt, 0, sizeof pl->last_cnt);
    if (a->n_vox == 0) return AMX_OK;
    int launches = 0;
    amx_fit_args d = *a;
    if (!has_extra) d.flags &= ~AMX_FLAG_EXTRA;
    const bool host = a->space == AMX_SPACE_HOST;
    const size_t n = (size_t)a->n_vox, m = (size_t)pl->m, nm = (size_t)pl->n_maps;
    const size_t ysz = a->y_dtype == AMX_F64 ? 8 : 4;
    const size_t extra_w = has_extra ? (pl->model == AMX_MODEL_NODDI ? 2 : m) : 0;
    // The caller's stream: everything is ordered after / joined back into it.  Device-resident inputs are produced by the caller's
    // own stream-ordered work, so a NULL stream means the LEGACY DEFAULT stream (what a cudaStream_t of 0 denotes everywhere else,
    // and what PyTorch hands out as its default stream), never the plan's private non-blocking stream: running there let a fit
    // start while the default stream was still writing y / dirs (seen as garbage maps of the first fit after a long-running
    // producer, and as NaNs in the sharded bench).  Host buffers: the plan's own stream.
    cudaStream_t st = pl->stream;
    if (!host) st = a->stream ? (cudaStream_t)a->stream : cudaStreamLegacy;

    // The volume is cut into voxel chunks that alternate between two compute streams (each with its own workspace):
    // the ragged end of one chunk's stage kernels is filled by the next chunk's kernels, and for host buffers the
    // H2D / D2H copies of neighbouring chunks hide behind the fit (3-stream pipeline over two staging sets).
    // (measured: for device-resident data one chunk on one stream is fastest -- per-chunk binning and smaller direction
    //  bins cost more than the ragged kernel ends -- so chunking defaults on for host buffers only, and the second
    //  compute stream stays an option, AMX_COMPUTE_STREAMS=2)
    //  Host buffers: the copy engine moves a voxel ~5x faster than the fit consumes it, so the chunks GROW geometrically
    //  (n/16, n/4, rest): only the small first chunk's upload and the last chunk's download are exposed, and the stage
    //  kernels see three ragged ends instead of one per fixed-size chunk.  AMX_HOST_CHUNK=<voxels> restores equal chunks.
    // Pageable or float64 host signals go through pinned staging filled by host threads (stage_signal): equal chunks, so that
    // the conversion of chunk i + 1 overlaps the upload of chunk i and the fit of chunk i - 1.
    const bool staged_host = host && env_int("AMX_HOST_STAGE", 1) && (a->y_dtype == AMX_F64 || !is_pinned(a->y));
    std::vector<long long> bounds{0};
    {
        long long fixed = env_int(host ? "AMX_HOST_CHUNK" : "AMX_DEVICE_CHUNK", host ? 0 : 0x7fffffff);
        if (staged_host && fixed <= 0 && (long long)n >= 131072)
            fixed = std::max<long long>(32768, ((long long)n / std::max(2, env_int("AMX_STAGE_CHUNKS", 8)) + 7) & ~7LL);
        if (fixed > 0) {
            const long long c = std::max<long long>(8192, fixed);
            if ((long long)n <= c + c / 2) bounds.push_back((long long)n);
            else for (long long o = c; ; o += c) { bounds.push_back(std::min<long long>(o, (long long)n)); if (o >= (long long)n) break; }
        } else if ((long long)n < 131072) {
            bounds.push_back((long long)n);
        } else {
            const long long first = std::max(2, env_int("AMX_HOST_FIRST", 16)), growth = std::max(1, env_int("AMX_HOST_GROWTH", 4));
            long long c = (((long long)n / first) + 7) & ~7LL, o = 0;
            for (;;) {
                o += c;
                if (o + c / 2 >= (long long)n) break;  // the last chunk takes the remainder
                bounds.push_back(o);
                c *= growth;
            }
            bounds.push_back((long long)n);
        }
    }
    const long long n_chunks = (long long)bounds.size() - 1;
    long long chunk = 0;
    for (long long i = 0; i < n_chunks; ++i) chunk = std::max(chunk, bounds[i + 1] - bounds[i]);
    const int nset = n_chunks > 1 ? 2 : 1;
    const int ncs = (nset > 1 && env_int("AMX_COMPUTE_STREAMS", 1) > 1) ? 2 : 1;
    for (int b = 0; b < ncs; ++b) CK(pl->work[b].status.reserve(64));
    if (host) {
        for (int b = 0; b < nset; ++b) {
            amx_plan::Stage &sg = pl->stg[b];
            CK(sg.y.reserve((size_t)chunk * m * ysz));
            CK(sg.est.reserve((size_t)chunk * nm * sizeof(double)));
            if (a->dirs) CK(sg.dirs.reserve((size_t)chunk * 3 * sizeof(double)));
            if (d.flags & AMX_FLAG_RMSE) CK(sg.rmse.reserve((size_t)chunk * sizeof(double)));
            if (d.flags & AMX_FLAG_NRMSE) CK(sg.nrmse.reserve((size_t)chunk * sizeof(double)));
            if (has_extra) CK(sg.extra.reserve((size_t)chunk * extra_w * sizeof(double)));
            if (a->support_out) CK(sg.sup.reserve((size_t)chunk * sizeof(int)));
            if (a->coeff_out) CK(sg.coef.reserve((size_t)chunk * pl->n * sizeof(double)));
            if (a->lut_out) CK(sg.lut.reserve((size_t)chunk * sizeof(int)));
        }
    }
    const int n_host_threads = host_threads();
    if (staged_host) {
        for (int b = 0; b < nset; ++b) {
            const size_t want = (size_t)chunk * m * ysz;
            if (pl->hpin_cap[b] < want) {
                if (pl->hpin[b]) cudaFreeHost(pl->hpin[b]);
                pl->hpin[b] = nullptr; pl->hpin_cap[b] = 0;
                CK(cudaHostAlloc(&pl->hpin[b], want + want / 8, cudaHostAllocDefault));
                pl->hpin_cap[b] = want + want / 8;
            }
        }
    }
    // A device -> host copy into PAGEABLE memory blocks the calling thread until the data are there, i.e. until that chunk's fit
    // has finished -- which would serialise the host-side staging of the next chunk behind it.  Pageable outputs therefore land in
    // pinned buffers and are moved to the caller's arrays by host threads after the last chunk.
    double *o_est = a->estimates, *o_dirs = a->dirs, *o_rmse = a->rmse, *o_nrmse = a->nrmse;
    bool landed[4] = {false, false, false, false};
    if (host && n_chunks > 1 && env_int("AMX_HOST_STAGE", 1)) {
        double **dst[4] = {&o_est, &o_dirs, &o_rmse, &o_nrmse};
        const size_t bytes[4] = {n * nm * sizeof(double), a->dirs ? n * 3 * sizeof(double) : 0, (d.flags & AMX_FLAG_RMSE) ? n * sizeof(double) : 0,
                                 (d.flags & AMX_FLAG_NRMSE) ? n * sizeof(double) : 0};
        for (int k = 0; k < 4; ++k) {
            if (!bytes[k] || !*dst[k] || is_pinned(*dst[k])) continue;
            if (pl->hout_cap[k] < bytes[k]) {
                if (pl->hout[k]) cudaFreeHost(pl->hout[k]);
                pl->hout[k] = nullptr; pl->hout_cap[k] = 0;
                CK(cudaHostAlloc(&pl->hout[k], bytes[k], cudaHostAllocDefault));
                pl->hout_cap[k] = bytes[k];
            }
            *dst[k] = (double *)pl->hout[k];
            landed[k] = true;
        }
    }
    cudaStream_t cs[2] = {st, ncs > 1 ? pl->cs[1] : st};
    cudaStream_t s_in = (host && nset > 1) ? pl->s_in : st, s_out = (host && nset > 1) ? pl->s_out : st;
    CK(cudaEventRecord(pl->ev[3], st));
    {
        long long init[8] = {0, (long long)1 << 62, 0, 0, 0, 0, 0, 0};
        for (int b = 0; b < ncs; ++b) CK(cudaMemcpyAsync(pl->work[b].status.p, init, sizeof init, cudaMemcpyHostToDevice, st));
    }
    {
        int rc = prepare_tables(pl, &d, st, &launches);
        if (rc) return rc;
    }
    if (nset > 1) {  // fork
        CK(cudaEventRecord(pl->ev_fork, st));
        if (ncs > 1) CK(cudaStreamWaitEvent(cs[1], pl->ev_fork, 0));
        if (host) { CK(cudaStreamWaitEvent(s_in, pl->ev_fork, 0)); CK(cudaStreamWaitEvent(s_out, pl->ev_fork, 0)); }
    }
    for (long long i = 0; i < n_chunks; ++i) {
        const int b = (int)(i % nset), wb = (int)(i % ncs);
        const size_t off = (size_t)bounds[i], cnt = (size_t)(bounds[i + 1] - bounds[i]);
        amx_fit_args c = d;
        c.n_vox = (int64_t)cnt;
        if (host) {
            amx_plan::Stage &sg = pl->stg[b];
            if (i >= 2) {
                CK(cudaStreamWaitEvent(s_in, pl->ev_comp[b], 0));  // inputs of this set consumed by chunk i-2 ...
                CK(cudaStreamWaitEvent(s_in, pl->ev_out[b], 0));   // ... and its flipped dirs (same buffer) read back
            }
            if (staged_host) {
                if (i >= 2) CK(cudaEventSynchronize(pl->ev_in[b]));  // the upload of chunk i - 2 has left this pinned set
                c.y_dtype = stage_signal((const char *)a->y + off * m * ysz, a->y_dtype, pl->hpin[b], cnt * m, n_host_threads);
                CK(cudaMemcpyAsync(sg.y.p, pl->hpin[b], cnt * m * (c.y_dtype == AMX_F64 ? 8 : 4), cudaMemcpyHostToDevice, s_in));
            } else {
                CK(cudaMemcpyAsync(sg.y.p, (const char *)a->y + off * m * ysz, cnt * m * ysz, cudaMemcpyHostToDevice, s_in));
            }
            if (a->dirs) CK(cudaMemcpyAsync(sg.dirs.p, a->dirs + off * 3, cnt * 3 * sizeof(double), cudaMemcpyHostToDevice, s_in));
            c.y = sg.y.p; c.estimates = (double *)sg.est.p;
            c.dirs = a->dirs ? (double *)sg.dirs.p : nullptr;
            c.rmse = (d.flags & AMX_FLAG_RMSE) ? (double *)sg.rmse.p : nullptr;
            c.nrmse = (d.flags & AMX_FLAG_NRMSE) ? (double *)sg.nrmse.p : nullptr;
            c.extra = has_extra ? (double *)sg.extra.p : nullptr;
            c.support_out = a->support_out ? (int *)sg.sup.p : nullptr;
            c.coeff_out = a->coeff_out ? (double *)sg.coef.p : nullptr;
            c.lut_out = a->lut_out ? (int *)sg.lut.p : nullptr;
            if (nset > 1) {
                CK(cudaEventRecord(pl->ev_in[b], s_in));
                CK(cudaStreamWaitEvent(cs[wb], pl->ev_in[b], 0));
                if (i >= 2) CK(cudaStreamWaitEvent(cs[wb], pl->ev_out[b], 0));  // outputs of this set drained by chunk i-2
            }
        } else {
            c.y = (const char *)a->y + off * m * ysz;
            c.estimates = a->estimates + off * nm;
            c.dirs = a->dirs ? a->dirs + off * 3 : nullptr;
            c.rmse = a->rmse ? a->rmse + off : nullptr;
            c.nrmse = a->nrmse ? a->nrmse + off : nullptr;
            c.extra = (has_extra && a->extra) ? a->extra + off * extra_w : nullptr;
            c.support_out = a->support_out ? a->support_out + off : nullptr;
            c.coeff_out = a->coeff_out ? a->coeff_out + off * pl->n : nullptr;
            c.lut_out = a->lut_out ? a->lut_out + off : nullptr;
        }
        const bool early_dirs = host && nset > 1 && a->dirs;
        int rc = fit_device(pl, pl->work[wb], &c, cs[wb], &launches, (long long)off, i == 0, early_dirs ? pl->ev_lut[b] : nullptr);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        if (host) {
            // the reference flips DIRs in place (amico/lut.pyx:335-338): hand the flipped directions back -- as soon as the LUT
            // kernel has flipped them, i.e. behind the fit instead of after it
            if (early_dirs) {
                CK(cudaStreamWaitEvent(s_out, pl->ev_lut[b], 0));
                CK(cudaMemcpyAsync(o_dirs + off * 3, c.dirs, cnt * 3 * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            }
            if (nset > 1) {
                CK(cudaEventRecord(pl->ev_comp[b], cs[wb]));
                CK(cudaStreamWaitEvent(s_out, pl->ev_comp[b], 0));
            }
            if (a->dirs && !early_dirs) CK(cudaMemcpyAsync(o_dirs + off * 3, c.dirs, cnt * 3 * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            CK(cudaMemcpyAsync(o_est + off * nm, c.estimates, cnt * nm * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            if (c.rmse) CK(cudaMemcpyAsync(o_rmse + off, c.rmse, cnt * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            if (c.nrmse) CK(cudaMemcpyAsync(o_nrmse + off, c.nrmse, cnt * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            if (c.extra) CK(cudaMemcpyAsync(a->extra + off * extra_w, c.extra, cnt * extra_w * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            if (c.support_out) CK(cudaMemcpyAsync(a->support_out + off, c.support_out, cnt * sizeof(int), cudaMemcpyDeviceToHost, s_out));
            if (c.coeff_out) CK(cudaMemcpyAsync(a->coeff_out + off * pl->n, c.coeff_out, cnt * pl->n * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            if (c.lut_out) CK(cudaMemcpyAsync(a->lut_out + off, c.lut_out, cnt * sizeof(int), cudaMemcpyDeviceToHost, s_out));
            if (nset > 1) CK(cudaEventRecord(pl->ev_out[b], s_out));
        }
    }
    if (nset > 1) {  // join everything back into the caller's stream
        if (ncs > 1) {
            CK(cudaEventRecord(pl->ev_join[1], cs[1]));
            CK(cudaStreamWaitEvent(st, pl->ev_join[1], 0));
        }
        if (host) {
            CK(cudaEventRecord(pl->ev_join[0], s_out));
            CK(cudaStreamWaitEvent(st, pl->ev_join[0], 0));
        }
    }
    CK(cudaEventRecord(pl->ev[4], st));
    long long h_status[2][8] = {{0, (long long)1 << 62, 0, 0, 0, 0, 0, 0}, {0, (long long)1 << 62, 0, 0, 0, 0, 0, 0}};
    for (int b = 0; b < ncs; ++b) CK(cudaMemcpyAsync(h_status[b], pl->work[b].status.p, sizeof h_status[b], cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {   // pageable outputs: pinned landing buffers -> the caller's arrays
        double *user[4] = {a->estimates, a->dirs, a->rmse, a->nrmse};
        const size_t bytes[4] = {n * nm * sizeof(double), n * 3 * sizeof(double), n * sizeof(double), n * sizeof(double)};
        for (int k = 0; k < 4; ++k) {
            if (!landed[k]) continue;
            constexpr size_t BLK = 1 << 20;
            const char *src = (const char *)pl->hout[k];
            char *dst = (char *)user[k];
            const size_t total = bytes[k];
            parallel_blocks((total + BLK - 1) / BLK, n_host_threads, [&](size_t b) { memcpy(dst + b * BLK, src + b * BLK, std::min(BLK, total - b * BLK)); });
        }
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pl->ev[0], pl->ev[1]) == cudaSuccess) pl->last_ms[0] = ms;
    if (cudaEventElapsedTime(&ms, pl->ev[3], pl->ev[4]) == cudaSuccess) pl->last_ms[2] = ms;
    // chunks overlap on two streams, so the fit kernels are timed as one span: whole call minus the first chunk's binning
    // (single chunk: exactly the event pair around the fit kernels)
    if (nset == 1) { if (cudaEventElapsedTime(&ms, pl->ev[1], pl->ev[2]) == cudaSuccess) pl->last_ms[1] = ms; }
    else pl->last_ms[1] = pl->last_ms[2] - pl->last_ms[0];
    pl->timing_valid = true;
    pl->last_cnt[0] = launches;
    const long long bad = h_status[0][0] | h_status[1][0], bad_vox = std::min(h_status[0][1], h_status[1][1]);
    pl->last_cnt[2] = h_status[0][3] + h_status[1][3];
    pl->last_cnt[6] = h_status[0][6] + h_status[1][6] + ((pl->model == AMX_MODEL_NODDI) ? 0 : h_status[0][2] + h_status[1][2]);
    pl->last_cnt[8] = h_status[0][7] + h_status[1][7];
    if (bad) {
        if (err_voxel) *err_voxel = bad_vox;
        return fail(AMX_E_LUT_RANGE, "\"amico.lut.dir_to_lut_idx\" index out of bounds (voxel %lld)", bad_vox);
    }
    if (pl->last_cnt[2])
        return fail(AMX_E_CAPACITY, "%lld voxel(s) outgrew the %d-atom active-set workspace of a kernel without slow path",
                    (long long)pl->last_cnt[2], LC);
    return AMX_OK;
}

int amx_lut_indices(amx_plan *pl, int space, double *dirs, int64_t n, int32_t *idx)
{
    if (!pl || !dirs || !idx || n < 0) return fail(AMX_E_INVALID, "bad argument");
    if (pl->model == AMX_MODEL_SANDI) return fail(AMX_E_INVALID, "SANDI has no direction LUT");
    if (n == 0) return AMX_OK;
    CK(cudaSetDevice(pl->device));
    cudaStream_t st = pl->stream;
    double *d_dirs = dirs;
    int *d_idx = idx;
    if (space == AMX_SPACE_HOST) {
        CK(pl->stg[0].dirs.reserve((size_t)n * 3 * sizeof(double)));
        CK(pl->work[0].lut.reserve((size_t)n * sizeof(int)));
        d_dirs = (double *)pl->stg[0].dirs.p; d_idx = (int *)pl->work[0].lut.p;
        CK(cudaMemcpyAsync(d_dirs, dirs, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CK(pl->work[0].status.reserve(64));
    long long init[3] = {0, (long long)1 << 62, 0};
    CK(cudaMemcpyAsync(pl->work[0].status.p, init, sizeof init, cudaMemcpyHostToDevice, st));
    k_lut<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_dirs, n, pl->d_htable, pl->ndirs, d_idx, nullptr, (long long *)pl->work[0].status.p, 0);
    CK(cudaGetLastError());
    if (space == AMX_SPACE_HOST) {
        CK(cudaMemcpyAsync(dirs, d_dirs, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(idx, d_idx, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    long long h_status[2];
    CK(cudaMemcpyAsync(h_status, pl->work[0].status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_status[0]) return fail(AMX_E_LUT_RANGE, "\"amico.lut.dir_to_lut_idx\" index out of bounds (voxel %lld)", h_status[1]);
    return AMX_OK;
}

int amx_plan_last_timing(amx_plan *pl, double *out_ms, int n)
{
    if (!pl || !out_ms) return fail(AMX_E_INVALID, "bad argument");
    if (!pl->timing_valid) return fail(AMX_E_INVALID, "no completed fit on this plan");
    for (int i = 0; i < n && i < 8; ++i) out_ms[i] = pl->last_ms[i];
    return AMX_OK;
}

int amx_plan_last_counters(amx_plan *pl, int64_t *out, int n)
{
    if (!pl || !out) return fail(AMX_E_INVALID, "bad argument");
    for (int i = 0; i < n && i < 16; ++i) out[i] = pl->last_cnt[i];
    return AMX_OK;
}

}  // extern "C"
