// FP64 throughput microbenchmark ("measure with an FP64 FMA/DMMA microbenchmark"): the
// denominator of the honest roofline of the per-voxel solvers, which are FP64 issue-bound, not HBM-bound.
//
//   kind 0: DFMA   -- 8 independent fma chains per thread (2 flop per lane and instruction)
//   kind 1: DMMA   -- mma.sync.m8n8k4.f64 (the instruction the A^T Y micro-GEMMs use), 4 independent accumulators
//                     per warp (8*8*4*2 = 512 flop per warp and instruction)
//   kind 2: issue  -- independent integer LOP3 / IADD3 chains: warp instructions per second (the issue-slot roofline)
//
// Built by __graft_entry__.build() into tools/libfp64peak.so; called by bench.py through ctypes.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a, double b)
{
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
    const double fa = a + threadIdx.x * 1e-9, fb = b;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_issue(double *out, int iters, int a, int b)
{
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = (x[i] ^ a) + b;  // LOP3 + IADD3: full-rate ALU instructions (IMAD issues at half rate)
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (double)s;
}

}  // namespace

// result: kind 0/1 -> TFLOP/s; kind 2 -> 1e12 warp instructions per second.  Returns 0 on success, the CUDA error otherwise.
extern "C" int fp64_peak(int device, int kind, double *result)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int block = 256, grid = sms * 8, iters = 4096;
    double *out = nullptr;
    if ((e = cudaMalloc(&out, (size_t)grid * block * sizeof(double))) != cudaSuccess) return (int)e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) k_dfma<<<grid, block>>>(out, iters, 0.999999, 1e-7);
        else if (kind == 1) k_dmma<<<grid, block>>>(out, iters, 0.5, 0.25);
        else k_issue<<<grid, block>>>(out, iters, 3, 1);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;  // first launch is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return (int)e;
    const double threads = (double)grid * block, warps = threads / 32.0, n_inst = (double)iters * 8.0;
    if (kind == 0) *result = threads * n_inst * 8.0 * 2.0 / (best * 1e-3) / 1e12;
    else if (kind == 1) *result = warps * n_inst * 4.0 * 512.0 / (best * 1e-3) / 1e12;
    else *result = warps * n_inst * 8.0 * 2.0 / (best * 1e-3) / 1e12;  // two instructions per element and step
    return 0;
}
