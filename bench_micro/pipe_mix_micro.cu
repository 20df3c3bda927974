// pipe_mix_micro.cu -- design microbenchmark: do DMMA.8x8x4 (FP64 tensor MMA) and DFMA share ONE pipe on sm_100a?
// Per iteration a warp issues ND independent DMMA chains and NF independent DFMA chains.  Pipe time per
// sub-partition at the measured peaks: 16 cycles per DMMA (256 FMA at 16 FMA/clk), 2 cycles per DFMA.
// time(mix) ~ time(DMMA only) + time(DFMA only): one shared pipe;  ~ max of the two: separate pipes.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ND, int NF>
__global__ void mix_kernel(double *out, int iters, double a, double b) {
    double c[ND > 0 ? ND : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < ND; ++i) { c[i][0] = 0.0; c[i][1] = threadIdx.x; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (ND > NF ? ND : NF); ++i) {
            if (i < ND) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < (ND > 0 ? NF / ND : NF); ++j) {
                const int k = (ND > 0) ? i * (NF / ND) + j : i;
                if (k < NF && (ND > 0 ? i < ND : j == 0)) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[k]) : "d"(a), "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ND; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K> double run(K kern, int blocks, int threads, double *out, int iters) {
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CHECK(cudaEventRecord(e0)); kern<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1)); CHECK(cudaGetLastError());
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int khz = 0; CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    double *out; CHECK(cudaMalloc(&out, 148 * 16 * 128 * 8));
    const int iters = 20000;
    for (int warps = 4; warps <= 8; warps *= 2) {
        const int threads = 128, blocks = p.multiProcessorCount * warps / 4;
        const double td = run(mix_kernel<4, 0>, blocks, threads, out, iters);
        const double tf = run(mix_kernel<0, 32>, blocks, threads, out, iters);
        const double tm = run(mix_kernel<4, 32>, blocks, threads, out, iters);
        const double wps = warps / 4.0;      // warps per sub-partition
        const double cyc = khz * 1e3 / iters / wps;
        printf("%d warps/SM: per warp-iteration cycles (nominal clock)  4 DMMA: %.1f | 32 DFMA: %.1f | 4 DMMA + 32 DFMA: %.1f   (sum %.1f, max %.1f)\n",
               warps, td * cyc, tf * cyc, tm * cyc, (td + tf) * cyc, (td > tf ? td : tf) * cyc);
    }
    return 0;
}
