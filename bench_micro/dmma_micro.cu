// dmma_micro.cu -- design microbenchmark: FP64 tensor MMA (mma.sync.m8n8k4.f64) throughput on
// sm_100a, with register operands and with the B operand streamed from shared memory.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC, bool SMEM>
__global__ void dmma_kernel(double *out, int iters) {
    __shared__ double sb[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sb[i] = 1e-3 * (i % 17);
    __syncthreads();
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-6, b = 0.5;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            double bb = SMEM ? sb[((it + i) & 63) * 32 + lane] : b;
            dmma(c[i][0], c[i][1], a, bb);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K> double run(K kern, int blocks, int threads, double *out, int iters) {
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CHECK(cudaEventRecord(e0)); kern<<<blocks, threads>>>(out, iters); CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1)); CHECK(cudaGetLastError());
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (r && ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    double *out; CHECK(cudaMalloc(&out, 148 * 8 * 1024 * 8));
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        int threads = 128, blocks = p.multiProcessorCount * warps / 4;
        double s1 = run(dmma_kernel<8, false>, blocks, threads, out, iters);
        double s2 = run(dmma_kernel<8, true>, blocks, threads, out, iters);
        double s3 = run(dmma_kernel<2, false>, blocks, threads, out, iters);
        double fl = 2.0 * 256 * 8 * (double)iters * blocks * (threads / 32);
        printf("%2d warps/SM: reg operands %.2f TFLOP/s | B from smem %.2f TFLOP/s | 2 accumulators (latency) %.2f TFLOP/s\n",
               warps, fl / s1 / 1e12, fl / s2 / 1e12, fl / 4 / s3 / 1e12);
    }
    return 0;
}
