// tps_micro.cu -- design microbenchmark (not product code): thread-per-scenario FP64 building
// blocks at 1 and 2 warps per SM sub-partition.  Decides where the shared constant matrices of
// the fast-path step kernel should live (constant bank vs shared-memory broadcast vs global
// uniform loads) and what an in-shared-memory thread-per-scenario Cholesky costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tps_micro tps_micro.cu && ./tps_micro
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int R1 = 28, C1 = 22;   // e = Px x
constexpr int R2 = 20, C2 = 28;   // f = E^T t
constexpr int R3 = 20, C3 = 42;   // y' = [Myx Myu][x; v]
constexpr int NCONST = R1 * C1 + R2 * C2 + R3 * C3;   // 2016
constexpr int NV = 20, NL = NV * (NV + 1) / 2;

__constant__ double cM[NCONST + 3 * NL];

template <int MODE>   // 0: constant bank, 1: shared broadcast, 2: global uniform
__device__ __forceinline__ double getc(const double *sm, const double *gm, int idx) {
    if (MODE == 0) return cM[idx];
    if (MODE == 1) return sm[idx];
    return __ldg(gm + idx);
}

template <int MODE>
__global__ void matvec_kernel(const double *__restrict__ gm, double *out, int iters) {
    extern __shared__ double smem[];
    if (MODE == 1) {
        for (int i = threadIdx.x; i < NCONST; i += blockDim.x) smem[i] = gm[i];
        __syncthreads();
    }
    double x[C1], v[NV];
#pragma unroll
    for (int j = 0; j < C1; ++j) x[j] = 1.0 + 1e-3 * j + 1e-6 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = 0.5 + 1e-3 * j;
    for (int it = 0; it < iters; ++it) {
        double e[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < C1; ++j) a = fma(getc<MODE>(smem, gm, r * C1 + j), x[j], a);
            e[r] = a;
        }
        double f[R2];
#pragma unroll
        for (int c = 0; c < R2; ++c) {
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < C2; ++r) a = fma(getc<MODE>(smem, gm, R1 * C1 + c * C2 + r), e[r], a);
            f[c] = a;
        }
#pragma unroll
        for (int r = 0; r < R3; ++r) {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < C1; ++j) a = fma(getc<MODE>(smem, gm, R1 * C1 + R2 * C2 + r * C3 + j), x[j], a);
#pragma unroll
            for (int j = 0; j < NV; ++j) a = fma(getc<MODE>(smem, gm, R1 * C1 + R2 * C2 + r * C3 + C1 + j), f[j], a);
            v[r] = a * 1e-3;
        }
#pragma unroll
        for (int j = 0; j < C1; ++j) x[j] = x[j] * 0.999 + 1e-6 * v[j % NV];
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < C1; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// thread-per-scenario Cholesky, L in shared memory as [idx][tid]; H built on the fly from three
// constant packed matrices (Ht + w0 H0 + w1 H1 + diag); forward solve fused; backward solve from smem.
template <int MODE, int CB>   // CB: columns per block (1 or 2)
__global__ void chol_kernel(const double *__restrict__ gm, double *out, int iters) {
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    double *Ls = smem;                   // [NL][T]
    double *sc = smem + NL * T;          // constants for MODE 1
    if (MODE == 1) {
        for (int i = tid; i < 3 * NL; i += T) sc[i] = gm[NCONST + i];
        __syncthreads();
    }
    const double *gc = gm + NCONST;
    double w0 = 0.1 + 1e-4 * tid, w1 = 0.4, wd = 9.0;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        double y[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) y[i] = 1.0 + 0.01 * i + acc * 1e-9;
        double dinv[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double a[NV];
#pragma unroll
            for (int i = k; i < NV; ++i) {
                const int p = i * (i + 1) / 2 + k;
                double h;
                if (MODE == 0) h = fma(w1, cM[NCONST + 2 * NL + p], fma(w0, cM[NCONST + NL + p], cM[NCONST + p]));
                else if (MODE == 1) h = fma(w1, sc[2 * NL + p], fma(w0, sc[NL + p], sc[p]));
                else h = fma(w1, __ldg(gc + 2 * NL + p), fma(w0, __ldg(gc + NL + p), __ldg(gc + p)));
                a[i] = h + (i == k ? wd : 0.0);
            }
            double yk = y[k];
#pragma unroll
            for (int j = 0; j < k; ++j) {
                const double lkj = Ls[(k * (k + 1) / 2 + j) * T + tid];
                a[k] = fma(-lkj, lkj, a[k]);
#pragma unroll
                for (int i = k + 1; i < NV; ++i) a[i] = fma(-Ls[(i * (i + 1) / 2 + j) * T + tid], lkj, a[i]);
                yk = fma(-lkj, y[j], yk);
            }
            const double inv = rsqrt(a[k]);
            dinv[k] = inv;
            y[k] = yk * inv;
#pragma unroll
            for (int i = k + 1; i < NV; ++i) Ls[(i * (i + 1) / 2 + k) * T + tid] = a[i] * inv;
        }
        // backward solve L^T z = y
#pragma unroll
        for (int j = NV - 1; j >= 0; --j) {
            const double zj = y[j] * dinv[j];
            y[j] = zj;
#pragma unroll
            for (int i = 0; i < j; ++i) y[i] = fma(-Ls[(j * (j + 1) / 2 + i) * T + tid], zj, y[i]);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) acc += y[i];
    }
    out[blockIdx.x * T + tid] = acc;
}

template <typename K>
double time_kernel(K launch, int reps = 3) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
        CHECK(cudaEventRecord(e0));
        launch();
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    std::vector<double> h(NCONST + 3 * NL);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 1e-2 * ((double)((i * 2654435761u) % 1000) / 1000.0 - 0.5);
    // make the packed matrices diagonally dominant enough
    double *gm, *out;
    CHECK(cudaMalloc(&gm, h.size() * 8)); CHECK(cudaMemcpy(gm, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpyToSymbol(cM, h.data(), h.size() * 8));
    CHECK(cudaMalloc(&out, sms * 8 * 256 * 8));
    const int iters = 2000;
    printf("SMs %d\n", sms);
    const char *names[3] = {"constant-bank", "smem-broadcast", "global-uniform"};
    for (int wps = 1; wps <= 2; ++wps) {   // warps per SM sub-partition
        const int threads = 128;
        const int blocks_per_sm = wps;
        const size_t smem = (wps == 1) ? 200 * 1024 : 100 * 1024;   // forces the residency
        auto run = [&](auto kern, const char *nm) {
            CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            double s = time_kernel([&] { kern<<<sms * blocks_per_sm, threads, smem>>>(gm, out, iters); });
            double macs = (double)sms * blocks_per_sm * threads * iters * (double)NCONST;
            printf("matvec %-15s %d warp/SMSP: %.3f ms  %.2f TFLOP/s  (%.1f cycles per 32-scenario iteration @1.965GHz)\n", nm, wps,
                   s * 1e3, 2 * macs / s / 1e12, s * 1.965e9 / iters);
        };
        run(matvec_kernel<0>, names[0]);
        run(matvec_kernel<1>, names[1]);
        run(matvec_kernel<2>, names[2]);
    }
    {
        const int threads = 128;
        const size_t smem = (size_t)NL * threads * 8 + 3 * NL * 8 + 64;
        auto run = [&](auto kern, const char *nm) {
            CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int it2 = 500;
            double s = time_kernel([&] { kern<<<sms, threads, smem>>>(gm, out, it2); });
            double fl = (double)sms * threads * it2 * (NV * NV * NV / 3.0 + 2.0 * NV * NV + 3.0 * 2 * NL);
            printf("chol+solve %-15s 1 warp/SMSP: %.3f ms  %.2f TFLOP/s  (%.0f cycles per 32-scenario factor+solve)\n", nm, s * 1e3,
                   fl / s / 1e12, s * 1.965e9 / it2);
        };
        run(chol_kernel<0, 1>, names[0]);
        run(chol_kernel<1, 1>, names[1]);
        run(chol_kernel<2, 1>, names[2]);
    }
    return 0;
}
