// Micro-benchmark: fp64 latency / throughput on one SM (tools only, not part of the library).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void rsq_chain(double* out, long long* cyc, int iters) {
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double y;
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
            x = y;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_chain(double* out, long long* cyc, int iters) {
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 31) & 31);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* cyc) {
    const int iters = 2000;
    chain<ILP><<<1, 32 * warps>>>(out, cyc, iters, 0.999, 0.001);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 8 * ILP * warps;  // warp-level DFMA instructions on the SM
    printf("warps %2d ILP %d: %8.2f clk per dependent step, %6.3f warp-DFMA/clk/SM\n", warps, ILP,
           (double)c / (iters * 8), ops / c);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1 << 12);
    for (int warps : {1, 4, 8, 16, 32}) {
        run<1>(warps, out, cyc);
        run<2>(warps, out, cyc);
        run<4>(warps, out, cyc);
        run<8>(warps, out, cyc);
    }
    long long c;
    rsq_chain<<<1, 32>>>(out, cyc, 2000);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("rsqrt.approx.f64 dependent latency: %.1f clk\n", (double)c / 16000);
    shfl_chain<<<1, 32>>>(out, cyc, 2000);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("64-bit shuffle dependent latency: %.1f clk\n", (double)c / 16000);
    return 0;
}
