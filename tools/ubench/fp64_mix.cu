// Micro-benchmark: what limits the fp64 pipe of one SM in code that looks like K2 (tools only, not part of the library).
//   A  x = fma(x, a, b)            two operands constant (operand-reuse cache): the textbook peak
//   B  x[i] = fma(x[i+1], x[i+2], x[i])   three distinct, changing register operands
//   C  x[i] = fma(x[i+1], a, x[i])        two distinct register operands
//   D  as A, one 64-bit shared-memory load per 8 DFMA mixed in
//   E  as B, one 64-bit shared-memory load per 8 DFMA mixed in
//   F  as A with DADD + DMUL instead of DFMA
// nvcc -arch=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 12;

template <int MODE>
__global__ void mix(double* out, long long* cyc, int iters, double a, double b) {
    __shared__ double sm[32 * 33];
    double x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = threadIdx.x * 1e-3 + i * 0.01 + 1.0;
    for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    const double* sp = sm + (threadIdx.x & 31);
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 0 || MODE == 3) x[i] = fma(x[i], a, b);
                if (MODE == 1 || MODE == 4) x[i] = fma(x[(i + 1) % N], x[(i + 2) % N], x[i]);
                if (MODE == 2) x[i] = fma(x[(i + 1) % N], a, x[i]);
                if (MODE == 5) x[i] = (i & 1) ? x[i] * a : x[i] + b;
                if ((MODE == 3 || MODE == 4) && (i % 8) == 0) acc += sp[((it + r + i) & 31) * 33];
            }
        }
    }
    long long t1 = clock64();
    double s = acc;
#pragma unroll
    for (int i = 0; i < N; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* what, int warps, double* out, long long* cyc) {
    const int iters = 2000;
    mix<MODE><<<1, 32 * warps>>>(out, cyc, iters, 0.9999999, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 4 * N * warps;
    printf("%-58s warps %2d: %6.3f warp-instr(fp64)/clk/SM\n", what, warps, ops / c);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1 << 12);
    for (int warps : {4, 12, 16, 32}) {
        run<0>("A fma(x, const, const)", warps, out, cyc);
        run<1>("B fma(x[i+1], x[i+2], x[i])  three distinct registers", warps, out, cyc);
        run<2>("C fma(x[i+1], const, x[i])   two distinct registers", warps, out, cyc);
        run<3>("D A + one LDS.64 per 8 DFMA", warps, out, cyc);
        run<4>("E B + one LDS.64 per 8 DFMA", warps, out, cyc);
        run<5>("F DADD / DMUL alternating (x op const)", warps, out, cyc);
    }
    return 0;
}
