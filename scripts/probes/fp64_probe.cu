// fp64_probe.cu -- B200 fp64 microbenchmarks that drive the IPM kernel design:
// DFMA vs DMMA (mma.sync.m8n8k4.f64) throughput and dependent-issue latency, rsqrt/div cost, shuffle and
// shared-memory latency.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dfma(double* out, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters, double a, double b)
{
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: CH DFMA chains + CM DMMA chains in the same warp (do the pipes overlap?)
template <int CH, int CM>
__global__ void k_mixed(double* out, int iters, double a, double b)
{
    double x[CH], c0[CM], c1[CM];
#pragma unroll
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < CM; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CM; i++) dmma(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i];
#pragma unroll
    for (int i = 0; i < CM; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
__global__ void k_special(double* out, int iters, double a)
{
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (OP == 0) x[i] = rsqrt(x[i]) + a;
            if (OP == 1) x[i] = a / x[i] + 1.0;
            if (OP == 2) x[i] = sqrt(x[i]) + a;
            if (OP == 3) { double s, c; sincos(x[i], &s, &c); x[i] = s + c + a; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}

// latency probes: one warp, dependent chain, clock64
__global__ void k_lat(long long* res, double* sink, double a, double b)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = lane; sm[lane + 32] = 0;
    __syncwarp();
    double x = lane * 1e-3, c0 = x, c1 = 0;
    const int R = 256;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) x = fma(x, a, b);
    long long t1 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) dmma(c0, c1, a, b);
    long long t2 = clock64();
    double y = x;
#pragma unroll 16
    for (int i = 0; i < R; i++) y = __shfl_xor_sync(0xffffffffu, y, 1) + a;
    long long t3 = clock64();
    int idx = lane;
    double z = 0;
#pragma unroll 16
    for (int i = 0; i < R; i++) { z += sm[idx]; idx = (int)z & 31; }
    long long t4 = clock64();
    double w = 1.0 + lane;
#pragma unroll 16
    for (int i = 0; i < R; i++) w = rsqrt(w) + a;
    long long t5 = clock64();
    if (lane == 0) {
        res[0] = (t1 - t0) / R; res[1] = (t2 - t1) / R; res[2] = (t3 - t2) / R; res[3] = (t4 - t3) / R; res[4] = (t5 - t4) / R;
    }
    sink[lane] = x + c0 + c1 + y + z + w;
}

template <class F>
float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, clk);
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    const int iters = 4096;
    for (int wps : {4, 8, 16, 32}) {
        const int blocks = 148 * 2, threads = wps * 32 / 2;
        float ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        double fl = 2.0 * blocks * threads * 8.0 * iters;
        printf("DFMA  warps/SM %2d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM\n", wps, ms, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / (clk * 1e3) / 148);
        ms = timeit([&] { k_dmma<4><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        fl = 2.0 * blocks * (threads / 32) * 4.0 * iters * 256.0;
        printf("DMMA  warps/SM %2d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM\n", wps, ms, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / (clk * 1e3) / 148);
        ms = timeit([&] { k_mixed<8, 4><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        fl = 2.0 * blocks * (threads / 32) * iters * (4.0 * 256.0 + 8.0 * 32.0);
        printf("MIXED warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
    }
    const char* names[] = {"rsqrt", "div", "sqrt", "sincos"};
    for (int op = 0; op < 4; op++) {
        const int blocks = 148 * 2, threads = 512;
        float ms = 0;
        if (op == 0) ms = timeit([&] { k_special<0><<<blocks, threads>>>(out, 512, 0.5); });
        if (op == 1) ms = timeit([&] { k_special<1><<<blocks, threads>>>(out, 512, 0.5); });
        if (op == 2) ms = timeit([&] { k_special<2><<<blocks, threads>>>(out, 512, 0.5); });
        if (op == 3) ms = timeit([&] { k_special<3><<<blocks, threads>>>(out, 512, 0.5); });
        double ops = (double)blocks * threads * 4.0 * 512;
        printf("%-6s: %.3f ms  %.1f ops/clk/SM  (= %.1f DFMA-equivalents each at 64/clk/SM)\n", names[op], ms,
               ops / (ms * 1e-3) / (clk * 1e3) / 148, 64.0 / (ops / (ms * 1e-3) / (clk * 1e3) / 148));
    }
    long long* res;
    cudaMallocManaged(&res, 8 * sizeof(long long));
    k_lat<<<1, 32>>>(res, out, 1.000001, 1e-9);
    cudaDeviceSynchronize();
    printf("latency (cycles, dependent issue): DFMA %lld  DMMA %lld  shfl64+add %lld  lds+add+cvt %lld  rsqrt+add %lld\n", res[0], res[1], res[2], res[3], res[4]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
