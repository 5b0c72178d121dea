// Dev microbenchmark (not product): issue/pipe throughput of packed fp32x2 vs scalar fp32 on sm_100a,
// and L2-resident read bandwidth.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o pipes
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(float* out, int iters)
{
    float a[8], b = 1.0001f, c = 0.5f;
    u64 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
    const u64 pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), pc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);                   // FFMA
            if (MODE == 1) p[i] = ffma2(p[i], pb, pc);                // FFMA2
            if (MODE == 2) a[i] = a[i] + c;                           // FADD
            if (MODE == 3) p[i] = fadd2(p[i], pc);                    // FADD2
            if (MODE == 4) { a[i] = fmaf(a[i], b, c); p[i] = fadd2(p[i], pc); }   // mixed
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_l2read(const float4* __restrict__ in, float* out, long long n, int reps)
{
    float4 acc = make_float4(0, 0, 0, 0);
    for (int r = 0; r < reps; ++r)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const float4 v = __ldcg(&in[i]);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE> void run_pipe(const char* name, float* d_out, double flop_per_inst)
{
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_pipe<MODE><<<blocks, 256>>>(d_out, 64);
    cudaEventRecord(e0);
    k_pipe<MODE><<<blocks, 256>>>(d_out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * 256 * iters * 8 * (MODE == 4 ? 2 : 1);
    printf("%-8s %8.3f ms  %7.2f G thread-inst/s  = %6.1f inst/clk/SM @1.965GHz   %7.2f TFLOP/s\n", name, ms, inst / ms / 1e6,
           inst / ms / 1e6 * 1e9 / 148 / 1.965e9, inst * flop_per_inst / ms / 1e9);
}

int main()
{
    float* d_out; cudaMalloc(&d_out, 148 * 8 * 256 * sizeof(float) * 4);
    run_pipe<0>("FFMA", d_out, 2);
    run_pipe<1>("FFMA2", d_out, 4);
    run_pipe<2>("FADD", d_out, 1);
    run_pipe<3>("FADD2", d_out, 2);
    run_pipe<4>("FFMA+FADD2", d_out, 2);
    for (long long mb : {16LL, 32LL, 48LL, 64LL, 96LL, 256LL, 1024LL}) {
        const long long n = mb * 1024 * 1024 / 16;
        float4* d; cudaMalloc(&d, n * 16); cudaMemset(d, 0, n * 16);
        const int reps = mb <= 96 ? 20 : 4;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_l2read<<<148 * 8, 256>>>(d, d_out, n, 2);
        cudaEventRecord(e0);
        k_l2read<<<148 * 8, 256>>>(d, d_out, n, reps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("read %5lld MB x%d : %8.3f ms  %8.1f GB/s\n", mb, reps, ms, (double)n * 16 * reps / ms / 1e6);
        cudaFree(d);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
