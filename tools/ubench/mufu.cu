// Micro-benchmark: clocks per warp-instruction of MUFU.EX2, F2FP (cvt.rn.bf16x2.f32), FFMA2 on one SM, as a function of
// the number of warps per scheduler.  nvcc -arch=sm_100a -O3 -o mufu tools/ubench/mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int OP>
__global__ void k(float* out, long long* clk, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 1) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 15])); acc ^= r; }
            if (OP == 2) { unsigned long long a, b; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x[i]), "f"(x[(i + 1) & 15]));
                           asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(b) : "l"(a)); float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(b)); x[i] = lo; }
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 1024);
    const char* names[] = {"MUFU.EX2", "F2FP.BF16.PACK", "FFMA2", "FFMA"};
    for (int op = 0; op < 4; ++op)
        for (int warps : {4, 8, 16, 32}) {
            const int iters = 2000;
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) k<0><<<1, warps * 32>>>(out, clk, iters);
                if (op == 1) k<1><<<1, warps * 32>>>(out, clk, iters);
                if (op == 2) k<2><<<1, warps * 32>>>(out, clk, iters);
                if (op == 3) k<3><<<1, warps * 32>>>(out, clk, iters);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
            const double per_sched = (double)c / (iters * 16.0 * (warps / 4));
            printf("%-16s %2d warps/SM (%d per scheduler): %6.2f clk per warp-instruction per scheduler -> %5.1f lanes/clk/SM\n",
                   names[op], warps, warps / 4, per_sched, 4 * 32 / per_sched);
        }
    return 0;
}
