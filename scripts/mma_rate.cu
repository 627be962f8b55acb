// Microbenchmark: issue rate of the legacy warp-level MMAs on this GPU (per SM), the number that bounds a 3xTF32
// skinny GEMM built on mma.sync.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wmar_b200/csrc/build/mma_rate scripts/mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int KIND, int NACC>
__global__ void __launch_bounds__(512, 1) k(float *out, int iters, long long *cyc) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[i][e] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int NACC>
void run(const char *name, int warps) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<KIND, NACC><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND, NACC><<<148, warps * 32>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * NACC * warps;
    printf("%-28s warps %2d chains %d : %.2f cycles per MMA per SM, %.2f MMA/ns/SM, chain latency %.1f cycles (%s)\n", name, warps, NACC,
           (double)c / n, n / (ms * 1e6), (double)c / iters / 1.0 / (NACC > 0 ? 1 : 1), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>("tf32 m16n8k8", 1);
    run<0, 1>("tf32 m16n8k8", 16);
    run<0, 2>("tf32 m16n8k8", 16);
    run<0, 4>("tf32 m16n8k8", 16);
    run<0, 8>("tf32 m16n8k8", 16);
    run<0, 4>("tf32 m16n8k8", 8);
    run<0, 4>("tf32 m16n8k8", 4);
    run<3, 4>("tf32 m16n8k4", 16);
    run<1, 1>("bf16 m16n8k16", 1);
    run<1, 4>("bf16 m16n8k16", 16);
    run<1, 8>("bf16 m16n8k16", 16);
    run<2, 4>("f16 m16n8k16", 16);
    return 0;
}
