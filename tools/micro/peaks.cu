// On-box ceilings that MEASURED_PEAKS.json does not hold: FP64 FMA throughput and the rate of
// dependent-free random 16-byte gathers from an L2-resident table (what bounds the EGM step and the simulator).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_gather(const double2 *tab, unsigned n, double *out, int iters) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double acc = 0;
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        const double2 v = tab[s % n];  // independent of the previous load: throughput, not latency
        acc += v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    double *out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    const int it = 20000;
    k_dfma<<<148 * 2, 1024>>>(out, 100);
    cudaEventRecord(e0); k_dfma<<<148 * 2, 1024>>>(out, it); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 8 * it * 148.0 * 2 * 1024;
    printf("{\"fp64_fma_tflops\": %.2f", flop / (ms * 1e-3) / 1e12);
    for (int mb = 32; mb <= 64; mb += 32) {  // table sizes that stay in the 126 MB L2
        const unsigned n = (unsigned)((size_t)mb << 20) / 16;
        double2 *tab; cudaMalloc(&tab, (size_t)n * 16); cudaMemset(tab, 0, (size_t)n * 16);
        k_gather<<<148 * 4, 512>>>(tab, n, out, 10);
        cudaEventRecord(e0); k_gather<<<148 * 4, 512>>>(tab, n, out, 2000); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"gather16_l2_%dMB_Gps\": %.1f", mb, 2000.0 * 148 * 4 * 512 / (ms * 1e-3) / 1e9);
        cudaFree(tab);
    }
    printf(", \"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
