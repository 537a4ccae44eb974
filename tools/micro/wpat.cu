// Write-pattern microbenchmark: every lane owns an "agent record" of NT*REC bytes and fills it P periods at a time,
// like the simulator's output stream.  Reports achieved DRAM write bandwidth per chunk size P*REC.
#include <cstdio>
#include <cuda_runtime.h>
template <int P>
__global__ void k(double *out, long nagents, int nt, int nso) {
    const int lane = threadIdx.x & 31;
    const long ntiles = nagents / 32;
    const int k16 = lane & 7, a0 = lane >> 3;
    for (long tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles; tile += (long)gridDim.x * (blockDim.x >> 5)) {
        for (int it = 0; it < nt; it += P) {
            // 32 agents x P periods x nso doubles, written as 16-byte pieces: agent = 4*t + lane/8, piece = lane%8 (+8*j)
            double *dst = out + ((size_t)tile * 32 * nt + it) * nso;
            const int pieces = P * nso / 2;  // per agent
            for (int t = 0; t < 8; t++) {
                const int a = 4 * t + a0;
                for (int pc = k16; pc < pieces; pc += 8)
                    __stcs(reinterpret_cast<double2 *>(dst + (size_t)a * nt * nso + 2 * pc), make_double2((double)it, (double)a));
            }
            // emulate the compute between visits
            __nanosleep(P * 400);
        }
    }
}
int main() {
    const long nagents = 4000000; const int nt = 50, nso = 14;
    double *out; cudaMalloc(&out, sizeof(double) * nagents * nt * nso);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = 8.0 * nagents * nt * nso / 1e9;
#define RUN(P) { k<P><<<592, 256>>>(out, nagents, nt, nso); cudaEventRecord(e0); for (int r = 0; r < 3; r++) k<P><<<592, 256>>>(out, nagents, nt, nso); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); printf("P=%2d chunk=%4d B  %.2f ms  %.0f GB/s\n", P, P * nso * 8, ms / 3, gb / (ms / 3) * 1e3); }
    RUN(1) RUN(2) RUN(5) RUN(10) RUN(25) RUN(50)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
