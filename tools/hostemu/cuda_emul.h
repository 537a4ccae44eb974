// cuda_emul.h -- a small emulation of the CUDA execution model on host threads.
//
// DEVELOPMENT AID ONLY.  The container this project is developed in has no GPU; this header lets
// g++ compile the project's .cu/.cuh files unchanged (-DEGDST_HOSTEMU -x c++ -std=c++20) and run the
// kernels with one std::thread per CUDA thread, one block at a time, so that kernel *logic*
// (indexing, barriers, shuffles, scans, atomics) can be debugged before a GPU run.  The product
// (egdst_b200/capi.py) only ever loads the nvcc-built library and fails loudly without it; nothing
// under tools/ is imported by the package, the tests' checker, or the benchmark.
#pragma once

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
inline bool __all_sync(unsigned, int pred);
inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
inline void *emu_dyn_smem = nullptr;
#define EGDST_DYN_SMEM(type, name) type *name = (type *)emu_dyn_smem
#define EGDST_LDCG(p) (*(p))
#define EGDST_GRID_CONSTANT
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };

inline thread_local uint3_ threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

struct EmuWarp {
    uint64_t buf[32];
    std::unique_ptr<std::barrier<>> bar;
    int nlanes;
};
struct EmuBlock {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<EmuWarp> warps;
};
inline EmuBlock *emu_block = nullptr;
inline thread_local int emu_lin = 0;

inline void __syncthreads() { emu_block->bar->arrive_and_wait(); }
inline EmuWarp &emu_warp() { return emu_block->warps[emu_lin >> 5]; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp().bar->arrive_and_wait(); }

template <class T>
inline uint64_t emu_bits(T v) { uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T emu_from(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

template <class T>
inline T emu_exchange(T v, int src) {
    EmuWarp &W = emu_warp();
    const int lane = emu_lin & 31;
    W.buf[lane] = emu_bits(v);
    W.bar->arrive_and_wait();
    T r = (src >= 0 && src < W.nlanes) ? emu_from<T>(W.buf[src]) : v;
    W.bar->arrive_and_wait();
    return r;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, int o) { const int lane = emu_lin & 31; return emu_exchange(v, lane + o < 32 ? lane + o : -1); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) { const int lane = emu_lin & 31; return emu_exchange(v, lane ^ o); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int o) { const int lane = emu_lin & 31; return emu_exchange(v, lane - o); }
inline unsigned __ballot_sync(unsigned, int pred) {
    EmuWarp &W = emu_warp();
    const int lane = emu_lin & 31;
    W.buf[lane] = pred ? 1 : 0;
    W.bar->arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < W.nlanes; i++) if (W.buf[i]) r |= 1u << i;
    W.bar->arrive_and_wait();
    return r;
}
inline bool __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }

inline int atomicAdd(int *p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
inline double atomicAdd(double *p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return std::atomic_ref<unsigned long long>(*p).fetch_add(v); }
inline int atomicCAS(int *p, int cmp, int val) { std::atomic_ref<int>(*p).compare_exchange_strong(cmp, val); return cmp; }
inline int atomicMax(int *p, int v) {
    std::atomic_ref<int> a(*p);
    int old = a.load();
    while (v > old && !a.compare_exchange_weak(old, v)) {}
    return old;
}
inline int atomicMin(int *p, int v) {
    std::atomic_ref<int> a(*p);
    int old = a.load();
    while (v < old && !a.compare_exchange_weak(old, v)) {}
    return old;
}

// ---- launch ------------------------------------------------------------------------------------
inline void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    std::vector<unsigned char> dyn(smem + 16);
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                EmuBlock B;
                B.bar = std::make_unique<std::barrier<>>(nthreads);
                B.warps.resize(nwarps);
                for (int w = 0; w < nwarps; w++) {
                    B.warps[w].nlanes = std::min(32, nthreads - 32 * w);
                    B.warps[w].bar = std::make_unique<std::barrier<>>(B.warps[w].nlanes);
                }
                emu_block = &B;
                emu_dyn_smem = dyn.data();
                std::vector<std::thread> ts;
                ts.reserve(nthreads);
                for (int t = 0; t < nthreads; t++)
                    ts.emplace_back([&, t] {
                        emu_lin = t;
                        threadIdx.x = t % block.x; threadIdx.y = (t / block.x) % block.y; threadIdx.z = t / (block.x * block.y);
                        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                        blockDim = block; gridDim = grid;
                        body();
                        B.warps[t >> 5].bar->arrive_and_drop();
                        B.bar->arrive_and_drop();
                    });
                for (auto &th : ts) th.join();
                emu_block = nullptr;
            }
}
#define EGDST_LAUNCH(kernel, grid, block, smem, stream, ...) emu_launch((grid), (block), (size_t)(smem), [&] { kernel(__VA_ARGS__); })

// ---- runtime API subset --------------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
typedef int cudaEvent_t;
typedef void *cudaGraphExec_t;
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = 0; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 2; return 0; }
