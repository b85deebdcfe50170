// TEST INFRASTRUCTURE — a minimal CUDA execution model on the CPU, enough to run the SOURCE of the small index / reduction
// kernels (csrc/arah_image.cu, csrc/arah_loss.cu) in the build container, which has no GPU.
//
// One OS thread per CUDA thread, one block at a time: `__syncthreads()` is a std::barrier over the block, warp shuffles / ballots
// exchange through a per-warp buffer between two warp barriers (so a shuffle that not all 32 lanes reach dead-locks here exactly
// where it would be undefined on the GPU), `__shared__` is function-static storage (blocks run one after another), atomics take a
// global lock, a thread that leaves the kernel drops out of the barriers.  Launches go through ARAH_LAUNCH, which the .cu files
// expand to `kernel<<<grid, block, 0, stream>>>(...)` under nvcc.  This checks indexing, launch geometry, reduction trees and
// argument marshalling of the very code nvcc compiles — not performance, not the memory model.  Never part of the product.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define ARAH_CUDA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; EmuDim3() {} EmuDim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef EmuDim3 dim3;
inline thread_local EmuDim3 threadIdx, blockIdx;
inline EmuDim3 blockDim, gridDim;

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

namespace emu {
struct BlockCtx {
    std::barrier<> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<unsigned long long> xchg;             // one 8-byte slot per thread
    explicit BlockCtx(int n) : block_bar(n), xchg(n) { for (int w = 0; w < (n + 31) / 32; ++w) { const int m = n - 32 * w < 32 ? n - 32 * w : 32; warp_bar.emplace_back(new std::barrier<>(m)); } }
};
inline BlockCtx* ctx = nullptr;
inline std::mutex atomic_lock;
inline std::barrier<>& warp() { return *ctx->warp_bar[threadIdx.x >> 5]; }

template <class F> void launch(EmuDim3 grid, EmuDim3 block, F&& body) {
    gridDim = grid; blockDim = block;
    const int n = (int)block.x;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            BlockCtx c(n);
            ctx = &c;
            std::vector<std::thread> th;
            th.reserve(n);
            for (int t = 0; t < n; ++t)
                th.emplace_back([&, t] {
                    threadIdx = EmuDim3((unsigned)t); blockIdx = EmuDim3(bx, by);
                    body();
                    c.warp_bar[t >> 5]->arrive_and_drop();           // an exited thread no longer takes part in barriers
                    c.block_bar.arrive_and_drop();
                });
            for (auto& x : th) x.join();
            ctx = nullptr;
        }
}
}  // namespace emu

#define ARAH_LAUNCH(kernel, grid, block, stream, ...) emu::launch(EmuDim3(grid), EmuDim3(block), [&] { kernel(__VA_ARGS__); })

inline void __syncthreads() { emu::ctx->block_bar.arrive_and_wait(); }

template <class T> T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "8-byte exchange slots");
    const unsigned t = threadIdx.x;
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::ctx->xchg[t] = raw;
    emu::warp().arrive_and_wait();
    const unsigned src = (t & ~31u) | ((t ^ (unsigned)lane_mask) & 31u);
    raw = src < emu::ctx->xchg.size() ? emu::ctx->xchg[src] : raw;
    emu::warp().arrive_and_wait();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}

inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned t = threadIdx.x;
    emu::ctx->xchg[t] = pred ? 1ull : 0ull;
    emu::warp().arrive_and_wait();
    unsigned m = 0;
    for (unsigned l = 0; l < 32; ++l) { const unsigned s = (t & ~31u) | l; if (s < emu::ctx->xchg.size() && emu::ctx->xchg[s]) m |= 1u << l; }
    emu::warp().arrive_and_wait();
    return m;
}

inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
template <class T> T __ldg(const T* p) { return *p; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }

template <class T> T atomicMin(T* p, T v) { std::lock_guard<std::mutex> g(emu::atomic_lock); const T o = *p; if (v < o) *p = v; return o; }
template <class T> T atomicMax(T* p, T v) { std::lock_guard<std::mutex> g(emu::atomic_lock); const T o = *p; if (v > o) *p = v; return o; }
template <class T> T atomicAdd(T* p, T v) { std::lock_guard<std::mutex> g(emu::atomic_lock); const T o = *p; *p = o + v; return o; }
