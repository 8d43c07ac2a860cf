// TEST INFRASTRUCTURE ONLY -- never part of the product path (libsushi_b200.so is built by nvcc from
// sushi_b200/csrc and has no CPU fallback).
//
// Host stand-ins for the CUDA device environment, so that g++ can compile the kernels of
// sushi_b200/csrc/sb_fused2.cu unchanged and tests/test_kernel_emulation.py can run their LOGIC on the CPU:
// the 512 CUDA threads of a CTA as fibers on one OS thread per warp (or as 512 OS threads, -DSB_EMU_THREADS, for
// ThreadSanitizer), std::barrier for bar.sync, a rendezvous per warp for shuffles, plain arrays for shared and
// tensor memory.  It checks index algebra, bookkeeping (mbarrier phases, tensor-memory
// columns, which thread parks what) and the arithmetic of the screening loops; it says nothing about races,
// memory-model fences, alignment rules of the copy engine or performance -- the GPU tests do that.
#pragma once
#include <cuda_runtime.h>          // host mode: vector types, make_float2/4, no device declarations
#include <stdint.h>

#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __launch_bounds__(...)

namespace emu {

constexpr int kThreads = 512, kWarps = kThreads / 32;

struct MBar { unsigned count = 0, pending = 0, phase = 0; long long tx = 0; };

// Two ways to run the 512 CUDA threads of a CTA:
//   default          16 OS threads (one per warp), each switching between its 32 lanes as fibers (ucontext) at
//                    every rendezvous -- fast: no 512-thread futex storms;
//   -DSB_EMU_THREADS 512 OS threads -- what ThreadSanitizer needs (tests/emu/run_tsan.py).
struct Warp {
#ifdef SB_EMU_THREADS
    std::barrier<> bar{32};
#else
    int count = 0, gen = 0;              // rendezvous of the 32 lanes (they run one at a time on one OS thread)
    int bar_count = 0, bar_gen = 0;      // lanes waiting at the CTA barrier
#endif
    unsigned long long slot[32];
};

struct Cta {
#ifdef SB_EMU_THREADS
    std::barrier<> bar{kThreads};
#else
    std::barrier<> bar{kWarps};
#endif
    std::unique_ptr<Warp> warps[kWarps];
    std::mutex mu;
    std::map<const void*, MBar> mbars;
    std::vector<float> tmem;           // [128 lanes][512 columns]
    unsigned tmem_cols = 0;
    std::vector<std::string> errors;
    Cta() : tmem(128 * 512) { for (auto& w : warps) w = std::make_unique<Warp>(); }
    void fail(const char* what) { std::lock_guard<std::mutex> g(mu); if (errors.size() < 16) errors.push_back(what); }
};

inline Cta*& cta() { static Cta* c = nullptr; return c; }
inline thread_local int t_lane = 0, t_warp = 0;

#ifdef SB_EMU_THREADS
inline void yield_lane() { std::this_thread::yield(); }
inline void warp_rendezvous() { cta()->warps[t_warp]->bar.arrive_and_wait(); }
inline void cta_barrier() { cta()->bar.arrive_and_wait(); }
#else
void yield_lane();                      // emu_driver.cpp: back to the warp's scheduler, which resumes the next lane
inline void warp_rendezvous() {
    Warp& w = *cta()->warps[t_warp];
    const int gen = w.gen;
    if (++w.count == 32) { w.count = 0; ++w.gen; }
    else while (w.gen == gen) yield_lane();
}
inline void cta_barrier() {
    Warp& w = *cta()->warps[t_warp];
    const int gen = w.bar_gen;
    if (++w.bar_count == 32) { cta()->bar.arrive_and_wait(); w.bar_count = 0; ++w.bar_gen; }   // the warp's last lane waits for the other warps
    else while (w.bar_gen == gen) yield_lane();
}
#endif

template <typename T> T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of up to 64 bits");
    Warp& w = *cta()->warps[t_warp];
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.slot[t_lane] = raw;
    warp_rendezvous();
    raw = w.slot[src];
    warp_rendezvous();
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}

// Runs body(t) for t = 0 .. 511 as the threads of the current CTA (emu_driver.cpp)
void run_cta(const std::function<void(int)>& body);

}  // namespace emu

// ---- built-in variables -------------------------------------------------------------------------------
inline thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
inline dim3 gridDim(1, 1, 1), blockDim(emu::kThreads, 1, 1);

// ---- intrinsics the kernels call directly -----------------------------------------------------------
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) { return emu::shfl(v, emu::t_lane ^ o); }
inline int __any_sync(unsigned, int pred) {
    int v = pred ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) v |= emu::shfl(v, emu::t_lane ^ o);
    return v;
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, int o) { return emu::shfl(v, emu::t_lane >= o ? emu::t_lane - o : emu::t_lane); }
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned out = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7))) & 0xffu;
        if (sel & 8) byte = (byte & 0x80u) ? 0xffu : 0u;      // sign replication mode
        out |= byte << (8 * i);
    }
    return out;
}
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (shift & 31u));
}
inline int __float2int_rn(float f) { return (int)std::nearbyintf(f); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline void __nanosleep(unsigned) { emu::yield_lane(); }
inline void __syncthreads() { emu::cta_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_rendezvous(); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
