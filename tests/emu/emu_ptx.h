// TEST INFRASTRUCTURE ONLY.  Host implementations of the wrappers in sushi_b200/csrc/sb_ptx.cuh (same names,
// same signatures), selected by -DSB_EMULATE: see emu_cuda.h.
#pragma once
#include "emu_cuda.h"

namespace sbf {

inline float rsqrt_fast(float x) { return 1.0f / std::sqrt(x); }

inline void cp_async16(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 16); }
inline void cp_async4(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 4); }
inline void cp_async_commit_wait_all() {}

// ---- mbarrier + bulk copies: phase completes when every expected arrival and every expected byte is in
inline void mbar_complete_if_done(emu::MBar& b) { if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; } }
inline void mbar_init(unsigned long long* bar, unsigned count) {
    std::lock_guard<std::mutex> g(emu::cta()->mu);
    emu::MBar& b = emu::cta()->mbars[bar];
    b = emu::MBar(); b.count = b.pending = count;
}
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    std::lock_guard<std::mutex> g(emu::cta()->mu);
    auto it = emu::cta()->mbars.find(bar);
    if (it == emu::cta()->mbars.end()) { emu::cta()->errors.push_back("mbar_expect_tx on an uninitialised barrier"); return; }
    it->second.tx += bytes; it->second.pending -= 1;
    mbar_complete_if_done(it->second);
}
inline void mbar_arrive(unsigned long long* bar) {
    std::lock_guard<std::mutex> g(emu::cta()->mu);
    emu::MBar& b = emu::cta()->mbars[bar];
    b.pending -= 1;
    mbar_complete_if_done(b);
}
inline void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    if ((reinterpret_cast<uintptr_t>(smem_dst) & 15) || (reinterpret_cast<uintptr_t>(gmem_src) & 15) || (bytes & 15))
        emu::cta()->fail("tma_load_1d: address or size not a multiple of 16");
    std::memcpy(smem_dst, gmem_src, bytes);
    std::lock_guard<std::mutex> g(emu::cta()->mu);
    emu::MBar& b = emu::cta()->mbars[bar];
    b.tx -= bytes;
    mbar_complete_if_done(b);
}
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
    for (long spins = 0;; ++spins) {
        {
            std::lock_guard<std::mutex> g(emu::cta()->mu);
            if (emu::cta()->mbars[bar].phase != parity) return;      // the phase with this parity has completed
        }
        if (spins > 20000000) { emu::cta()->fail("mbar_wait: never completed"); return; }
        emu::yield_lane();
        if ((spins & 63) == 63) std::this_thread::yield();
    }
}
inline void mbar_wait_sleep(unsigned long long* bar, unsigned parity, unsigned) { mbar_wait(bar, parity); }
inline void fence_proxy_async() {}

template <int ID> inline void csync() { emu::cta_barrier(); }     // one-role kernels only (ID 0)

// ---- tensor memory: [128 lanes][512 columns] of 32-bit cells; a warp reaches the 32 lanes of its quarter
inline void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    std::lock_guard<std::mutex> g(emu::cta()->mu);
    if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) emu::cta()->errors.push_back("tmem_alloc: column count not a power of two in 32..512");
    emu::cta()->tmem_cols = ncols;
    *smem_slot = 0;
}
inline void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if (taddr != 0 || ncols != emu::cta()->tmem_cols) emu::cta()->fail("tmem_dealloc: does not match the allocation");
}
inline float* tmem_cell(uint32_t taddr, int n) {
    const unsigned lane0 = taddr >> 16, col = taddr & 0xffffu;
    if (lane0 != 32u * (emu::t_warp & 3)) { emu::cta()->fail("tensor memory: lane field is not this warp's quarter"); return nullptr; }
    if (col + n > emu::cta()->tmem_cols) { emu::cta()->fail("tensor memory: column outside the allocation"); return nullptr; }
    return emu::cta()->tmem.data() + (size_t)(lane0 + emu::t_lane) * 512 + col;
}
inline void tmem_st8(uint32_t taddr, float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
    if (float* c = tmem_cell(taddr, 8)) { c[0] = v0; c[1] = v1; c[2] = v2; c[3] = v3; c[4] = v4; c[5] = v5; c[6] = v6; c[7] = v7; }
}
inline void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    if (float* c = tmem_cell(taddr, 16)) for (int i = 0; i < 16; ++i) v[i] = c[i];
}
inline void tmem_wait_st() {}
inline void tmem_wait_ld() {}
inline void tmem_fence_before() {}
inline void tmem_fence_after() {}
template <int N> inline void setmaxnreg_inc() {}
template <int N> inline void setmaxnreg_dec() {}

}  // namespace sbf
