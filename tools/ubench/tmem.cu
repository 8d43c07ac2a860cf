#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(a), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
                   "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__global__ void k(float* out) {
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&s_taddr, 512);
    tmem_fence_before(); __syncthreads(); tmem_fence_after();
    const uint32_t base = s_taddr;
    if (warp < 4) {
        float v[8];
        for (int m = 0; m < 64; ++m) {
            for (int j = 0; j < 8; ++j) v[j] = tid * 1000.0f + m * 8 + j;
            tmem_st8(base + ((uint32_t)(warp * 32) << 16) + m * 8, v);
        }
        tmem_wait_st();
    }
    tmem_fence_before(); __syncthreads(); tmem_fence_after();
    // warps 4..7 read what warps 0..3 wrote (same lane quarter = warp % 4)
    if (warp >= 4) {
        float v[16];
        const int q = warp & 3;
        for (int c = 0; c < 32; ++c) {
            tmem_ld16(base + ((uint32_t)(q * 32) << 16) + c * 16, v);
            tmem_wait_ld();
            for (int j = 0; j < 16; ++j) out[((tid - 128) * 32 + c) * 16 + j] = v[j];
        }
    }
    tmem_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(base, 512);
}

__global__ void __launch_bounds__(640, 1) kt(long long* cyc, float* sink) {
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&s_taddr, 512);
    tmem_fence_before(); __syncthreads(); tmem_fence_after();
    const uint32_t base = s_taddr;
    long long t0 = clock64();
    if (warp >= 16) {            // 4 producer warps fill 256 columns (128 KB)
        float v[8];
        for (int m = 0; m < 32; ++m) {
            for (int j = 0; j < 8; ++j) v[j] = tid + m * 8 + j;
            tmem_st8(base + ((uint32_t)((warp & 3) * 32) << 16) + m * 8, v);
        }
        tmem_wait_st();
    }
    long long t1 = clock64();
    tmem_fence_before(); __syncthreads(); tmem_fence_after();
    long long t2 = clock64();
    float acc = 0.f;
    if (warp < 16) {             // 16 consumer warps drain them: 64 columns each
        float v[16];
        for (int c = 0; c < 4; ++c) {
            tmem_ld16(base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64 + c * 16, v);
            tmem_wait_ld();
            for (int j = 0; j < 16; ++j) acc += v[j];
        }
    }
    long long t3 = clock64();
    sink[blockIdx.x * 640 + tid] = acc;
    if (tid == 0 || tid == 512) { cyc[blockIdx.x * 4 + (tid ? 2 : 0)] = t1 - t0; cyc[blockIdx.x * 4 + (tid ? 3 : 1)] = t3 - t2; }
    tmem_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(base, 512);
}
int main() {
    {
        long long* c; float* s; cudaMalloc(&c, 148 * 4 * 8); cudaMalloc(&s, 148 * 640 * 4);
        for (int rep = 0; rep < 2; ++rep) { kt<<<148, 640>>>(c, s); cudaDeviceSynchronize(); }
        long long h[148 * 4]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        printf("kt: %s; CTA0: 128KB tcgen05.st by 4 warps %lld cycles (thread 512), tcgen05.ld by 16 warps %lld cycles (thread 0)\n",
               cudaGetErrorString(cudaGetLastError()), h[2], h[1]);
    }
    float* d; cudaMalloc(&d, 128 * 512 * 4);
    k<<<1, 256>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    static float h[128 * 512];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 128; ++t) for (int c = 0; c < 512; ++c) if (h[t * 512 + c] != t * 1000.0f + c) { if (bad < 5) printf("mismatch t=%d c=%d got %f\n", t, c, h[t*512+c]); ++bad; }
    printf("tmem roundtrip mismatches: %d\n", bad);
    // timing: cycles for 128KB st + ld with 16 warps
    return 0;
}
