// Microbenchmark: issue rate of scalar FFMA/FADD versus the packed FFMA2/FADD2 forms on sm_100a,
// and LDS.64 vs LDS.128 shared-memory reads.  Prints warp-instructions per cycle per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void __launch_bounds__(512, 1) k(float2* out, float2 s, long long* cyc) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 m = s, c = make_float2(s.y, s.x);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }   // 2 FFMA
            if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);                                           // 1 FFMA2
            if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                         // 2 FADD
            if (MODE == 3) a[i] = __fadd2_rn(a[i], c);                                               // 1 FADD2
            if (MODE == 4) { a[i].x = a[i].x * m.x; a[i].y = a[i].y * m.y; }                         // 2 FMUL
            if (MODE == 5) a[i] = __fmul2_rn(a[i], m);                                               // 1 FMUL2
        }
    }
    long long t1 = clock64();
    float2 r = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// shared-memory read rate: every thread reads consecutive 8- or 16-byte words
template <int W> __global__ void __launch_bounds__(512, 1) ks(float* out, long long* cyc, int stride) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 8192; i += 512) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1024; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int idx = (threadIdx.x * stride + r * 512 + it) & 8191;
            if (W == 16) { float4 v = sm[idx]; acc += v.x + v.y + v.z + v.w; }
            else { float2 v = reinterpret_cast<float2*>(sm)[idx * 2]; float2 w = reinterpret_cast<float2*>(sm)[idx * 2 + 1]; acc += v.x + v.y + w.x + w.y; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float2* out; long long* cyc; cudaMalloc(&out, 148 * 512 * sizeof(float2)); cudaMalloc(&cyc, 148 * 8);
    long long h[148];
    const char* names[6] = {"FFMA x2", "FFMA2", "FADD x2", "FADD2", "FMUL x2", "FMUL2"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            float2 s = make_float2(1.0001f, 0.9999f);
            switch (mode) {
                case 0: k<0><<<148, 512>>>(out, s, cyc); break; case 1: k<1><<<148, 512>>>(out, s, cyc); break;
                case 2: k<2><<<148, 512>>>(out, s, cyc); break; case 3: k<3><<<148, 512>>>(out, s, cyc); break;
                case 4: k<4><<<148, 512>>>(out, s, cyc); break; case 5: k<5><<<148, 512>>>(out, s, cyc); break;
            }
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double ops = 8.0 * ITERS * 16;   // float2 updates per SM-warp-slot: 16 warps x 8 x ITERS
        printf("%-8s cycles %.0f  float2-updates/cycle/SM %.3f (warp granularity)  err=%s\n", names[mode], avg, ops / avg, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFuncSetAttribute(ks<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    cudaFuncSetAttribute(ks<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    for (int stride = 1; stride <= 17; stride += 16) for (int w = 8; w <= 16; w += 8) {
        for (int rep = 0; rep < 2; ++rep) {
            if (w == 16) ks<16><<<148, 512, 8192 * 16>>>((float*)out, cyc, stride); else ks<8><<<148, 512, 8192 * 16>>>((float*)out, cyc, stride);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("LDS.%d stride %d: cycles %.0f  bytes/cycle/SM %.1f err=%s\n", w * 8, stride, avg, 1024.0 * 16 * 512 * 16 / avg, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
