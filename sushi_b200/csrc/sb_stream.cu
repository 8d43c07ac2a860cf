// Resident streams: upload, running sums (the integral image OpenCV rebuilds on
// every matchTemplate call, reference wav.py:185), and the per-stream block
// spectra that every query against the stream shares.
#include "sb_internal.h"

using namespace sb;

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 4096 samples per CTA

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Phase A: per-tile totals of x and x^2 in fp64 (exact integers for u8 input).
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
k_tile_totals(const T* __restrict__ x, int64_t n, double* __restrict__ tsum, double* __restrict__ tsq) {
    __shared__ double s_a[SCAN_THREADS / 32], s_b[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t j = base + (int64_t)i * SCAN_THREADS + threadIdx.x;   // coalesced
        if (j < n) { double v = (double)x[j]; a += v; b += v * v; }
    }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
        tsum[blockIdx.x] = ta; tsq[blockIdx.x] = tb;
    }
}

// Phase B: exclusive scan of the tile totals.  One CTA walks them in slabs of 1024 x TOT_ITEMS (a 90-minute
// stream has 15 880 tiles: one slab), every thread scanning TOT_ITEMS consecutive totals in registers.
constexpr int TOT_ITEMS = 16;
__global__ void __launch_bounds__(1024)
k_scan_tile_totals(double* __restrict__ tsum, double* __restrict__ tsq, int64_t ntiles) {
    __shared__ double s_a[32], s_b[32];
    __shared__ double carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = 0.0; carry_b = 0.0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < ntiles; base += 1024 * TOT_ITEMS) {
        const int64_t j0 = base + (int64_t)threadIdx.x * TOT_ITEMS;
        double va[TOT_ITEMS], vb[TOT_ITEMS];
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int i = 0; i < TOT_ITEMS; ++i) {
            va[i] = j0 + i < ntiles ? tsum[j0 + i] : 0.0; vb[i] = j0 + i < ntiles ? tsq[j0 + i] : 0.0;
            a += va[i]; b += vb[i];
        }
        double ia = a, ib = b;                          // inclusive warp scan of the threads' totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ta; ib += tb; }
        }
        if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
        __syncthreads();
        if (warp == 0) {
            double wa = s_a[lane], wb = s_b[lane];
            double xa = wa, xb = wb;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
                if (lane >= o) { xa += ta; xb += tb; }
            }
            s_a[lane] = xa - wa; s_b[lane] = xb - wb;   // exclusive warp offsets
        }
        __syncthreads();
        double ea = carry_a + s_a[warp] + (ia - a), eb = carry_b + s_b[warp] + (ib - b);
#pragma unroll
        for (int i = 0; i < TOT_ITEMS; ++i) {
            if (j0 + i < ntiles) { tsum[j0 + i] = ea; tsq[j0 + i] = eb; }
            ea += va[i]; eb += vb[i];
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a = ea; carry_b = eb; }
        __syncthreads();
    }
}

// Phase C: in-tile inclusive scan + tile offset -> pfx[i+1] = (sum, sum of squares).
// Thread t scans SCAN_ITEMS consecutive samples; the results go through shared memory so that the
// 16-byte stores to HBM are coalesced (a blocked arrangement would scatter them 256 bytes apart).
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
k_tile_scan(const T* __restrict__ x, int64_t n, const double* __restrict__ osum, const double* __restrict__ osq,
            double2* __restrict__ pfx) {
    __shared__ double s_a[SCAN_THREADS / 32], s_b[SCAN_THREADS / 32];
    extern __shared__ __align__(16) unsigned char scan_smem[];
    double2* s_out = reinterpret_cast<double2*>(scan_smem);             // [SCAN_TILE + SCAN_TILE/16] padded
    const int64_t tile0 = (int64_t)blockIdx.x * SCAN_TILE;
    const int64_t base = tile0 + (int64_t)threadIdx.x * SCAN_ITEMS;
    double va[SCAN_ITEMS], vb[SCAN_ITEMS];
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t j = base + i;
        double v = j < n ? (double)x[j] : 0.0;
        a += v; b += v * v;
        va[i] = a; vb[i] = b;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
    __syncthreads();
    double wa = 0.0, wb = 0.0;
    for (int w = 0; w < warp; ++w) { wa += s_a[w]; wb += s_b[w]; }
    const double offa = osum[blockIdx.x] + wa + (ia - a);
    const double offb = osq[blockIdx.x] + wb + (ib - b);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int e = threadIdx.x * SCAN_ITEMS + i;
        s_out[e + (e >> 4)] = make_double2(offa + va[i], offb + vb[i]);   // +1 slot per 16: conflict-free
    }
    __syncthreads();
#pragma unroll 4
    for (int e = threadIdx.x; e < SCAN_TILE; e += SCAN_THREADS) {
        const int64_t j = tile0 + e;
        if (j < n) pfx[j + 1] = s_out[e + (e >> 4)];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) pfx[0] = make_double2(0.0, 0.0);
}

// ---- uint8 streams: the same three phases on exact integers, written for HBM bandwidth -------------------
// The running sums cost 16 bytes per sample to write against 1 byte to read, so the scan is a store stream:
// phase C below gives every store instruction of a warp 512 contiguous bytes (lane l owns sample 32*it + l,
// the scan across the lanes is five shuffle steps on 32-bit integers -- sums inside a 4096-sample tile fit:
// 4096 * 255^2 < 2^31), with no trip through shared memory and no barrier between loads and stores, so the
// 8 warps of a CTA stream independently.  Phase A reads 16 bytes per lane and sums with dp4a.
__global__ void __launch_bounds__(SCAN_THREADS)
k_tile_totals_u8(const uint8_t* __restrict__ x, int64_t n, double* __restrict__ tsum, double* __restrict__ tsq) {
    __shared__ unsigned s_a[SCAN_THREADS / 32], s_b[SCAN_THREADS / 32];
    const int64_t j = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * 16;     // stream allocations carry 16 bytes of slack
    unsigned a = 0, b = 0;
    if (j < n) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(x + j));
        if (j + 16 > n) {                            // mask the bytes past the end
            const int keep = (int)(n - j);
            unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int kb = keep - 4 * k;
                w[k] = kb >= 4 ? w[k] : (kb <= 0 ? 0u : (w[k] & (0xffffffffu >> (8 * (4 - kb)))));
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        a = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
        b = __dp4a(v.x, v.x, __dp4a(v.y, v.y, __dp4a(v.z, v.z, __dp4a(v.w, v.w, 0u))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ta = 0, tb = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
        tsum[blockIdx.x] = (double)ta; tsq[blockIdx.x] = (double)tb;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_tile_scan_u8(const uint8_t* __restrict__ x, int64_t n, const double* __restrict__ osum, const double* __restrict__ osq,
               double2* __restrict__ pfx) {
    constexpr int NW = SCAN_THREADS / 32, PER_WARP = SCAN_TILE / NW;        // 512 samples per warp
    __shared__ unsigned s_a[NW], s_b[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)warp * PER_WARP;
    {   // totals of the warps in front of this one inside the tile
        unsigned a = 0, b = 0;
        const int64_t j = w0 + lane * 16;
        if (j < n) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(x + j));
            if (j + 16 > n) {
                const int keep = (int)(n - j);
                unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int kb = keep - 4 * k;
                    w[k] = kb >= 4 ? w[k] : (kb <= 0 ? 0u : (w[k] & (0xffffffffu >> (8 * (4 - kb)))));
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            a = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
            b = __dp4a(v.x, v.x, __dp4a(v.y, v.y, __dp4a(v.z, v.z, __dp4a(v.w, v.w, 0u))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) { s_a[warp] = a; s_b[warp] = b; }
    }
    __syncthreads();
    unsigned cs = 0, cq = 0;                           // integer carry inside the tile (exact)
    for (int w = 0; w < warp; ++w) { cs += s_a[w]; cq += s_b[w]; }
    const double bs = osum[blockIdx.x], bq = osq[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) pfx[0] = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int it = 0; it < PER_WARP / 32; ++it) {
        const int64_t j = w0 + it * 32 + lane;
        if (w0 + it * 32 >= n) break;                  // warp-uniform
        const unsigned v = j < n ? (unsigned)__ldg(x + j) : 0u;
        unsigned s = v, q = v * v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned ts = __shfl_up_sync(0xffffffffu, s, o), tq = __shfl_up_sync(0xffffffffu, q, o);
            if (lane >= o) { s += ts; q += tq; }
        }
        if (j < n) pfx[j + 1] = make_double2(bs + (double)(cs + s), bq + (double)(cq + q));
        cs += __shfl_sync(0xffffffffu, s, 31);
        cq += __shfl_sync(0xffffffffu, q, 31);
    }
}

// Centred float rows for the block spectra: row k holds image[kB .. kB+2B) - c,
// zero beyond the end of the stream; rows are (2B+2) floats apart (in-place R2C).
template <typename T>
__global__ void __launch_bounds__(256)
k_gather_blocks(const T* __restrict__ x, int64_t n, const double2* __restrict__ pfx, int B, float* __restrict__ rows,
                int64_t k_first, int chunks_per_row) {
    const float c = sizeof(T) == 1 ? (float)rint(pfx[n].x / (double)n) : (float)(pfx[n].x / (double)n);   // = centre_of<T>
    const int64_t row = blockIdx.x / chunks_per_row;
    const int chunk = blockIdx.x % chunks_per_row;
    const int64_t k = k_first + row;
    float* out = rows + row * (int64_t)(2 * B + 2);
    const int i0 = chunk * 2048 + threadIdx.x * 2;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int i = i0 + r * 512;
        if (i < 2 * B) {
            int64_t j = k * B + i;
            float2 v;
            v.x = j < n ? (float)x[j] - c : 0.f;
            v.y = j + 1 < n ? (float)x[j + 1] - c : 0.f;
            *reinterpret_cast<float2*>(out + i) = v;
        }
    }
}

int build_prefix_u8(sb_stream* s) {
    Ctx& c = ctx();
    const int64_t n = s->n;
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    double* d_t = nullptr;
    SB_TRY(pool_alloc((void**)&d_t, sizeof(double) * 2 * ntiles));
    double* tsum = d_t; double* tsq = d_t + ntiles;
    const uint8_t* x = static_cast<const uint8_t*>(s->d_raw);
    {
        ProfScope ps("scan_tile_totals");
        k_tile_totals_u8<<<(unsigned)ntiles, SCAN_THREADS, 0, c.stream>>>(x, n, tsum, tsq);
    }
    {
        ProfScope ps("scan_tile_offsets");
        k_scan_tile_totals<<<1, 1024, 0, c.stream>>>(tsum, tsq, ntiles);
    }
    {
        ProfScope ps("scan_tiles");
        k_tile_scan_u8<<<(unsigned)ntiles, SCAN_THREADS, 0, c.stream>>>(x, n, tsum, tsq, s->d_pfx);
    }
    SB_CUDA(cudaGetLastError());
    pool_free(d_t);            // reused only by later work on the same stream
    return SB_OK;
}

template <typename T>
int build_prefix(sb_stream* s) {
    Ctx& c = ctx();
    const int64_t n = s->n;
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    double* d_t = nullptr;
    SB_TRY(pool_alloc((void**)&d_t, sizeof(double) * 2 * ntiles));
    double* tsum = d_t; double* tsq = d_t + ntiles;
    const T* x = static_cast<const T*>(s->d_raw);
    {
        ProfScope ps("scan_tile_totals");
        k_tile_totals<T><<<(unsigned)ntiles, SCAN_THREADS, 0, c.stream>>>(x, n, tsum, tsq);
    }
    {
        ProfScope ps("scan_tile_offsets");
        k_scan_tile_totals<<<1, 1024, 0, c.stream>>>(tsum, tsq, ntiles);
    }
    {
        ProfScope ps("scan_tiles");
        const size_t smem = sizeof(double2) * (SCAN_TILE + SCAN_TILE / 16);
        static bool attr_set = false;
        if (!attr_set) {
            SB_CUDA(cudaFuncSetAttribute(k_tile_scan<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        k_tile_scan<T><<<(unsigned)ntiles, SCAN_THREADS, smem, c.stream>>>(x, n, tsum, tsq, s->d_pfx);
    }
    SB_CUDA(cudaGetLastError());
    pool_free(d_t);            // reused only by later work on the same stream
    return SB_OK;
}

int stream_finish(sb_stream* s) {
    SB_TRY(pool_alloc((void**)&s->d_pfx, sizeof(double2) * (s->n + 1)));
    if (s->dtype == SB_U8) return build_prefix_u8(s);
    return build_prefix<float>(s);
}

int stream_alloc(int64_t n, int dtype, sb_stream** out, const char* who) {
    Ctx& c = ctx();
    if (!c.inited) SB_FAIL(SB_ESTATE, "%s: library not initialised (call sb_init)", who);
    if (!out) SB_FAIL(SB_EINVAL, "%s: NULL output handle", who);
    if (n < 1) SB_FAIL(SB_EINVAL, "%s: stream length %lld < 1", who, (long long)n);
    if (dtype != SB_U8 && dtype != SB_F32) SB_FAIL(SB_EINVAL, "%s: unknown dtype %d", who, dtype);
    sb_stream* s = new (std::nothrow) sb_stream();
    if (!s) SB_FAIL(SB_ENOMEM, "%s: out of host memory", who);
    s->n = n; s->dtype = dtype;
    const size_t esz = dtype == SB_U8 ? 1 : 4;
    if (pool_alloc(&s->d_raw, esz * n + 16) != SB_OK) { delete s; return SB_ENOMEM; }
    *out = s;
    return SB_OK;
}

}  // namespace

namespace sb {

int stream_finish_public(sb_stream* s) { return stream_finish(s); }

// Block spectra in the quad layout of the packed kernels (sb_fused2.cu): B = 16384, hop B.
// Rows are built on demand: [k_lo, k_hi) is what the batch at hand reads.  A rank of an event-sharded job
// touches only the part of the destination stream its own events' search windows cover, so it transforms
// that part only (the replicated stream preparation is the Amdahl term of strong scaling).
int ensure_spectra_quad(sb_stream* s, int64_t k_lo, int64_t k_hi) {
    Ctx& c = ctx();
    if (!packed_supports(c.B)) SB_FAIL(SB_EINVAL, "internal: quad-layout spectra need a lag block of 16384");
    const int64_t nblk = (s->n + c.B - 1) / c.B;
    if (k_lo < 0) k_lo = 0;
    if (k_hi > nblk) k_hi = nblk;
    if (k_lo >= k_hi) return SB_OK;
    if (!s->d_specq) {
        SB_TRY(pool_alloc((void**)&s->d_specq, sizeof(float2) * (size_t)nblk * kQuadRowF2));
        s->nblkq = nblk; s->specq_lo = s->specq_hi = k_lo;
    }
    auto build = [&](int64_t a, int64_t b) -> int {
        if (a >= b) return SB_OK;
        ProfScope ps("block_spectra");
        return launch_block_spectra_quad(s, a, b - a, s->d_specq + a * (int64_t)kQuadRowF2);
    };
    if (s->specq_lo == s->specq_hi) {                // nothing built yet
        SB_TRY(build(k_lo, k_hi));
        s->specq_lo = k_lo; s->specq_hi = k_hi;
        return SB_OK;
    }
    // keep the built range contiguous: extend it on either side (a gap between an old and a new range is filled)
    if (k_lo < s->specq_lo) { SB_TRY(build(k_lo, s->specq_lo)); s->specq_lo = k_lo; }
    if (k_hi > s->specq_hi) { SB_TRY(build(s->specq_hi, k_hi)); s->specq_hi = k_hi; }
    return SB_OK;
}

// Build (or fetch) the block spectra of `s`: row k = FFT_2B of samples [k*H, k*H + 2B), H = B/hd.
int ensure_spectra(sb_stream* s, int hd) {
    Ctx& c = ctx();
    // classic row layout ([B+1] complex per row): engine 1's kernels (also behind engines 2 / 3 for the
    // queries routed through the blocked multiply) or the cuFFT pipeline
    const int eng = (c.engine >= 1 && fused_supports(c.B)) ? 1 : 0;
    if (s->d_spec && s->specB == c.B && s->specHD == hd && s->specEngine == eng) return SB_OK;
    if (s->d_spec) { pool_free(s->d_spec); s->d_spec = nullptr; }
    const int B = c.B, H = B / hd;
    const int64_t nblk = (s->n + H - 1) / H;
    SB_TRY(pool_alloc((void**)&s->d_spec, sizeof(float2) * (size_t)nblk * (B + 1)));
    if (c.engine >= 1 && fused_supports(B)) {        // hand-written gather + forward FFT, one launch
        {
            ProfScope ps("block_spectra");
            SB_TRY(launch_block_spectra(s, hd, 0, nblk, s->d_spec));
        }
        s->specB = B; s->specHD = hd; s->nblk = nblk; s->specEngine = 1;
        return SB_OK;
    }
    if (hd != 1) SB_FAIL(SB_EINVAL, "internal: the cuFFT engine only builds spectra at hop B");
    const int chunks = (2 * B + 2047) / 2048;
    const int64_t sub = 1024;                        // rows per cuFFT call
    for (int64_t k = 0; k < nblk; k += sub) {
        const int64_t rows = (nblk - k) < sub ? (nblk - k) : sub;
        float* dst = reinterpret_cast<float*>(s->d_spec + k * (B + 1));
        {
            ProfScope ps("gather_blocks");
            if (s->dtype == SB_U8)
                k_gather_blocks<uint8_t><<<(unsigned)(rows * chunks), 256, 0, c.stream>>>(
                    static_cast<const uint8_t*>(s->d_raw), s->n, s->d_pfx, B, dst, k, chunks);
            else
                k_gather_blocks<float><<<(unsigned)(rows * chunks), 256, 0, c.stream>>>(
                    static_cast<const float*>(s->d_raw), s->n, s->d_pfx, B, dst, k, chunks);
        }
        cufftHandle plan;
        SB_TRY(get_plan(CUFFT_R2C, rows, &plan));
        {
            ProfScope ps("cufft_r2c_blocks", 0);
            SB_CUFFT(cufftExecR2C(plan, dst, reinterpret_cast<cufftComplex*>(dst)));
        }
    }
    SB_CUDA(cudaGetLastError());
    s->specB = B; s->specHD = 1; s->nblk = nblk; s->specEngine = 0;
    return SB_OK;
}

}  // namespace sb

extern "C" {

int sb_stream_create(const void* host_samples, int64_t n, int dtype, sb_stream** out) {
    if (!host_samples) SB_FAIL(SB_EINVAL, "sb_stream_create: NULL samples");
    SB_TRY(stream_alloc(n, dtype, out, "sb_stream_create"));
    sb_stream* s = *out;
    Ctx& c = ctx();
    const size_t esz = dtype == SB_U8 ? 1 : 4;
    cudaError_t e = cudaMemcpyAsync(s->d_raw, host_samples, esz * n, cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) { sb_stream_destroy(s); *out = nullptr; SB_FAIL(SB_ECUDA, "sb_stream_create: H2D copy: %s", cudaGetErrorString(e)); }
    int rc = stream_finish(s);
    if (rc != SB_OK) { sb_stream_destroy(s); *out = nullptr; }
    return rc;
}

int sb_stream_create_device(const void* dev_samples, int64_t n, int dtype, sb_stream** out) {
    if (!dev_samples) SB_FAIL(SB_EINVAL, "sb_stream_create_device: NULL samples");
    SB_TRY(stream_alloc(n, dtype, out, "sb_stream_create_device"));
    sb_stream* s = *out;
    Ctx& c = ctx();
    const size_t esz = dtype == SB_U8 ? 1 : 4;
    // enqueue only: the caller may have ordered the library stream behind a broadcast that is still in flight
    cudaError_t e = cudaMemcpyAsync(s->d_raw, dev_samples, esz * n, cudaMemcpyDeviceToDevice, c.stream);
    if (e != cudaSuccess) { sb_stream_destroy(s); *out = nullptr; SB_FAIL(SB_ECUDA, "sb_stream_create_device: D2D copy: %s", cudaGetErrorString(e)); }
    int rc = stream_finish(s);
    if (rc != SB_OK) { sb_stream_destroy(s); *out = nullptr; }
    return rc;
}

int sb_stream_destroy(sb_stream* s) {
    if (!s) return SB_OK;
    Ctx& c = ctx();
    (void)c;
    pool_free(s->d_raw); pool_free(s->d_pfx); pool_free(s->d_spec); pool_free(s->d_specq); pool_free(s->d_loadhist);
    delete s;
    return SB_OK;
}

const void* sb_stream_device_ptr(const sb_stream* s) { return s ? s->d_raw : nullptr; }
int64_t sb_stream_length(const sb_stream* s) { return s ? s->n : -1; }
int sb_stream_dtype(const sb_stream* s) { return s ? s->dtype : -1; }

int sb_stream_read(const sb_stream* s, int64_t off, int64_t n, void* host_out) {
    Ctx& c = ctx();
    if (!c.inited) SB_FAIL(SB_ESTATE, "sb_stream_read: library not initialised");
    if (!s || !host_out) SB_FAIL(SB_EINVAL, "sb_stream_read: NULL argument");
    if (off < 0 || n < 0 || off + n > s->n) SB_FAIL(SB_EINVAL, "sb_stream_read: range [%lld,+%lld) outside stream of %lld",
                                                     (long long)off, (long long)n, (long long)s->n);
    const size_t esz = s->dtype == SB_U8 ? 1 : 4;
    SB_CUDA(cudaMemcpyAsync(host_out, static_cast<const char*>(s->d_raw) + esz * off, esz * n,
                            cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    return SB_OK;
}

}  // extern "C"
