// Fused lag-block kernel, 16 values per thread (64 registers, 32 resident warps per SM) -- the
// occupancy-oriented sibling of sb_fused.cu; same phases, same arithmetic, four Stockham passes of
// radix 16 x 16 x 16 x {4|2} instead of three.  For one (query, lag block) item a CTA
//   1. accumulates Y[bin] = sum_p conj(T^_p[bin]) * X^_{k+p}[bin] straight from L2 into registers,
//   2. packs the Hermitian half spectrum into a half-size complex sequence in shared memory,
//   3. runs the inverse FFT entirely in shared memory (Stockham, radix 16 x 16 x 16 x {4|2}),
//   4. slides the window sums through its lags, screens every lag in fp32 and evaluates only the
//      lags that can still be the minimum with OpenCV's rule in fp64, and
//   5. merges its (value, first index) into the query's result with one 64-bit atomicMin.
// The B-long correlation block never touches HBM (the cuFFT pipeline in sb_matcher.cu writes and
// re-reads it twice); the only global traffic is the spectra (L2 resident) and 2 B per lag of raw
// samples.  See DESIGN.md section 4 for the roofline of each phase.
#include "sb_internal.h"
#include <cmath>
#include <cstdlib>
#include <vector>

#include "sb_fused_common.cuh"

using namespace sb;
using namespace sbf;

namespace {

// ---------------------------------------------------------------- configuration
template <int LOGN> struct Cfg;
template <> struct Cfg<14> {   // B = 16384 lags per item, one CTA of 1024 threads per SM
    static constexpr int N = 16384, T = 1024, R4 = 4, MINB = 1;
};
template <> struct Cfg<13> {   // B = 8192 lags per item, two CTAs of 512 threads per SM
    static constexpr int N = 8192, T = 512, R4 = 2, MINB = 2;
};

struct FusedTables {
    const float2* w;      // [N/2+1]    exp(+i*pi*m/N)                       (Hermitian unpacking)
    const float2* t2;     // [16][16]   exp(+2*pi*i*r*k/256),  k = j & 15      (pass 2)
    const float2* a3;     // [16][32]   exp(+2*pi*i*r*k/4096), k = lane        (pass 3, low part)
    const float2* b3;     // [16][8]    exp(+2*pi*i*r*kh/128), kh = (j>>5)&7   (pass 3, high part)
    const float2* a4;     // [R4][32]   exp(+2*pi*i*r*k/N),    k = lane        (pass 4, low part)
    const float2* b4;     // [R4][128]  exp(+2*pi*i*r*kh*32/N), kh = j >> 5    (pass 4, high part)
};

// one padding slot per 16 complex values keeps the radix-16 scatter of pass 1 conflict free
__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

// ---------------------------------------------------------------- the kernel
template <int LOGN, typename S>
__global__ void __launch_bounds__(Cfg<LOGN>::T, Cfg<LOGN>::MINB)
k_match_fused16(const float2* __restrict__ That, int64_t part_first,
              const float2* __restrict__ Xhat, int64_t nblk,
              const S* __restrict__ img, int64_t img_n,
              const double2* __restrict__ ipfx, const double2* __restrict__ tpfx,
              const QueryDesc* __restrict__ desc, const int* __restrict__ item_query, int64_t item_first,
              FusedTables tab, unsigned long long* __restrict__ keys, float* __restrict__ curve_out) {
    typedef Cfg<LOGN> C;
    constexpr int N = C::N, T = C::T, B = C::N;
    constexpr int NB = B + 1;                       // bins per spectrum row
    constexpr int LAGS_PER_ROUND = T * 8;           // 8 consecutive lags per thread per round
    constexpr int ROUNDS = B / LAGS_PER_ROUND;      // 4
    constexpr int NW = T / 32;
    constexpr int R4 = C::R4;
    constexpr int NT2 = 16 * 16, NA3 = 16 * 32, NB3 = 16 * 8, NA4 = R4 * 32, NB4 = R4 * 128;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);                  // pad16(N) complex values
    float2* s_t2 = buf + pad16(N) + 1;                                  // twiddle tables
    float2* s_a3 = s_t2 + NT2;
    float2* s_b3 = s_a3 + NA3;
    float2* s_a4 = s_b3 + NB3;
    float2* s_b4 = s_a4 + NA4;
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(s_b4 + NB4);   // [NW]
    float* s_min = reinterpret_cast<float*>(s_best + NW);               // [NW]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t item = item_first + blockIdx.x;
    const int q = __ldg(item_query + blockIdx.x);     // which query this lag block belongs to
    // stage the FFT twiddle tables (16 KB); first needed in pass 2, several barriers from here
    for (int i = tid; i < NT2; i += T) s_t2[i] = __ldg(tab.t2 + i);
    for (int i = tid; i < NA3; i += T) s_a3[i] = __ldg(tab.a3 + i);
    for (int i = tid; i < NB3; i += T) s_b3[i] = __ldg(tab.b3 + i);
    for (int i = tid; i < NA4; i += T) s_a4[i] = __ldg(tab.a4 + i);
    for (int i = tid; i < NB4; i += T) s_b4[i] = __ldg(tab.b4 + i);
    const QueryDesc d = desc[q];
    const int64_t k = d.k0 + (item - d.itemBase);

    // ---------------- 1+2. spectral multiply-accumulate and Hermitian packing ---------------
    {
        int P = d.P;
        if (k + P > nblk) P = (int)(nblk - k);      // blocks past the end of the stream are zero
        const float2* tp = That + (d.partBase - part_first) * (int64_t)NB;
        const float2* xp = Xhat + k * (int64_t)NB;
        constexpr int U = 2;                        // bin pairs in flight per thread
        static_assert((B / 2) % (U * T) == 0, "pair loop must tile B/2");
        // pairs (m, B-m), m = 1 .. B/2-1, plus m = 0 whose partner is the Nyquist bin B
        for (int m0 = tid; m0 < B / 2; m0 += U * T) {
            float2 ym[U], yp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { ym[u] = make_float2(0.f, 0.f); yp[u] = make_float2(0.f, 0.f); }
            for (int p = 0; p < P; ++p) {
                const float2* t = tp + (int64_t)p * NB;
                const float2* x = xp + (int64_t)p * NB;
                float2 t1[U], x1[U], t2[U], x2[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int m = m0 + u * T;
                    t1[u] = __ldg(t + m); x1[u] = __ldg(x + m);
                    t2[u] = __ldg(t + (B - m)); x2[u] = __ldg(x + (B - m));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    ym[u].x += t1[u].x * x1[u].x + t1[u].y * x1[u].y;  ym[u].y += t1[u].x * x1[u].y - t1[u].y * x1[u].x;
                    yp[u].x += t2[u].x * x2[u].x + t2[u].y * x2[u].y;  yp[u].y += t2[u].x * x2[u].y - t2[u].y * x2[u].x;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int m = m0 + u * T;
                const float2 w = __ldg(tab.w + m);                     // exp(+i*pi*m/B)
                // Z[m]   = (Ym + conj(Yp)) + i*(Ym - conj(Yp))*w
                const float2 e = make_float2(ym[u].x + yp[u].x, ym[u].y - yp[u].y);
                const float2 o = cmul(make_float2(ym[u].x - yp[u].x, ym[u].y + yp[u].y), w);
                buf[pad16(m)] = make_float2(e.x - o.y, e.y + o.x);
                // Z[B-m] = conj(e) + i*conj(o)   (w^(B-m) = -conj(w^m))
                if (m > 0) buf[pad16(B - m)] = make_float2(e.x + o.y, -e.y + o.x);
            }
        }
        if (tid < 32) {                             // the self-paired bin m = B/2 (w = i): Z = 2*conj(Y)
            float2 y = make_float2(0.f, 0.f);
            for (int p = lane; p < P; p += 32) {
                const float2 t1 = __ldg(tp + (int64_t)p * NB + B / 2), x1 = __ldg(xp + (int64_t)p * NB + B / 2);
                y.x += t1.x * x1.x + t1.y * x1.y;  y.y += t1.x * x1.y - t1.y * x1.x;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { y.x += __shfl_xor_sync(0xffffffffu, y.x, o); y.y += __shfl_xor_sync(0xffffffffu, y.y, o); }
            if (lane == 0) buf[pad16(B / 2)] = make_float2(2.f * y.x, -2.f * y.y);
        }
    }
    __syncthreads();

    // ---------------- 3. inverse FFT of N complex points in shared memory -------------------
    // Stockham passes: butterfly j reads in[j + r*N/R], twiddles by exp(2*pi*i*r*k/(Ns*R)),
    // k = j mod Ns, and writes out[(j-k)*R + k + r*Ns]; every value sits in a register between
    // the two barriers, so the pass is in place.
    {   // pass 1: radix 16, Ns = 1 (no twiddles); one butterfly per thread
        static_assert(N / 16 == T, "one radix-16 butterfly per thread");
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = buf[pad16(tid + r * T)];
        __syncthreads();
        dft_dif<16>(v);
#pragma unroll
        for (int r = 0; r < 16; ++r) buf[pad16(tid * 16 + r)] = v[brev<16>(r)];
        __syncthreads();
    }
    {   // pass 2: radix 16, Ns = 16, k = j & 15
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = buf[pad16(tid + r * T)];
        __syncthreads();
        const int k = tid & 15;
#pragma unroll
        for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], s_t2[r * 16 + k]);
        dft_dif<16>(v);
        const int j0 = (tid - k) * 16 + k;
#pragma unroll
        for (int r = 0; r < 16; ++r) buf[pad16(j0 + r * 16)] = v[brev<16>(r)];
        __syncthreads();
    }
    {   // pass 3: radix 16, Ns = 256, k = j & 255 = kh * 32 + lane
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = buf[pad16(tid + r * T)];
        __syncthreads();
        const int k = tid & 255, kh = k >> 5;
#pragma unroll
        for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], cmul(s_a3[r * 32 + lane], s_b3[r * 8 + kh]));
        dft_dif<16>(v);
        const int j0 = (tid - k) * 16 + k;
#pragma unroll
        for (int r = 0; r < 16; ++r) buf[pad16(j0 + r * 256)] = v[brev<16>(r)];
        __syncthreads();
    }
    {   // pass 4: radix R4, Ns = 4096, k = j (j < 4096); only the first half of the outputs is
        // needed: z[0 .. N/2) carries the B valid lags of the 2B-point real sequence
        constexpr int Ns = 4096, PER = Ns / T;
        static_assert(N / R4 == Ns, "pass 4 is the last pass");
        float2 v[PER][R4];
#pragma unroll
        for (int b = 0; b < PER; ++b)
#pragma unroll
            for (int r = 0; r < R4; ++r) v[b][r] = buf[pad16(tid + b * T + r * Ns)];
        __syncthreads();
#pragma unroll
        for (int b = 0; b < PER; ++b) {
            const int j = tid + b * T;
            const int kh = j >> 5;
#pragma unroll
            for (int r = 1; r < R4; ++r) v[b][r] = cmul(v[b][r], cmul(s_a4[r * 32 + lane], s_b4[r * 128 + kh]));
            dft_dif<R4, true>(v[b]);
#pragma unroll
            for (int r = 0; r < R4 / 2; ++r) buf[pad16(j + r * Ns)] = v[b][brev<R4>(r)];
        }
        __syncthreads();
    }
    // now buf[pad16(i)] = (x[2i], x[2i+1]) for i < N/2: correlation at lags 2i, 2i+1 (times 2B)

    // ---------------- 4. window sums, fp32 screening, fp64 exact evaluation -----------------
    // Every thread owns runs of 8 consecutive lags.  The exact window sums at the first lag of a
    // run come from the interleaved fp64 running sums (two 16-byte loads per run); inside the run
    // the window slides on the raw samples: W[j+1] = W[j] + I[j+n]^k - I[j]^k.
    const int64_t n = d.tlen;
    const int64_t jlo = d.lag0, jhi = d.lag0 + d.nlags;
    const int64_t j_blk = k * B;
    const double2 t_hi = tpfx[d.toff + n], t_lo = tpfx[d.toff];
    const double tsum = t_hi.x - t_lo.x, tsq = t_hi.y - t_lo.y;
    const double a = (double)Acc<S>::centre(ipfx[img_n].x, (double)img_n);
    const double b = (double)Acc<S>::centre(tsum, (double)n);
    const double n_ab = (double)n * a * b;
    const double scale = 1.0 / (double)(2 * B);
    const double k_const = a * tsum - n_ab;
    const float f_tsq = (float)tsq, f_b = (float)b, f_scale = (float)scale;
    const bool interior = j_blk >= jlo && j_blk + B <= jhi;           // every lag of the item is valid

    // exact window sums at the head of each run, fetched up front (two 16-byte loads per run)
    double w0s[ROUNDS], w0q[ROUNDS];
    bool live[ROUNDS];
#pragma unroll
    for (int c = 0; c < ROUNDS; ++c) {
        const int64_t j0 = j_blk + c * LAGS_PER_ROUND + tid * 8;
        live[c] = j0 < jhi && j0 + 8 > jlo;                           // the run holds a valid lag
        w0s[c] = 0.0; w0q[c] = 0.0;
        if (live[c]) {
            const double2 p_hi = ipfx[j0 + n], p_lo = ipfx[j0];
            w0s[c] = p_hi.x - p_lo.x; w0q[c] = p_hi.y - p_lo.y;
        }
    }

    float vf[ROUNDS][8];
    float tmin = 2.0f;
#pragma unroll
    for (int c = 0; c < ROUNDS; ++c) {
        const int m0 = c * LAGS_PER_ROUND + tid * 8;
        const int64_t j0 = j_blk + m0;
#pragma unroll
        for (int i = 0; i < 8; ++i) vf[c][i] = 2.0f;                  // sentinel: not a valid lag
        if (live[c]) {
            const float f_w0q = (float)w0q[c];
            const float f_k0 = (float)(b * w0s[c] + k_const);
            float cc[8];
#pragma unroll
            for (int h = 0; h < 4; ++h) { const float2 z = buf[pad16((m0 >> 1) + h)]; cc[2 * h] = z.x; cc[2 * h + 1] = z.y; }
            if (sizeof(S) == 1) {
                const uint8_t* img8 = reinterpret_cast<const uint8_t*>(img);
                const unsigned long long lo8 = __ldg(reinterpret_cast<const unsigned long long*>(img8 + j0));   // j0 % 8 == 0
                const unsigned long long hi8 = load8(img8, j0 + n);
                int rq = 0, rs = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float wq = f_w0q + (float)rq;
                    const float sit = fmaf(cc[i], f_scale, fmaf(f_b, (float)rs, f_k0));
                    const float num = fmaxf((wq + f_tsq) - 2.0f * sit, 0.0f);
                    const float pr = wq * f_tsq;
                    const float v = pr > 0.0f ? fminf(num * rsqrtf(pr), 1.0f) : 1.0f;
                    if (interior || (j0 + i >= jlo && j0 + i < jhi)) { vf[c][i] = v; tmin = fminf(tmin, v); }
                    const int lo = (int)((lo8 >> (8 * i)) & 0xffu), hi = (int)((hi8 >> (8 * i)) & 0xffu);
                    rq += hi * hi - lo * lo; rs += hi - lo;
                }
            } else {
                double rq = 0.0, rs = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t j = j0 + i;
                    const float wq = f_w0q + (float)rq;
                    const float sit = fmaf(cc[i], f_scale, fmaf(f_b, (float)rs, f_k0));
                    const float num = fmaxf((wq + f_tsq) - 2.0f * sit, 0.0f);
                    const float pr = wq * f_tsq;
                    const float v = pr > 0.0f ? fminf(num * rsqrtf(pr), 1.0f) : 1.0f;
                    if (j >= jlo && j < jhi) { vf[c][i] = v; tmin = fminf(tmin, v); }
                    if (j + n < img_n) {
                        const double lo = (double)img[j], hi = (double)img[j + n];
                        rq += hi * hi - lo * lo; rs += hi - lo;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
    if (lane == 0) s_min[warp] = tmin;
    __syncthreads();
    float bmin = s_min[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) bmin = fminf(bmin, s_min[w]);
    const float thr = curve_out ? 1.5f : bmin + kScreenMargin;       // debug curve: evaluate everything

    // lags that can still be the minimum (bit c*8+i), then ONE copy of the fp64 path
    unsigned cand = 0;
#pragma unroll
    for (int c = 0; c < ROUNDS; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) cand |= (vf[c][i] <= thr) ? (1u << (c * 8 + i)) : 0u;
    unsigned long long best = ~0ull;
    while (cand) {
        const int bit = __ffs(cand) - 1;
        cand &= cand - 1;
        const int m = (bit >> 3) * LAGS_PER_ROUND + tid * 8 + (bit & 7);
        const int64_t j = j_blk + m;
        const float2 z = buf[pad16(m >> 1)];
        const double cc = (double)((m & 1) ? z.y : z.x) * scale;
        const double2 p_hi = ipfx[j + n], p_lo = ipfx[j];
        const float v = sqdiff_exact(cc, p_hi.x - p_lo.x, p_hi.y - p_lo.y, a, b, tsum, tsq, n_ab);
        if (curve_out) curve_out[d.curveOff + (j - jlo)] = v;
        const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(j - jlo);
        best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NW; ++w) best = s_best[w] < best ? s_best[w] : best;
        if (best != ~0ull) atomicMin(keys + q, best);
    }
}

// ---------------------------------------------------------------- host side
template <int LOGN> size_t fused_smem_bytes() {
    typedef Cfg<LOGN> C;
    const size_t padded = (size_t)(C::N + (C::N >> 4) + 1);
    const size_t tables = 16 * 16 + 16 * 32 + 16 * 8 + C::R4 * 32 + C::R4 * 128;
    const size_t nw = C::T / 32;
    return (padded + tables) * sizeof(float2) + nw * sizeof(unsigned long long) + nw * sizeof(float) + 64;
}

struct TableSet { float2* dev = nullptr; FusedTables tab; };
TableSet g_tables[2];     // [0]: LOGN 13, [1]: LOGN 14

template <int LOGN> int ensure_tables(FusedTables* out) {
    typedef Cfg<LOGN> C;
    TableSet& ts = g_tables[LOGN - 13];
    if (!ts.dev) {
        const int N = C::N, R4 = C::R4;
        const size_t nw = N / 2 + 1;
        std::vector<float2> h;
        const double pi = 3.14159265358979323846;
        auto push = [&](double ang) { h.push_back(make_float2((float)cos(ang), (float)sin(ang))); };
        for (size_t m = 0; m < nw; ++m) push(pi * m / N);
        const size_t o_t2 = h.size();
        for (int r = 0; r < 16; ++r) for (int k = 0; k < 16; ++k) push(2.0 * pi * r * k / 256.0);
        const size_t o_a3 = h.size();
        for (int r = 0; r < 16; ++r) for (int k = 0; k < 32; ++k) push(2.0 * pi * r * k / 4096.0);
        const size_t o_b3 = h.size();
        for (int r = 0; r < 16; ++r) for (int k = 0; k < 8; ++k) push(2.0 * pi * r * k / 128.0);
        const size_t o_a4 = h.size();
        for (int r = 0; r < R4; ++r) for (int k = 0; k < 32; ++k) push(2.0 * pi * r * k / N);
        const size_t o_b4 = h.size();
        for (int r = 0; r < R4; ++r) for (int k = 0; k < 128; ++k) push(2.0 * pi * r * (k * 32.0) / N);
        SB_CUDA(cudaMalloc(&ts.dev, h.size() * sizeof(float2)));
        SB_CUDA(cudaMemcpy(ts.dev, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
        ts.tab.w = ts.dev; ts.tab.t2 = ts.dev + o_t2; ts.tab.a3 = ts.dev + o_a3; ts.tab.b3 = ts.dev + o_b3;
        ts.tab.a4 = ts.dev + o_a4; ts.tab.b4 = ts.dev + o_b4;
    }
    *out = ts.tab;
    return SB_OK;
}

// item_query[i] = query of item (item_first + i): one CTA per query fills its own range
__global__ void k_fill_item_query16(const QueryDesc* __restrict__ desc, int q_begin, int64_t item_first,
                                  int* __restrict__ item_query) {
    const int q = q_begin + blockIdx.x;
    const int64_t base = desc[q].itemBase - item_first;
    const int nk = desc[q].nk;
    for (int i = threadIdx.x; i < nk; i += blockDim.x) item_query[base + i] = q;
}

int* g_item_query16 = nullptr;
int64_t g_item_query16_cap = 0;

template <int LOGN, typename S>
int launch_typed(const sb_stream* image, const sb_stream* tmpl, const float2* d_parts, int64_t part_first,
                 const QueryDesc* d_desc, int q_begin, int q_end, int64_t item_first, int64_t n_items,
                 unsigned long long* d_keys, float* d_curve) {
    Ctx& c = ctx();
    FusedTables tab;
    SB_TRY(ensure_tables<LOGN>(&tab));
    static bool attr_set = false;
    // SB_FUSED_EXTRA_SMEM (bytes): occupancy experiments only -- inflates the dynamic shared memory so
    // fewer CTAs fit on an SM
    static const size_t extra = getenv("SB_FUSED_EXTRA_SMEM") ? (size_t)atol(getenv("SB_FUSED_EXTRA_SMEM")) : 0;
    const size_t smem = fused_smem_bytes<LOGN>() + extra;
    if (!attr_set) {
        SB_CUDA(cudaFuncSetAttribute(k_match_fused16<LOGN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    if (g_item_query16_cap < n_items) {
        cudaStreamSynchronize(c.stream);
        cudaFree(g_item_query16); g_item_query16 = nullptr; g_item_query16_cap = 0;
        SB_CUDA(cudaMalloc(&g_item_query16, sizeof(int) * (size_t)n_items));
        g_item_query16_cap = n_items;
    }
    k_fill_item_query16<<<(unsigned)(q_end - q_begin), 128, 0, c.stream>>>(d_desc, q_begin, item_first, g_item_query16);
    c.launches += 1;
    const int64_t max_grid = 1 << 30;
    for (int64_t i0 = 0; i0 < n_items; i0 += max_grid) {
        const int64_t ni = std::min<int64_t>(max_grid, n_items - i0);
        k_match_fused16<LOGN, S><<<(unsigned)ni, Cfg<LOGN>::T, smem, c.stream>>>(
            d_parts, part_first, image->d_spec, image->nblk, static_cast<const S*>(image->d_raw), image->n,
            image->d_pfx, tmpl->d_pfx, d_desc, g_item_query16 + i0, item_first + i0,
            tab, d_keys, d_curve);
    }
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

}  // namespace

namespace sb {

int launch_match_fused16(const sb_stream* image, const sb_stream* tmpl, const float2* d_parts, int64_t part_first,
                       const QueryDesc* d_desc, int q_begin, int q_end, int64_t item_first, int64_t n_items,
                       unsigned long long* d_keys, float* d_curve) {
    const int B = ctx().B;
    const bool u8 = image->dtype == SB_U8;
    if (B == 16384)
        return u8 ? launch_typed<14, uint8_t>(image, tmpl, d_parts, part_first, d_desc, q_begin, q_end, item_first, n_items, d_keys, d_curve)
                  : launch_typed<14, float>(image, tmpl, d_parts, part_first, d_desc, q_begin, q_end, item_first, n_items, d_keys, d_curve);
    if (B == 8192)
        return u8 ? launch_typed<13, uint8_t>(image, tmpl, d_parts, part_first, d_desc, q_begin, q_end, item_first, n_items, d_keys, d_curve)
                  : launch_typed<13, float>(image, tmpl, d_parts, part_first, d_desc, q_begin, q_end, item_first, n_items, d_keys, d_curve);
    SB_FAIL(SB_EINVAL, "fused engine supports lag blocks of 8192 or 16384 samples, not %d", B);
}

void fused16_release_tables() {
    for (auto& t : g_tables) { if (t.dev) cudaFree(t.dev); t.dev = nullptr; }
    cudaFree(g_item_query16); g_item_query16 = nullptr; g_item_query16_cap = 0;
}

}  // namespace sb
