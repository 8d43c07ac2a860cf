// Loader kernels: the arithmetic of WavStream.__init__ (reference wav.py:64-91,108-156).
//   K1  sb_load_pcm   int16/int24 decode, channel average, per-chunk nearest-neighbour resample
//                     (cv2.resize INTER_NEAREST index map), edge padding      wav.py:64-91,125-141
//   K2  sb_normalise  3 x median clip, rescale to [0,1], optional uint8 quantisation  wav.py:145-156
// Everything is float32 with explicitly rounded operations (no FMA contraction), in the order the
// reference applies them, so the result is bit-identical to the NumPy/OpenCV loader.
#include "sb_internal.h"
#include <vector>
#include <cstring>

using namespace sb;

namespace sb { int stream_finish_public(sb_stream* s); }

namespace {

struct ResampleGeom {
    int64_t frames;        // frames in the file
    int framerate;         // frames per chunk (READ_CHUNK_SIZE = 1 s, wav.py:105,126)
    int64_t nfull;         // number of full chunks
    int len_last;          // frames in the trailing partial chunk (0 if none)
    int out_full;          // output samples of a full chunk   (wav.py:127)
    int out_last;          // output samples of the partial chunk
    int resample;          // downsample_rate != 1 (wav.py:131)
    double ifx_full;       // 1 / (out_full / framerate)   -- OpenCV's inverse scale, fp64
    double ifx_last;
    int64_t padding;       // wav.py:120
    int64_t total;         // wav.py:119
    int64_t written;       // samples produced by the chunk loop
};

__device__ __forceinline__ float decode_frame(const unsigned char* __restrict__ pcm, int64_t frame,
                                              int channels, int width) {
    const unsigned char* p = pcm + frame * (int64_t)channels * width;
    float acc = 0.f;
    for (int c = 0; c < channels; ++c) {
        const unsigned char* q = p + c * width + (width == 3 ? 1 : 0);     // int24: bytes 1,2 (wav.py:71-74)
        const short v = (short)((unsigned short)q[0] | ((unsigned short)q[1] << 8));
        acc = c == 0 ? (float)v : __fadd_rn(acc, (float)v);                // left-to-right float32 sum (wav.py:88-89)
    }
    return channels == 1 ? acc : __fdiv_rn(acc, (float)channels);          // wav.py:90
}

// content sample o (o in [0, written)) of the resampled stream
__device__ __forceinline__ float content_sample(const unsigned char* __restrict__ pcm, const ResampleGeom g,
                                                int64_t o, int channels, int width) {
    int64_t c; int x, len; double ifx;
    if (g.out_full > 0 && o < g.nfull * (int64_t)g.out_full) {
        c = o / g.out_full; x = (int)(o - c * g.out_full); len = g.framerate; ifx = g.ifx_full;
    } else {
        c = g.nfull; x = (int)(o - g.nfull * (int64_t)g.out_full); len = g.len_last; ifx = g.ifx_last;
    }
    int sx = x;
    if (g.resample) {
        sx = (int)floor((double)x * ifx);                                    // OpenCV resizeNN: cvFloor(x * ifx)
        if (sx > len - 1) sx = len - 1;
    }
    return decode_frame(pcm, c * (int64_t)g.framerate + sx, channels, width);
}

__global__ void __launch_bounds__(256)
k_decode_resample_pad(const unsigned char* __restrict__ pcm, ResampleGeom g, int channels, int width,
                      float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.total) return;
    const int64_t tail0 = g.total - g.padding;          // first sample of the tail padding
    float v;
    if (i < g.padding) {                                 // head: repeat the first content sample (wav.py:140)
        v = g.written > 0 ? content_sample(pcm, g, 0, channels, width) : 0.f;
    } else if (i < tail0) {
        const int64_t o = i - g.padding;
        v = o < g.written ? content_sample(pcm, g, o, channels, width) : 0.f;   // gap: np.empty -> 0
    } else {                                             // tail: repeat data[-padding-1] (wav.py:141)
        const int64_t o = tail0 - 1 - g.padding;
        v = (o >= 0 && o < g.written) ? content_sample(pcm, g, o, channels, width) : 0.f;
    }
    out[i] = v;
}

// ---- medians by radix select ----------------------------------------------------------------
__device__ __forceinline__ unsigned int float_key(float f) {      // order-preserving map float -> uint
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
float key_float(unsigned int k) {                                  // host-side inverse of float_key
    const unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f; memcpy(&f, &b, sizeof(f)); return f;
}
// subset 0: x >= 0, subset 1: x <= 0  (both contain the zeros, wav.py:145-146)
__device__ __forceinline__ bool in_subset(float f, int subset) { return subset == 0 ? f >= 0.f : f <= 0.f; }

// histogram of byte `shift/8` of the keys of subset members whose higher bytes equal `prefix`
__global__ void __launch_bounds__(256)
k_select_hist(const float* __restrict__ x, int64_t n, int subset, unsigned int prefix, unsigned int mask, int shift,
              unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float f = x[i];
        if (in_subset(f, subset)) {
            const unsigned int k = float_key(f);
            if ((k & mask) == prefix) atomicAdd(&s_h[(k >> shift) & 0xffu], 1u);
        }
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)s_h[threadIdx.x]);
}

// smallest key strictly greater than `key` among subset members
__global__ void __launch_bounds__(256)
k_select_next(const float* __restrict__ x, int64_t n, int subset, unsigned int key, unsigned int* __restrict__ out) {
    unsigned int best = 0xffffffffu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float f = x[i];
        if (in_subset(f, subset)) {
            const unsigned int k = float_key(f);
            if (k > key && k < best) best = k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned int t = __shfl_xor_sync(0xffffffffu, best, o); best = t < best ? t : best; }
    if ((threadIdx.x & 31) == 0 && best != 0xffffffffu) atomicMin(out, best);
}

__global__ void __launch_bounds__(256)
k_normalise(const float* __restrict__ x, int64_t n, float lo, float hi, float* __restrict__ out_f32,
            unsigned char* __restrict__ out_u8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    v = fminf(fmaxf(v, lo), hi);                         // np.clip (wav.py:148)
    v = __fsub_rn(v, lo);                                // wav.py:150
    v = __fdiv_rn(v, __fsub_rn(hi, lo));                 // wav.py:151 (float32 difference, float32 divide)
    if (out_u8) {
        v = __fmul_rn(v, 255.0f);                        // wav.py:154
        v = __fadd_rn(v, 0.5f);                          // wav.py:155
        out_u8[i] = (unsigned char)(int)v;               // astype('uint8'): truncation (wav.py:156)
    } else {
        out_f32[i] = v;
    }
}

// median of the subset as NumPy computes it: middle element, or the float32 mean of the two middle ones
int subset_median(const float* d_x, int64_t n, int subset, unsigned long long* d_hist, unsigned int* d_next,
                  float* median_out, int64_t* count_out) {
    Ctx& c = ctx();
    const int grid = c.sm_count * 8;
    unsigned long long h[256];
    unsigned int prefix = 0, mask = 0;
    int64_t count = 0, k = 0, below = 0;      // k: rank searched; below: members with key < current prefix range
    int64_t equal = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        SB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(h), c.stream));
        {
            ProfScope ps("median_select_hist");
            k_select_hist<<<grid, 256, 0, c.stream>>>(d_x, n, subset, prefix, mask, shift, d_hist);
        }
        SB_CUDA(cudaMemcpyAsync(h, d_hist, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
        SB_CUDA(cudaStreamSynchronize(c.stream));
        if (pass == 0) {
            for (int b = 0; b < 256; ++b) count += (int64_t)h[b];
            if (count == 0) { *median_out = nanf(""); *count_out = 0; return SB_OK; }   // np.median([]) is nan
            k = (count - 1) / 2;                 // lower middle element
        }
        int64_t acc = below;
        int b = 0;
        for (; b < 256; ++b) {
            if (acc + (int64_t)h[b] > k) break;
            acc += (int64_t)h[b];
        }
        below = acc; equal = (int64_t)h[b];
        prefix |= (unsigned int)b << shift;
        mask |= 0xffu << shift;
    }
    const unsigned int key_lo = prefix;          // key of the element of rank k
    float med = key_float(key_lo);
    if (count % 2 == 0) {                        // need rank k+1 as well
        unsigned int key_hi = key_lo;
        if (below + equal <= k + 1) {            // rank k+1 is the next distinct value
            const unsigned int init = 0xffffffffu;
            SB_CUDA(cudaMemcpyAsync(d_next, &init, sizeof(init), cudaMemcpyHostToDevice, c.stream));
            {
                ProfScope ps("median_select_next");
                k_select_next<<<grid, 256, 0, c.stream>>>(d_x, n, subset, key_lo, d_next);
            }
            SB_CUDA(cudaMemcpyAsync(&key_hi, d_next, sizeof(key_hi), cudaMemcpyDeviceToHost, c.stream));
            SB_CUDA(cudaStreamSynchronize(c.stream));
        }
        const float a = key_float(key_lo), b2 = key_float(key_hi);
        med = (a + b2) * 0.5f;                   // np.mean of two float32 values, float32 arithmetic
    }
    *median_out = med; *count_out = count;
    return SB_OK;
}

int py2_round_pos(double x) { return (int)floor(x + 0.5); }     // round() of Python 2 for x >= 0 (wav.py:127)

}  // namespace

extern "C" {

int sb_load_pcm(const void* pcm_host, int64_t frames, int channels, int sample_width,
                int framerate, int sample_rate, int64_t padding, int64_t total_len,
                sb_stream** out_f32) {
    Ctx& c = ctx();
    if (!c.inited) SB_FAIL(SB_ESTATE, "sb_load_pcm: library not initialised (call sb_init)");
    if (!pcm_host || !out_f32) SB_FAIL(SB_EINVAL, "sb_load_pcm: NULL argument");
    if (sample_width != 2 && sample_width != 3) SB_FAIL(SB_EINVAL, "Unsupported sample width: %d", sample_width);
    if (frames < 0 || channels < 1 || framerate < 1 || sample_rate < 1 || padding < 0 || total_len < 1)
        SB_FAIL(SB_EINVAL, "sb_load_pcm: bad geometry");
    ResampleGeom g;
    g.frames = frames; g.framerate = framerate;
    g.nfull = frames / framerate;
    g.len_last = (int)(frames - g.nfull * framerate);
    const double rate = (double)sample_rate / (double)framerate;          // wav.py:114
    g.resample = rate != 1.0;
    g.out_full = py2_round_pos((double)framerate * rate);
    g.out_last = py2_round_pos((double)g.len_last * rate);
    g.ifx_full = g.out_full > 0 ? 1.0 / ((double)g.out_full / (double)framerate) : 0.0;
    g.ifx_last = (g.out_last > 0 && g.len_last > 0) ? 1.0 / ((double)g.out_last / (double)g.len_last) : 0.0;
    g.padding = padding; g.total = total_len;
    g.written = g.nfull * (int64_t)g.out_full + g.out_last;
    if (g.written > total_len - padding)
        SB_FAIL(SB_EINVAL, "sb_load_pcm: %lld resampled samples do not fit a buffer of %lld with %lld padding",
                (long long)g.written, (long long)total_len, (long long)padding);

    sb_stream* s = new (std::nothrow) sb_stream();
    if (!s) SB_FAIL(SB_ENOMEM, "sb_load_pcm: out of host memory");
    s->n = total_len; s->dtype = SB_F32;
    unsigned char* d_pcm = nullptr;
    const size_t pcm_bytes = (size_t)frames * channels * sample_width;
    int rc = pool_alloc(&s->d_raw, sizeof(float) * total_len + 16);
    if (rc == SB_OK) rc = pool_alloc((void**)&d_pcm, pcm_bytes + 16);
    if (rc != SB_OK) { pool_free(s->d_raw); delete s; return rc; }
    cudaError_t e = cudaMemcpyAsync(d_pcm, pcm_host, pcm_bytes, cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) {
        ProfScope ps("decode_resample_pad");
        k_decode_resample_pad<<<(unsigned)((total_len + 255) / 256), 256, 0, c.stream>>>(
            d_pcm, g, channels, sample_width, static_cast<float*>(s->d_raw));
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);            // pcm_host may be reused by the caller
    pool_free(d_pcm);
    if (e != cudaSuccess) { pool_free(s->d_raw); delete s; SB_FAIL(SB_ECUDA, "sb_load_pcm: %s", cudaGetErrorString(e)); }
    *out_f32 = s;                 // no running sums yet: only sb_normalise / sb_stream_read accept it
    return SB_OK;
}

int sb_normalise(const sb_stream* raw_f32, int dtype, sb_stream** out, float* min3_out, float* max3_out) {
    Ctx& c = ctx();
    if (!c.inited) SB_FAIL(SB_ESTATE, "sb_normalise: library not initialised (call sb_init)");
    if (!raw_f32 || !out) SB_FAIL(SB_EINVAL, "sb_normalise: NULL argument");
    if (raw_f32->dtype != SB_F32) SB_FAIL(SB_EINVAL, "sb_normalise: input must be a float32 stream from sb_load_pcm");
    if (dtype != SB_U8 && dtype != SB_F32) SB_FAIL(SB_EINVAL, "Unknown sample type of WAV stream, must be uint8 or float32");
    const int64_t n = raw_f32->n;
    const float* x = static_cast<const float*>(raw_f32->d_raw);
    unsigned long long* d_hist = nullptr; unsigned int* d_next = nullptr;
    SB_TRY(pool_alloc((void**)&d_hist, 256 * sizeof(unsigned long long)));
    int rc = pool_alloc((void**)&d_next, 256);
    float med_pos = 0.f, med_neg = 0.f; int64_t cnt = 0;
    if (rc == SB_OK) rc = subset_median(x, n, 0, d_hist, d_next, &med_pos, &cnt);
    if (rc == SB_OK) rc = subset_median(x, n, 1, d_hist, d_next, &med_neg, &cnt);
    pool_free(d_hist); pool_free(d_next);
    if (rc != SB_OK) return rc;
    const float hi = med_pos * 3.0f, lo = med_neg * 3.0f;               // wav.py:145-146 (float32 products)
    if (min3_out) *min3_out = lo;
    if (max3_out) *max3_out = hi;

    sb_stream* s = new (std::nothrow) sb_stream();
    if (!s) SB_FAIL(SB_ENOMEM, "sb_normalise: out of host memory");
    s->n = n; s->dtype = dtype;
    rc = pool_alloc(&s->d_raw, (dtype == SB_U8 ? 1 : 4) * (size_t)n + 16);
    if (rc != SB_OK) { delete s; return rc; }
    {
        ProfScope ps("normalise_quantise");
        k_normalise<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(
            x, n, lo, hi, dtype == SB_F32 ? static_cast<float*>(s->d_raw) : nullptr,
            dtype == SB_U8 ? static_cast<unsigned char*>(s->d_raw) : nullptr);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { sb_stream_destroy(s); SB_FAIL(SB_ECUDA, "sb_normalise: %s", cudaGetErrorString(e)); }
    rc = stream_finish_public(s);
    if (rc != SB_OK) { sb_stream_destroy(s); return rc; }
    *out = s;
    return SB_OK;
}

}  // extern "C"
