// Log-mel filterbank features on the GPU: the front of the extraction path.
// Replaces scripts/featureExtractor.py:8-30 (`mfsc` + `normalize`) = librosa 0.7.2 `stft(center=False)` + `filters.mel`
// (a third-party dependency that is not in the reference tree; its published algorithm is restated in
// oracle/feature_oracle.py and cross-checked there against an independent implementation):
//   y *= 32768; pre-emphasis over the whole signal (first sample scaled by 1 - c)            featureExtractor.py:16-18
//   frames of n_fft = 512 samples every `hop`, window of win_length taps centred in the frame  :19 (librosa.stft)
//   |rfft|, mel = melw[n_mels, 257] . |S|                                                        :20-21
//   log(max(1, mel))                                                                             :22
//   cepstral mean normalisation over the utterance's frames                                      :25-26
//
// One warp per frame.  The 512-point real FFT is a 256-point complex FFT (radix-2, in shared memory, bit-reversed
// load) plus the usual split step; ~10 kFLOP per frame instead of 263 kFLOP for the DFT as a matrix product, so the
// stage is bound by nothing in particular: 102 k frames (256 utterances x 4 s) take tens of microseconds.  Roofline:
// HBM, algorithmic bytes = 4 B per sample read + 4 * n_mels B per frame written.
#include "common.cuh"
#include <math.h>

namespace dasv {

constexpr int kFftN = 512;              // the reference's n_fft (featureExtractor.py:11)
constexpr int kFftH = kFftN / 2;        // complex FFT length
constexpr int kBins = kFftH + 1;        // 257 one-sided bins
constexpr int kFeatWarps = 8;

struct LogmelParams {
    const float* wave;          // [B][stride]
    const int32_t* n_samples;   // [B] valid samples per utterance
    long long stride;
    const float* window;        // [win_length]
    const float* melw;          // [n_mels][257]
    const int32_t* mel_range;   // [n_mels][2]: first bin, one past the last bin with a non-zero weight
    float* out;                 // [B][Tmax][n_mels]
    int B, Tmax, n_mels, win_length, hop;
    float preem, scale;
};

__global__ void __launch_bounds__(kFeatWarps * 32) logmel_kernel(const LogmelParams p) {
    __shared__ float2 tw[kFftH / 2];                 // exp(-2 pi i k / 256), k < 128
    __shared__ float2 tw2[kFftH / 2 + 1];            // exp(-2 pi i k / 512), k <= 128
    __shared__ float2 z[kFeatWarps][kFftH];
    __shared__ float mag[kFeatWarps][kBins + 3];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < kFftH / 2; k += blockDim.x) {
        float s, c;
        sincospif(-static_cast<float>(k) / 128.f, &s, &c);
        tw[k] = make_float2(c, s);
    }
    for (int k = tid; k <= kFftH / 2; k += blockDim.x) {
        float s, c;
        sincospif(-static_cast<float>(k) / 256.f, &s, &c);
        tw2[k] = make_float2(c, s);
    }
    __syncthreads();

    const int b = blockIdx.y;
    const int t = blockIdx.x * kFeatWarps + warp;
    const int n = p.n_samples[b];
    const int frames = n >= kFftN ? 1 + (n - kFftN) / p.hop : 0;
    if (t >= frames) return;                         // no block-level barrier below

    const float* w = p.wave + static_cast<long long>(b) * p.stride;
    const int pad = (kFftN - p.win_length) / 2;      // librosa pad_center: window centred in the frame
    const long long s0 = static_cast<long long>(t) * p.hop;
    float2* zw = z[warp];

    // windowed, pre-emphasised samples -> z[bitrev(j)] = x[2j] + i x[2j+1]
    for (int j = lane; j < kFftH; j += 32) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * j + e;                 // position in the frame
            const int wi = i - pad;
            float x = 0.f;
            if (wi >= 0 && wi < p.win_length) {
                const long long g = s0 + i;
                const float cur = w[g];
                x = (g == 0) ? cur * (1.f - p.preem) : cur - p.preem * w[g - 1];
                x *= p.scale * p.window[wi];
            }
            v[e] = x;
        }
        zw[__brev(static_cast<unsigned>(j)) >> 24] = make_float2(v[0], v[1]);
    }
    __syncwarp();

    // 256-point complex FFT, radix-2 decimation in time, in place
#pragma unroll 1
    for (int half = 1; half < kFftH; half <<= 1) {
        const int tstep = (kFftH / 2) / half;
        for (int j = lane; j < kFftH / 2; j += 32) {
            const int pos = j & (half - 1);
            const int i0 = ((j - pos) << 1) + pos, i1 = i0 + half;
            const float2 wv = tw[pos * tstep];
            const float2 a = zw[i0], c = zw[i1];
            const float2 tt = make_float2(wv.x * c.x - wv.y * c.y, wv.x * c.y + wv.y * c.x);
            zw[i0] = make_float2(a.x + tt.x, a.y + tt.y);
            zw[i1] = make_float2(a.x - tt.x, a.y - tt.y);
        }
        __syncwarp();
    }

    // split step: X[k] = (Z[k] + conj Z[256-k]) / 2 - i/2 e^{-2 pi i k / 512} (Z[k] - conj Z[256-k]),  k = 0 .. 256
    for (int k = lane; k < kBins; k += 32) {
        const float2 a = zw[k & (kFftH - 1)];
        const float2 c = zw[(kFftH - k) & (kFftH - 1)];
        const float er = 0.5f * (a.x + c.x), ei = 0.5f * (a.y - c.y);       // even part
        const float odr = 0.5f * (a.x - c.x), odi = 0.5f * (a.y + c.y);     // (Z[k] - conj Z[256-k]) / 2
        // twiddle e^{-2 pi i k / 512}: table holds k <= 128, the rest by symmetry  w(256 - k) = -conj w(k)
        float2 wv = (k <= 128) ? tw2[k] : make_float2(-tw2[256 - k].x, tw2[256 - k].y);
        // -i * w * od
        const float pr = wv.x * odr - wv.y * odi, pi = wv.x * odi + wv.y * odr;
        const float xr = er + pi, xi = ei - pr;
        mag[warp][k] = sqrtf(xr * xr + xi * xi);
    }
    __syncwarp();

    float* o = p.out + (static_cast<size_t>(b) * p.Tmax + t) * p.n_mels;
    for (int m = lane; m < p.n_mels; m += 32) {
        const int lo = p.mel_range[2 * m], hi = p.mel_range[2 * m + 1];
        const float* mw = p.melw + static_cast<size_t>(m) * kBins;
        float acc = 0.f;
        for (int k = lo; k < hi; ++k) acc = fmaf(mw[k], mag[warp][k], acc);
        o[m] = logf(fmaxf(1.f, acc));
    }
}

// Cepstral mean (and variance) normalisation (featureExtractor.py:25-26, data.py:21-30 'cmn' / 'cmvn'): one CTA per
// utterance, thread (m, slice) owns mel bin m and every 8th frame: fixed-order Kahan sums over the utterance's frames, then
// the subtraction (and the division by the population standard deviation where it exceeds 0.01, data.py:28-29); frames
// past the utterance are zeroed.
__global__ void cmn_kernel(float* feat, const int32_t* frames, int Tmax, int n_mels, int variance) {
    const int b = blockIdx.x;
    const int nf = min(max(frames[b], 0), Tmax);
    float* f = feat + static_cast<size_t>(b) * Tmax * n_mels;
    __shared__ float part[8][128];
    __shared__ float stat[2][128];
    const int m = threadIdx.x % 128, slice = threadIdx.x / 128;      // 1024 threads = 8 time slices x 128 bins
    float s = 0.f, c = 0.f;                          // Kahan sum
    if (m < n_mels)
        for (int t = slice; t < nf; t += 8) {
            const float y = f[static_cast<size_t>(t) * n_mels + m] - c;
            const float u = s + y;
            c = (u - s) - y;
            s = u;
        }
    part[slice][m] = s;
    __syncthreads();
    if (slice == 0) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += part[i][m];
        stat[0][m] = nf > 0 ? a / static_cast<float>(nf) : 0.f;
    }
    __syncthreads();
    const float mean = stat[0][m];
    float inv = 1.f;
    if (variance) {
        s = 0.f; c = 0.f;
        if (m < n_mels)
            for (int t = slice; t < nf; t += 8) {
                const float d = f[static_cast<size_t>(t) * n_mels + m] - mean;
                const float y = d * d - c;
                const float u = s + y;
                c = (u - s) - y;
                s = u;
            }
        part[slice][m] = s;
        __syncthreads();
        if (slice == 0) {
            float a = 0.f;
            for (int i = 0; i < 8; ++i) a += part[i][m];
            const float sd = nf > 0 ? sqrtf(a / static_cast<float>(nf)) : 1.f;
            stat[1][m] = sd > 0.01f ? 1.f / sd : 1.f;
        }
        __syncthreads();
        inv = stat[1][m];
    }
    if (m < n_mels)
        for (int t = slice; t < Tmax; t += 8) {
            const size_t i = static_cast<size_t>(t) * n_mels + m;
            f[i] = t < nf ? (f[i] - mean) * inv : 0.f;
        }
}

}  // namespace dasv

using namespace dasv;

extern "C" int dasv_logmel_f32(const float* wave, const int32_t* n_samples, int B, long long wave_stride,
                               const float* window, int win_length, int hop,
                               const float* melw, const int32_t* mel_range, int n_mels,
                               float preem, float scale, float* out, int Tmax, void* stream) {
    if (B < 0 || Tmax < 0) { set_error("logmel: negative shape"); return 1; }
    if (B == 0 || Tmax == 0) return 0;
    if (!wave || !n_samples || !window || !melw || !mel_range || !out) { set_error("logmel: null pointer"); return 1; }
    if (win_length < 1 || win_length > kFftN || hop < 1) { set_error("logmel: need 1 <= win_length <= 512 and hop >= 1 (got %d, %d)", win_length, hop); return 1; }
    if (n_mels < 1 || n_mels > 128) { set_error("logmel: n_mels %d outside [1, 128]", n_mels); return 1; }
    LogmelParams p{};
    p.wave = wave; p.n_samples = n_samples; p.stride = wave_stride; p.window = window; p.melw = melw; p.mel_range = mel_range;
    p.out = out; p.B = B; p.Tmax = Tmax; p.n_mels = n_mels; p.win_length = win_length; p.hop = hop; p.preem = preem; p.scale = scale;
    dim3 grid((Tmax + kFeatWarps - 1) / kFeatWarps, B);
    logmel_kernel<<<grid, kFeatWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("logmel");
}

extern "C" int dasv_cmn_f32(float* feat, const int32_t* frames, int B, int Tmax, int n_mels, int variance, void* stream) {
    if (B < 0 || Tmax < 0) { set_error("cmn: negative shape"); return 1; }
    if (B == 0 || Tmax == 0) return 0;
    if (!feat || !frames) { set_error("cmn: null pointer"); return 1; }
    if (n_mels < 1 || n_mels > 128) { set_error("cmn: n_mels %d outside [1, 128]", n_mels); return 1; }
    cmn_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(feat, frames, Tmax, n_mels, variance);
    return check_launch("cmn");
}
