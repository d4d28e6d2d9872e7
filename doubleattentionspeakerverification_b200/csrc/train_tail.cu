// Training-mode tail of SpeakerClassifier.forward (scripts/model.py:61-71) on the device:
//   * BatchNorm1d with batch statistics (model.py:67, `self.b2` in train mode): forward (+ running-statistics update,
//     momentum / unbiased variance as torch.nn.BatchNorm1d) and backward;
//   * AM-Softmax (scripts/loss.py:37-52): L2-normalised x and W, cosine logits, the margin subtracted AT THE LABEL ON THE
//     DEVICE (the reference scatters it on the CPU and copies it over, loss.py:45-48), annealing, scale; and its backward
//     through both normalisations.
// All of it is ~1.5 GFLOP per step against 13 TFLOP of convolutions: CUDA-core fp32 tiles, deterministic (no atomics),
// bound by launch latency.  The Linear layers around it (fc1, fc2, preLayer) are plain library GEMMs (cuBLAS via torch).
#include "common.cuh"
#include <math.h>

namespace dasv {

// ------------------------------------------------------------------------------ BatchNorm1d, batch statistics
constexpr int kBnCols = 32, kBnRows = 8;

__global__ void __launch_bounds__(kBnCols * kBnRows)
bn1d_train_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ y,
                      float* __restrict__ save_mean, float* __restrict__ save_invstd, int B, int E, float eps, float momentum) {
    __shared__ float red[kBnRows][kBnCols];
    const int fx = threadIdx.x % kBnCols, ry = threadIdx.x / kBnCols;
    const int f = blockIdx.x * kBnCols + fx;
    const bool ok = f < E;
    float s = 0.f;
    if (ok) for (int b = ry; b < B; b += kBnRows) s += x[static_cast<size_t>(b) * E + f];
    red[ry][fx] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int r = 0; r < kBnRows; ++r) mean += red[r][fx];
    mean /= static_cast<float>(B);
    __syncthreads();
    float q = 0.f;                                   // second pass over the (tiny) column: sum of squared deviations
    if (ok) for (int b = ry; b < B; b += kBnRows) { const float d = x[static_cast<size_t>(b) * E + f] - mean; q = fmaf(d, d, q); }
    red[ry][fx] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int r = 0; r < kBnRows; ++r) var += red[r][fx];
    var /= static_cast<float>(B);                    // biased: what normalises the batch
    const float invstd = rsqrtf(var + eps);
    if (ok) {
        const float g = gamma ? gamma[f] : 1.f, bt = beta ? beta[f] : 0.f;
        for (int b = ry; b < B; b += kBnRows) {
            const size_t i = static_cast<size_t>(b) * E + f;
            y[i] = fmaf((x[i] - mean) * invstd, g, bt);
        }
        if (ry == 0) {
            save_mean[f] = mean;
            save_invstd[f] = invstd;
            if (running_mean) running_mean[f] = (1.f - momentum) * running_mean[f] + momentum * mean;
            if (running_var) running_var[f] = (1.f - momentum) * running_var[f] + momentum * var * (B > 1 ? static_cast<float>(B) / (B - 1) : 1.f);
        }
    }
}

__global__ void __launch_bounds__(kBnCols * kBnRows)
bn1d_train_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                      const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                      float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int E) {
    __shared__ float r1[kBnRows][kBnCols], r2[kBnRows][kBnCols];
    const int fx = threadIdx.x % kBnCols, ry = threadIdx.x / kBnCols;
    const int f = blockIdx.x * kBnCols + fx;
    const bool ok = f < E;
    const float mean = ok ? save_mean[f] : 0.f, invstd = ok ? save_invstd[f] : 0.f;
    float sdy = 0.f, sdyx = 0.f;
    if (ok) for (int b = ry; b < B; b += kBnRows) {
        const size_t i = static_cast<size_t>(b) * E + f;
        const float g = dy[i];
        sdy += g;
        sdyx = fmaf(g, (x[i] - mean) * invstd, sdyx);
    }
    r1[ry][fx] = sdy; r2[ry][fx] = sdyx;
    __syncthreads();
    sdy = 0.f; sdyx = 0.f;
#pragma unroll
    for (int r = 0; r < kBnRows; ++r) { sdy += r1[r][fx]; sdyx += r2[r][fx]; }
    if (!ok) return;
    const float g = gamma ? gamma[f] : 1.f;
    const float k = g * invstd / static_cast<float>(B);
    for (int b = ry; b < B; b += kBnRows) {
        const size_t i = static_cast<size_t>(b) * E + f;
        const float xhat = (x[i] - mean) * invstd;
        dx[i] = k * (static_cast<float>(B) * dy[i] - sdy - xhat * sdyx);
    }
    if (ry == 0) { if (dgamma) dgamma[f] = sdyx; if (dbeta) dbeta[f] = sdy; }
}

// ------------------------------------------------------------------------------ small reductions
DASV_DEVICE float warp_sum_t(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// out[r] = 1 / max(||a[r,:]||, 1e-12)   (one warp per row)
__global__ void row_invnorm_kernel(const float* __restrict__ a, float* __restrict__ out, int R, int C) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = a[static_cast<size_t>(r) * C + c]; s = fmaf(v, v, s); }
    s = warp_sum_t(s);
    if (lane == 0) out[r] = 1.f / fmaxf(sqrtf(s), 1e-12f);
}
// out[c] = 1 / max(||a[:,c]||, 1e-12)   (one thread per column, rows in order: deterministic)
__global__ void col_invnorm_kernel(const float* __restrict__ a, float* __restrict__ out, int R, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int r = 0; r < R; ++r) { const float v = a[static_cast<size_t>(r) * C + c]; s = fmaf(v, v, s); }
    out[c] = 1.f / fmaxf(sqrtf(s), 1e-12f);
}
// G = dcosth + s * dlogits (either may be null), rb[b] = sum_s G*costh, then cs[s] = sum_b G*costh
__global__ void am_grad_rows_kernel(const float* __restrict__ dcosth, const float* __restrict__ dlogits, const float* __restrict__ costh,
                                    float* __restrict__ G, float* __restrict__ rb, int B, int S, float s) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    float acc = 0.f;
    for (int c = lane; c < S; c += 32) {
        const size_t i = static_cast<size_t>(b) * S + c;
        const float g = (dcosth ? dcosth[i] : 0.f) + (dlogits ? s * dlogits[i] : 0.f);
        G[i] = g;
        acc = fmaf(g, costh[i], acc);
    }
    acc = warp_sum_t(acc);
    if (lane == 0) rb[b] = acc;
}
__global__ void am_grad_cols_kernel(const float* __restrict__ G, const float* __restrict__ costh, float* __restrict__ cs, int B, int S) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S) return;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) { const size_t i = static_cast<size_t>(b) * S + c; acc = fmaf(G[i], costh[i], acc); }
    cs[c] = acc;
}

// ------------------------------------------------------------------------------ 64 x 64 fp32 tile GEMM with fused epilogues
// C[m,n] = sum_k A(m,k) * ka[k] * B(k,n);  A(m,k) = TA ? A[k*lda + m] : A[m*lda + k];  B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n].
struct GemmEpi {
    int mode;                 // 0: AM-Softmax forward, 1: dx of AM-Softmax, 2: dW of AM-Softmax
    const float* rowv;        // per-m vector (inverse norm of x rows / -)
    const float* colv;        // per-n vector (inverse norm of W columns)
    const float* aux;         // mode 1: rb [M]; mode 2: cs [N]
    const float* src;         // mode 1: x [M,N]; mode 2: W [M,N]
    const long long* label;   // mode 0
    float* out2;              // mode 0: logits
    float s, sm;              // mode 0: scale, s * m / (1 + alpha)
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_tile_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, const float* __restrict__ ka,
                 float* __restrict__ C, int M, int N, int K, const GemmEpi ep) {
    __shared__ float As[16][64 + 1], Bs[16][64 + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            int kk, mm;
            if (TA) { mm = i % 64; kk = i / 64; } else { kk = i % 16; mm = i / 16; }      // contiguous index fastest
            const int k = k0 + kk, m = m0 + mm;
            float v = 0.f;
            if (k < K && m < M) v = (TA ? A[static_cast<size_t>(k) * lda + m] : A[static_cast<size_t>(m) * lda + k]) * (ka ? ka[k] : 1.f);
            As[kk][mm] = v;
        }
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            int kk, nn;
            if (TB) { kk = i % 16; nn = i / 16; } else { nn = i % 64; kk = i / 64; }
            const int k = k0 + kk, n = n0 + nn;
            float v = 0.f;
            if (k < K && n < N) v = TB ? Bm[static_cast<size_t>(n) * ldb + k] : Bm[static_cast<size_t>(k) * ldb + n];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            const size_t o = static_cast<size_t>(m) * N + n;
            if (ep.mode == 0) {
                const float c = acc[i][j] * ep.rowv[m] * ep.colv[n];                    // loss.py:41-44
                C[o] = c;
                ep.out2[o] = ep.s * c - ((ep.label[m] == n) ? ep.sm : 0.f);             // loss.py:45-51
            } else if (ep.mode == 1) {
                const float ix = ep.rowv[m];
                C[o] = ix * acc[i][j] - ep.src[o] * ix * ix * ep.aux[m];
            } else {
                const float iw = ep.colv[n];
                C[o] = iw * acc[i][j] - ep.src[o] * iw * iw * ep.aux[n];
            }
        }
    }
}

}  // namespace dasv

using namespace dasv;

extern "C" int dasv_bn1d_train_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                   float* y, float* save_mean, float* save_invstd, int B, int E, float eps, float momentum, void* stream) {
    if (!x || !y || !save_mean || !save_invstd) { set_error("bn1d_train_fwd: null argument"); return 1; }
    if (B <= 0 || E <= 0) { set_error("bn1d_train_fwd: empty batch"); return 1; }
    bn1d_train_fwd_kernel<<<(E + kBnCols - 1) / kBnCols, kBnCols * kBnRows, 0, static_cast<cudaStream_t>(stream)>>>(
        x, gamma, beta, running_mean, running_var, y, save_mean, save_invstd, B, E, eps, momentum);
    return check_launch("bn1d_train_fwd");
}

extern "C" int dasv_bn1d_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                                   float* dx, float* dgamma, float* dbeta, int B, int E, void* stream) {
    if (!dy || !x || !save_mean || !save_invstd || !dx) { set_error("bn1d_train_bwd: null argument"); return 1; }
    if (B <= 0 || E <= 0) return 0;
    bn1d_train_bwd_kernel<<<(E + kBnCols - 1) / kBnCols, kBnCols * kBnRows, 0, static_cast<cudaStream_t>(stream)>>>(
        dy, x, gamma, save_mean, save_invstd, dx, dgamma, dbeta, B, E);
    return check_launch("bn1d_train_bwd");
}

// x [B,E], W [E,S] (reference layout, loss.py:20), label [B] int64 on the device -> costh [B,S], logits [B,S]; inv_x [B] and
// inv_w [S] are kept for the backward.  margin_scaled = s * m / (1 + alpha) (alpha: the annealing term, loss.py:28-35).
extern "C" int dasv_amsoftmax_fwd(const float* x, const float* W, const long long* label, float* costh, float* logits,
                                  float* inv_x, float* inv_w, int B, int E, int S, float s, float margin_scaled, void* stream) {
    if (!x || !W || !label || !costh || !logits || !inv_x || !inv_w) { set_error("amsoftmax_fwd: null argument"); return 1; }
    if (B <= 0 || S <= 0 || E <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    row_invnorm_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(x, inv_x, B, E);
    col_invnorm_kernel<<<(S + 255) / 256, 256, 0, st>>>(W, inv_w, E, S);
    GemmEpi ep{0, inv_x, inv_w, nullptr, nullptr, label, logits, s, margin_scaled};
    gemm_tile_kernel<false, false><<<dim3((S + 63) / 64, (B + 63) / 64), 256, 0, st>>>(x, E, W, S, nullptr, costh, B, S, E, ep);
    return check_launch("amsoftmax_fwd");
}

extern "C" size_t dasv_amsoftmax_bwd_workspace_bytes(int B, int S) {
    return (static_cast<size_t>(B > 0 ? B : 0) * (S > 0 ? S : 0) + (B > 0 ? B : 0) + (S > 0 ? S : 0)) * sizeof(float);
}

// dcosth / dlogits: gradients at the two outputs (either may be NULL) -> dx [B,E], dW [E,S].
extern "C" int dasv_amsoftmax_bwd(const float* dcosth, const float* dlogits, const float* x, const float* W, const float* costh,
                                  const float* inv_x, const float* inv_w, float* dx, float* dW, void* workspace,
                                  int B, int E, int S, float s, void* stream) {
    if (!x || !W || !costh || !inv_x || !inv_w || !dx || !dW || !workspace) { set_error("amsoftmax_bwd: null argument"); return 1; }
    if (B <= 0 || S <= 0 || E <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* G = static_cast<float*>(workspace);
    float* rb = G + static_cast<size_t>(B) * S;
    float* cs = rb + B;
    am_grad_rows_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(dcosth, dlogits, costh, G, rb, B, S, s);
    am_grad_cols_kernel<<<(S + 255) / 256, 256, 0, st>>>(G, costh, cs, B, S);
    // dx[b,e] = ix[b] * sum_s G[b,s] iw[s] W[e,s] - x[b,e] ix[b]^2 rb[b]
    GemmEpi e1{1, inv_x, nullptr, rb, x, nullptr, nullptr, 0.f, 0.f};
    gemm_tile_kernel<false, true><<<dim3((E + 63) / 64, (B + 63) / 64), 256, 0, st>>>(G, S, W, S, inv_w, dx, B, E, S, e1);
    // dW[e,s] = iw[s] * sum_b x[b,e] ix[b] G[b,s] - W[e,s] iw[s]^2 cs[s]
    GemmEpi e2{2, nullptr, inv_w, cs, W, nullptr, nullptr, 0.f, 0.f};
    gemm_tile_kernel<true, false><<<dim3((S + 63) / 64, (E + 63) / 64), 256, 0, st>>>(x, E, G, S, inv_x, dW, E, S, B, e2);
    return check_launch("amsoftmax_bwd");
}
