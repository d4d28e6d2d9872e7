// Embedding tail and trial scoring.
//   * dasv_fc_tail_f32: b2(relu(fc2(relu(fc1(pooled))))) in eval mode (scripts/model.py:56-57), one
//     fused kernel: 0.45 MFLOP per utterance, so the whole cost is weight traffic and launch latency.
//     An 8-CTA cluster owns up to 8 utterances; each CTA computes 1/8 of the columns of both layers.
//   * dasv_cosine_pairs / dasv_cosine_matrix: F.cosine_similarity(dim=-1, eps=1e-8) (scripts/utils.py:18-21)
//     batched over a trial list or a full enrol x test cross-product.
//   * dasv_attention_fwd: the single-query Attention pooling (scripts/poolings.py:22-27).
#include "common.cuh"
#include <math.h>

namespace dasv {

// A thread-block cluster of kTailCL CTAs owns UB utterances; CTA r computes output columns [r*CW, (r+1)*CW) of both
// layers, so every weight element is read once per cluster and the reads of one layer are spread over kTailCL SMs
// (the first version gave a single CTA all E columns: 78 us of dependent L2 round trips even for ONE utterance).
// Between the layers the relu(fc1) slices are exchanged through distributed shared memory (each CTA stores its slice
// into every CTA of the cluster) and one cluster barrier.
constexpr int kTailCL = 8;            // CTAs per cluster (portable maximum)
constexpr int kTailKS = 16;           // k slices per column (threads = 64 column lanes x kTailKS): the loads are latency-bound, so many short chains
constexpr int kTailThreads = 64 * kTailKS;

template <int UB>
DASV_DEVICE void tail_layer(const float* __restrict__ in_sm, int K, const float* __restrict__ wt, int E, int c0, int cw,
                            float* __restrict__ red_sm, float (&out)[UB]) {
    // out[u] = sum_k in_sm[u][k] * wt[k][c0 + el] for this thread's column el, reduced over the kTailKS k slices
    const int el = threadIdx.x & 63, ks = threadIdx.x >> 6;
    const int kper = (K + kTailKS - 1) / kTailKS;
    const int k0 = ks * kper, k1 = min(K, k0 + kper);
    float acc[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) acc[u] = 0.f;
    if (el < cw) {
        const float* wcol = wt + c0 + el;
        int k = k0;
        for (; k + 10 <= k1; k += 10) {          // 10 independent weight loads in flight per thread
            float w[10];
#pragma unroll
            for (int j = 0; j < 10; ++j) w[j] = __ldg(wcol + static_cast<size_t>(k + j) * E);
#pragma unroll
            for (int j = 0; j < 10; ++j)
#pragma unroll
                for (int u = 0; u < UB; ++u) acc[u] = fmaf(in_sm[u * K + k + j], w[j], acc[u]);
        }
        for (; k < k1; ++k) {
            const float w = __ldg(wcol + static_cast<size_t>(k) * E);
#pragma unroll
            for (int u = 0; u < UB; ++u) acc[u] = fmaf(in_sm[u * K + k], w, acc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) red_sm[(ks * UB + u) * 64 + el] = acc[u];
    __syncthreads();
    if (ks == 0) {
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < kTailKS; ++q) v += red_sm[(q * UB + u) * 64 + el];     // fixed order: deterministic
            out[u] = v;
        }
    }
}

template <int UB>
__global__ void __launch_bounds__(kTailThreads) fc_tail_kernel(const float* __restrict__ pooled, const float* __restrict__ w1t,
                                                             const float* __restrict__ b1, const float* __restrict__ w2t,
                                                             const float* __restrict__ b2, const float* __restrict__ bn_scale,
                                                             const float* __restrict__ bn_shift, float* __restrict__ emb,
                                                             int B, int Din, int E) {
    extern __shared__ float sm[];
    griddep_launch();
    griddep_wait();                          // programmatic dependent launch: `pooled` comes from the previous kernel
    float* in_sm = sm;                       // [UB][Din]
    float* h_sm = sm + UB * Din;             // [UB][E]   relu(fc1), assembled from all CTAs of the cluster
    float* red_sm = h_sm + UB * E;           // [kTailKS][UB][64]
    const uint32_t rank = cluster_ctarank();
    const int b0 = static_cast<int>(blockIdx.x / kTailCL) * UB;
    const int nb = min(UB, B - b0);
    const int CW = (E + kTailCL - 1) / kTailCL;          // columns per CTA (<= 64, checked by the host)
    const int c0 = static_cast<int>(rank) * CW, cw = max(0, min(CW, E - c0));
    for (int i = threadIdx.x; i < UB * Din; i += blockDim.x) {
        const int u = i / Din;
        in_sm[i] = u < nb ? pooled[static_cast<size_t>(b0) * Din + i] : 0.f;
    }
    __syncthreads();
    const int el = threadIdx.x & 63, ks = threadIdx.x >> 6;
    float v[UB];
    tail_layer<UB>(in_sm, Din, w1t, E, c0, cw, red_sm, v);
    if (ks == 0 && el < cw) {
        const float bb = b1[c0 + el];
        const uint32_t mine = smem_u32(h_sm + c0 + el);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const float h = fmaxf(v[u] + bb, 0.f);                                   // relu(fc1), model.py:56
            for (uint32_t r = 0; r < kTailCL; ++r) {                                 // into every CTA's copy of h
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(mine + static_cast<uint32_t>(u * E) * 4u), "r"(r));
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(h) : "memory");
            }
        }
    }
    cluster_sync_all();                      // release/acquire: every slice of h is visible in every CTA
    tail_layer<UB>(h_sm, E, w2t, E, c0, cw, red_sm, v);
    if (ks == 0 && el < cw) {
        const int e = c0 + el;
        const float bb = b2[e], sc = bn_scale[e], sh = bn_shift[e];
#pragma unroll
        for (int u = 0; u < UB; ++u)
            if (u < nb) emb[static_cast<size_t>(b0 + u) * E + e] = fmaf(fmaxf(v[u] + bb, 0.f), sc, sh);   // b2(relu(fc2)), model.py:57
    }
}

// ------------------------------------------------------------------------------ cosine scoring
DASV_DEVICE float warp_sum_f(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// one warp per trial
__global__ void cosine_pairs_kernel(const float* __restrict__ emb, const int32_t* __restrict__ ia, const int32_t* __restrict__ ib,
                                    float* __restrict__ scores, int n_pairs, int E) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const float* a = emb + static_cast<size_t>(ia[w]) * E;
    const float* b = emb + static_cast<size_t>(ib[w]) * E;
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int d = lane; d < E; d += 32) {
        const float x = a[d], y = b[d];
        ab = fmaf(x, y, ab); aa = fmaf(x, x, aa); bb = fmaf(y, y, bb);
    }
    ab = warp_sum_f(ab); aa = warp_sum_f(aa); bb = warp_sum_f(bb);
    if (lane == 0) scores[w] = ab / (fmaxf(sqrtf(aa), 1e-8f) * fmaxf(sqrtf(bb), 1e-8f));
}

__global__ void row_inv_norm_kernel(const float* __restrict__ x, float* __restrict__ inv, int n, int E) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    float s = 0.f;
    for (int d = lane; d < E; d += 32) { const float v = x[static_cast<size_t>(w) * E + d]; s = fmaf(v, v, s); }
    s = warp_sum_f(s);
    if (lane == 0) inv[w] = 1.f / fmaxf(sqrtf(s), 1e-8f);
}

// scores[i,j] = <enrol_i, test_j> * inv_e[i] * inv_t[j]; 64x64 tile, 16-deep slices, 4x4 per thread.
__global__ void __launch_bounds__(256) cosine_matrix_kernel(const float* __restrict__ en, const float* __restrict__ te,
                                                           const float* __restrict__ inv_e, const float* __restrict__ inv_t,
                                                           float* __restrict__ scores, int Ne, int Nt, int E) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const int lr = tid / 4, lk = (tid % 4) * 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < E; k0 += 16) {
        float a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            a[q] = (i0 + lr < Ne && k < E) ? en[static_cast<size_t>(i0 + lr) * E + k] : 0.f;
            b[q] = (j0 + lr < Nt && k < E) ? te[static_cast<size_t>(j0 + lr) * E + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) { As[lk + q][lr] = a[q]; Bs[lk + q][lr] = b[q]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i0 + ty * 4 + i;
        if (r >= Ne) continue;
        const float ie = inv_e[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = j0 + tx * 4 + j;
            if (c < Nt) scores[static_cast<size_t>(r) * Nt + c] = acc[i][j] * ie * inv_t[c];
        }
    }
}

// ------------------------------------------------------------------------------ EER threshold counts
// ge[k] = #{ i : scores[i] >= thresholds[k] } for the validation sweep of scripts/train.py:135-150 (200 thresholds,
// scripts/utils.py:5-15 `Score`): comparisons in double, like the reference's float(sc) >= float(th).  Each thread owns
// a score and walks the (shared-memory) threshold list; warp ballots keep the integer atomics to one per warp.
__global__ void __launch_bounds__(256) threshold_counts_kernel(const float* __restrict__ scores, int n,
                                                              const double* __restrict__ thresholds, int n_th,
                                                              unsigned long long* __restrict__ ge) {
    extern __shared__ double th_sm[];
    unsigned int* cnt_sm = reinterpret_cast<unsigned int*>(th_sm + n_th);
    for (int k = threadIdx.x; k < n_th; k += blockDim.x) { th_sm[k] = thresholds[k]; cnt_sm[k] = 0u; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x; base < n; base += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long i = base + threadIdx.x;
        const bool have = i < n;
        const double sc = have ? static_cast<double>(scores[i]) : 0.0;
        for (int k = 0; k < n_th; ++k) {
            const unsigned int m = __ballot_sync(0xffffffffu, have && sc >= th_sm[k]);
            if (lane == 0 && m) atomicAdd(&cnt_sm[k], __popc(m));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_th; k += blockDim.x)
        if (cnt_sm[k]) atomicAdd(&ge[k], static_cast<unsigned long long>(cnt_sm[k]));
}

// ------------------------------------------------------------------------------ Attention pooling
// scores: one warp per frame
__global__ void attention_scores_kernel(const unsigned char* __restrict__ x, int bf16, const float* __restrict__ att,
                                        float* __restrict__ align, int B, int T, int D) {
    const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= static_cast<long long>(B) * T) return;
    float s = 0.f;
    if (bf16) {
        const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(x) + w * D;
        for (int d = lane; d < D; d += 32) s = fmaf(__bfloat162float(r[d]), att[d], s);
    } else {
        const float* r = reinterpret_cast<const float*>(x) + w * D;
        for (int d = lane; d < D; d += 32) s = fmaf(r[d], att[d], s);
    }
    s = warp_sum_f(s);
    if (lane == 0) align[w] = s;
}
// softmax over time in place: one CTA per utterance
__global__ void attention_softmax_kernel(float* __restrict__ align, const int32_t* __restrict__ lengths,
                                         const uint8_t* __restrict__ keep, int T) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    float* a = align + static_cast<size_t>(b) * T;
    const int L = lengths ? min(max(lengths[b], 0), T) : T;
    if (keep != nullptr) {      // dropped positions score -inf (HeadAttention training mode, poolings.py:39-43)
        for (int t = threadIdx.x; t < L; t += blockDim.x)
            if (keep[static_cast<size_t>(b) * T + t] == 0) a[t] = -INFINITY;
        __syncthreads();
    }
    float m = -INFINITY;
    for (int t = threadIdx.x; t < L; t += blockDim.x) m = fmaxf(m, a[t]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = -INFINITY;
    for (int i = 0; i < (blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int t = threadIdx.x; t < L; t += blockDim.x) s += expf(a[t] - m);
    s = warp_sum_f(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
    const float inv = 1.f / s;      // everything masked -> NaN, as torch.softmax of all -inf
    for (int t = threadIdx.x; t < T; t += blockDim.x) a[t] = t < L ? expf(a[t] - m) * inv : 0.f;
}
// out[b,d] = sum_t p[b,t] x[b,t,d]; thread per (b, d)
__global__ void attention_sum_kernel(const unsigned char* __restrict__ x, int bf16, const float* __restrict__ align,
                                     const int32_t* __restrict__ lengths, float* __restrict__ out, int B, int T, int D) {
    const int b = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int L = lengths ? min(max(lengths[b], 0), T) : T;
    const float* p = align + static_cast<size_t>(b) * T;
    float acc = 0.f;
    if (bf16) {
        const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(x) + static_cast<size_t>(b) * T * D + d;
        for (int t = 0; t < L; ++t) acc = fmaf(p[t], __bfloat162float(r[static_cast<size_t>(t) * D]), acc);
    } else {
        const float* r = reinterpret_cast<const float*>(x) + static_cast<size_t>(b) * T * D + d;
        for (int t = 0; t < L; ++t) acc = fmaf(p[t], r[static_cast<size_t>(t) * D], acc);
    }
    out[static_cast<size_t>(b) * D + d] = acc;
}

}  // namespace dasv

using namespace dasv;

extern "C" int dasv_fc_tail_f32(const float* pooled, const float* w1t, const float* b1, const float* w2t,
                                const float* b2, const float* bn_scale, const float* bn_shift, float* emb,
                                int B, int Din, int E, void* stream) {
    if (!pooled || !w1t || !b1 || !w2t || !b2 || !bn_scale || !bn_shift || !emb) { set_error("fc_tail: null argument"); return 1; }
    if (B <= 0) return 0;
    if ((E + kTailCL - 1) / kTailCL > 64) { set_error("fc_tail: E=%d > %d", E, 64 * kTailCL); return 1; }
    // utterances per cluster: as many as shared memory holds (the fc1 input rows are the large part for the MHA pooling)
    int UB = 8;
    auto smem_of = [&](int ub) { return (static_cast<size_t>(ub) * (Din + E) + static_cast<size_t>(kTailKS) * ub * 64) * sizeof(float); };
    while (UB > 1 && (smem_of(UB) > 160 * 1024 || UB / 2 >= B)) UB /= 2;
    if (UB == 2) UB = 1;
    const size_t smem = smem_of(UB);
    if (smem > 200 * 1024) { set_error("fc_tail: Din=%d E=%d need %zu B of shared memory", Din, E, smem); return 1; }
    auto kern = UB == 8 ? fc_tail_kernel<8> : (UB == 4 ? fc_tail_kernel<4> : fc_tail_kernel<1>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("fc_tail: smem attribute: %s", cudaGetErrorString(e)); return 1; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>((B + UB - 1) / UB) * kTailCL);
    cfg.blockDim = dim3(kTailThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kTailCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    e = cudaLaunchKernelEx(&cfg, kern, pooled, w1t, b1, w2t, b2, bn_scale, bn_shift, emb, B, Din, E);
    if (e != cudaSuccess) { set_error("fc_tail: launch failed: %s", cudaGetErrorString(e)); return 1; }
    return check_launch("fc_tail");
}

extern "C" int dasv_cosine_pairs(const float* emb, const int32_t* ia, const int32_t* ib, float* scores,
                                 int n_pairs, int E, void* stream) {
    if (!emb || !ia || !ib || !scores) { set_error("cosine_pairs: null argument"); return 1; }
    if (n_pairs <= 0) return 0;
    const long long threads = static_cast<long long>(n_pairs) * 32;
    cosine_pairs_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        emb, ia, ib, scores, n_pairs, E);
    return check_launch("cosine_pairs");
}

extern "C" size_t dasv_cosine_matrix_workspace_bytes(int Ne, int Nt) {
    return (static_cast<size_t>(Ne > 0 ? Ne : 0) + (Nt > 0 ? Nt : 0)) * sizeof(float);
}

extern "C" int dasv_cosine_matrix(const float* enrol, const float* test, float* scores, void* workspace,
                                  int Ne, int Nt, int E, void* stream) {
    if (!enrol || !test || !scores || !workspace) { set_error("cosine_matrix: null argument"); return 1; }
    if (Ne <= 0 || Nt <= 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* inv_e = static_cast<float*>(workspace);
    float* inv_t = inv_e + Ne;
    row_inv_norm_kernel<<<(Ne * 32 + 255) / 256, 256, 0, s>>>(enrol, inv_e, Ne, E);
    row_inv_norm_kernel<<<(Nt * 32 + 255) / 256, 256, 0, s>>>(test, inv_t, Nt, E);
    dim3 grid((Nt + 63) / 64, (Ne + 63) / 64);
    cosine_matrix_kernel<<<grid, 256, 0, s>>>(enrol, test, inv_e, inv_t, scores, Ne, Nt, E);
    return check_launch("cosine_matrix");
}

extern "C" int dasv_threshold_counts(const float* scores, int n, const double* thresholds, int n_th,
                                     unsigned long long* ge_counts, void* stream) {
    if (!thresholds || !ge_counts || (n > 0 && !scores)) { set_error("threshold_counts: null argument"); return 1; }
    if (n_th <= 0 || n_th > 4096) { set_error("threshold_counts: n_th=%d must be in 1..4096", n_th); return 1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(ge_counts, 0, static_cast<size_t>(n_th) * sizeof(unsigned long long), s);
    if (e != cudaSuccess) { set_error("threshold_counts: memset: %s", cudaGetErrorString(e)); return 1; }
    if (n <= 0) return 0;
    int grid = (n + 255) / 256;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    const size_t smem = static_cast<size_t>(n_th) * (sizeof(double) + sizeof(unsigned int));
    threshold_counts_kernel<<<grid, 256, smem, s>>>(scores, n, thresholds, n_th, ge_counts);
    return check_launch("threshold_counts");
}

extern "C" int dasv_attention_fwd(const void* x, int x_dtype, const int32_t* lengths, const uint8_t* keep,
                                  const float* att, float* out, float* align, int B, int T, int D, void* stream) {
    if (!x || !att || !out || !align) { set_error("attention_fwd: null argument (align is required as workspace)"); return 1; }
    if (x_dtype != 0 && x_dtype != 1) { set_error("attention_fwd: bad dtype %d", x_dtype); return 1; }
    if (B <= 0 || T <= 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned char* xp = static_cast<const unsigned char*>(x);
    const long long threads = static_cast<long long>(B) * T * 32;
    attention_scores_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(xp, x_dtype, att, align, B, T, D);
    attention_softmax_kernel<<<B, 256, 0, s>>>(align, lengths, keep, T);
    attention_sum_kernel<<<dim3((D + 255) / 256, B), 256, 0, s>>>(xp, x_dtype, align, lengths, out, B, T, D);
    return check_launch("attention_fwd");
}
