// Shared pieces of the fused DoubleMHA pooling kernels (forward and backward).
//
// Work decomposition (both directions):
//   * one CTA owns one utterance at a time (all H heads must meet for the head softmax) and
//     loops persistently over utterances b = blockIdx.x, blockIdx.x + gridDim.x, ...
//   * a producer warp streams the utterance's valid frames HBM -> SMEM with 1-D bulk async
//     copies (cp.async.bulk, the TMA engine's linear mode) into a ring of `stages` buffers of
//     `fps` frames each, signalling per-stage "full" mbarriers; 8 consumer warps release
//     stages through "empty" mbarriers.  x[b] is contiguous, so every copy is one linear burst.
//   * consumer threads are split in groups of G lanes (G = 8/16/32, a sub-warp); a group owns
//     head(s) and reads a head's row of one frame as NV 16-byte vectors per lane; row
//     reductions are xor-shuffles inside the group.
#pragma once
#include "common.cuh"

namespace dasv {

constexpr int kDmhaConsumerWarps = 8;
constexpr int kDmhaConsumerThreads = kDmhaConsumerWarps * 32;
constexpr int kDmhaThreads = kDmhaConsumerThreads + 32;   // + producer warp
constexpr int kDmhaFB = 4;                                 // frames processed together (ILP)

struct DmhaPlan {
    int bf16, G, NV, HPG, S, fps, stages;
    size_t smem_bytes;
    int err;    // 0 ok
};

template <int VE, bool BF16>
DASV_DEVICE void load_row_vec(const unsigned char* p, float (&f)[VE]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    if constexpr (BF16) {
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    } else {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
}

template <int G>
DASV_DEVICE float group_sum(float v) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

DASV_DEVICE float warp_sum(float v) { return group_sum<32>(v); }
DASV_DEVICE float warp_max(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

struct DmhaFwdParams {
    const unsigned char* x;
    const int32_t* lengths;
    const float* query;
    const float* att;
    const uint8_t* keep;
    float* out;
    float* ctx;
    float* lse;
    float* headw;
    float* align;
    int B, T, D, H, dh;
    int fps, stages, S;
    float scale_log2;   // log2(e) / sqrt(H): scores are kept in log2 units for ex2.approx
    int pitch3, rowcopy3; // used by scripts/ubench/dmha_fwd3.cu only (warp-MMA variant, no longer part of the library)
    int* ws_cnt;        // dmha_fwd2.cu: utterance counter for the dynamic deal (NULL = static round-robin)
    // dmha_fwd2.cu, head split (small batches): an utterance is cut into hs groups of H heads; B, H, D describe the
    // pseudo-utterances (B = utterances x hs), ldx is the row pitch of x and Hq the row pitch of query, in elements.
    // hs <= 1: off (ldx = D, Hq = H).
    int hs, ldx, Hq;
};

// The dynamic deal's workspace: [0] next utterance, [1] CTAs that have left.  The caller provides it zeroed; the last CTA
// to leave (every CTA has taken its last ticket by then) zeroes it again for the next launch on the same stream.
DASV_DEVICE void dmha_release_counter(int* ws_cnt) {
    if (ws_cnt == nullptr) return;
    __threadfence();
    if (atomicAdd(ws_cnt + 1, 1) == static_cast<int>(gridDim.x) - 1) {
        ws_cnt[0] = 0;
        ws_cnt[1] = 0;
        __threadfence();
    }
}

struct DmhaFwdSmem {
    uint32_t ring, q, a, pacc, pm, pl, u, w, bars, total;
};

__host__ __device__ inline DmhaFwdSmem dmha_fwd_smem(int D, int H, int dh, int S, int stages, uint32_t stage_bytes) {
    DmhaFwdSmem s;
    uint32_t o = 0;
    s.ring = o; o += stages * stage_bytes;
    s.q = o;    o += D * 4;
    s.a = o;    o += dh * 4;
    s.pacc = o; o += S * D * 4;
    s.pm = o;   o += H * S * 4;
    s.pl = o;   o += H * S * 4;
    s.u = o;    o += H * 4;
    s.w = o;    o += H * 4;
    o = (o + 7u) & ~7u;
    s.bars = o; o += 2 * stages * 8;
    s.total = o;
    return s;
}


// v2 forward (dmha_fwd2.cu): returns 0 = launched, 1 = error (message set), -1 = shape outside its mapping.
int dmha_fwd2_launch(DmhaFwdParams p, int x_dtype, void* workspace, cudaStream_t stream);
size_t dmha_fwd2_workspace_bytes(int B, int D, int H);

struct DmhaBwdParams {
    const unsigned char* x;
    const int32_t* lengths;
    const float* query;
    const float* att;
    const float* g_out;
    const float* g_ctx;
    const float* ctx;
    const float* lse;
    const float* headw;
    unsigned char* dx;
    float* ws_dq;      // [grid][D]
    float* ws_da;      // [grid][dh]
    int B, T, D, H, dh;
    int fps, stages, S;
    float scale_log2, inv_sqrt_h;
};

// v2 backward (dmha_bwd2.cu): same return convention as dmha_fwd2_launch; *grid_out = number of per-CTA partials written.
int dmha_bwd2_launch(DmhaBwdParams p, int x_dtype, int max_grid, int* grid_out, cudaStream_t stream);

// Host-side plan shared by forward and backward so both walk the ring identically.
DmhaPlan dmha_make_plan(int x_dtype, int T, int D, int H, bool backward);

}  // namespace dasv
