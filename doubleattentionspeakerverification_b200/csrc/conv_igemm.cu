// bf16 tensor-core implicit-GEMM 3x3 convolution for the VGG front-end (scripts/CNNs.py:73-86),
// written for sm_100a: TMA -> swizzled SMEM -> tcgen05.mma (accumulators in TMEM) -> tcgen05.ld
// epilogue with bias + ReLU + length mask (+ 2x2 ceil-mode max-pool, + the front-end's final
// [B,T',C*F'] re-layout) fused, so activations make one HBM round trip per layer.
//
// GEMM orientation (chosen for the epilogue):  D[co, pixel] = sum_k W[co, k] * X[pixel, k],
//   k = (tap, ci).  A = packed weights [Cout_pad][9*Cin] (K-major, 128 rows per CTA tile = the 128
//   TMEM lanes), B = an activation patch of N = BF x BT x BB pixels (f, t, utterance) loaded by ONE
//   4-D TMA box per (tap, 64-channel slice): the box origin is shifted by the tap offset and TMA's
//   out-of-bounds zero fill implements the conv padding (also across utterance boundaries).
//   With output channels on TMEM lanes a thread owns one channel and sees pixels along TMEM
//   columns, so the 2x2 pool window (columns j, j+1, j+BF, j+BF+1) is thread-local, and the final
//   layer's feature order c*F'+f is a contiguous store.
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//   warps 4-11 = epilogue (two warps per TMEM lane quarter).  Pipelines: SMEM ring full/empty, double-buffered TMEM accumulator
//   full/empty, persistent static tile schedule (tile = blockIdx.x + i*gridDim.x).
#include "common.cuh"
#include "tmap.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

namespace dasv {

constexpr int kConvThreads = 384;              // 4 role warps + 8 epilogue warps
constexpr int kConvFuseThreads = 256;          // fused first layer: 8 more warps compute conv11 patches
constexpr int kConvTileM = 128;                 // output channels per CTA tile = TMEM lanes
constexpr int kConvKC = 64;                     // channels per K slice (64 bf16 = one 128-byte swizzle row)
constexpr uint32_t kConvABytes = kConvTileM * kConvKC * 2;
constexpr int kConvEpiChunk = 32;               // output pixels staged per epilogue chunk
constexpr uint32_t kConvEpiBytes = kConvEpiChunk * kConvTileM * 2;   // [32][128] bf16; double-buffered per epilogue half

struct ConvParams {
    const float* bias;
    const int32_t* lengths;
    void* y;
    int B, T, F, Cin, Cout;
    int BF, BT, BB;          // patch: BF frequency bins x BT frames x BB utterances
    int N, Npad;             // pixels per patch, rounded up to 16 (UMMA N)
    int n_ft, n_tt, n_bt, n_mt;
    int kchunks;             // K slices of 64 channels per tap: Cin / 64 (3 * Cin / 64 in split mode)
    int a_tap_stride;        // columns of the packed weights per tap: Cin (3 * Cin in split mode)
    int kcx_wrap;            // split mode: K slices >= this one re-read the activation slices from the start (x_hi again)
    int stages;
    int pool, ref_layout, y_f32;
    int relu;                // 0: linear epilogue (input-gradient pass), only without POOL
    const void* mask;        // optional bf16 tensor shaped like y (NHWC): outputs are zeroed where mask <= 0 (ReLU backward)
    uint32_t b_bytes;        // bytes one B box delivers
    uint32_t stage_bytes;    // per ring-1 stage: A + B(padded) (per-tap mode) or one B patch (tap-row reuse mode)
    uint32_t tmem_cols;
    int reuse;               // 1: B patches carry a +-1 frame halo and serve the three taps of a column (dy = -1,0,1)
    int gap_cols;            // unused accumulator columns between the blocks of consecutive utterances of a patch
                             // (2*BF halo rows in single-CTA reuse mode; the 8-row padding of a half in pair mode; else 0)
    int sa;                  // reuse mode: stages of the separate A (weight tile) ring
    int pair;                // 1: CTA pairs (cta_group::2): 256 channels x N pixels per pair, each CTA loads half of the patch
    int w_f16, x_f16;        // operand formats of the MMA: weights / activations are fp16 (else bf16)
    // fused first layer (FUSE11 kernels): conv11 (1 -> Cin channels, scripts/CNNs.py:72) is computed by four extra warps
    // into a per-CTA, double-buffered scratch patch in global memory (L2-resident) that the TMA producer reads instead of x
    const float* x0;         // [B,T,F] f32 input of conv11
    const float* w11;        // [Cin,1,3,3] f32
    const float* b11;        // [Cin] f32
    void* scratch;           // [grid][2][BT+2][BF+2][Cin] 16-bit
    uint32_t fuse_off;       // byte offset of the fused path's shared memory (input patch + barriers) from the barrier block
    // split-K for launches with too few tiles to fill the SMs (small batches): a tile's K slices are dealt to `splitk` CTAs,
    // each stores its raw fp32 accumulator to ws [splitk][B*T*F][Cout]; conv_splitk_finish_kernel sums, adds the bias and
    // applies ReLU / pool / format.  kpc = K slices per split.
    int splitk, kpc;
    float* ws;
    int wide_epi;            // pooled epilogue: 16- / 8-column TMEM loads where a pooled row of the patch allows it
    int lazy_mask;           // DASV_CONV_LAZY_MASK: all-masked tiles are only zero-filled where the next layer reads them (see conv_tile_needs_zeros)
    int balanced;            // ragged batches: tiles with valid frames are compacted through a per-CTA prefix table so that every
                             // CTA gets the same number of them (and of the all-masked tiles, which only store zeros)
    unsigned long long* trace;   // debugging aid (dasv_debug_conv_trace): per CTA 8 x globaltimer stamps, or nullptr
    int RT;                  // pair mode: frames of the patch half one CTA loads (without halo)
    int split_t;             // pair mode: halves split along t (BB == 1) or along the utterance (BB == 2)
};

struct ConvTile {
    int m, f0, t0, b0, split;
};

DASV_DEVICE ConvTile conv_decode_tile(const ConvParams& p, int tile, int n_mt_eff, int rank) {
    ConvTile c;
    c.split = 0;
    if (p.splitk > 1) { c.split = tile % p.splitk; tile /= p.splitk; }      // the K splits of a tile run side by side
    c.m = tile % n_mt_eff;
    if (p.pair) c.m = c.m * 2 + rank;        // a pair owns two adjacent 128-channel tiles
    int pt = tile / n_mt_eff;
    c.f0 = (pt % p.n_ft) * p.BF; pt /= p.n_ft;
    c.t0 = (pt % p.n_tt) * p.BT; pt /= p.n_tt;
    c.b0 = pt * p.BB;
    return c;
}

DASV_DEVICE int conv_len(const ConvParams& p, int b) {
    if (b >= p.B) return 0;
    return p.lengths ? min(max(p.lengths[b], 0), p.T) : p.T;
}

// A patch whose every frame lies at or beyond its utterance's valid length produces only zeros.
DASV_DEVICE bool conv_tile_masked(const ConvParams& p, const ConvTile& c) {
    if (p.lengths == nullptr) return false;
    for (int bb = 0; bb < p.BB; ++bb)
        if (conv_len(p, c.b0 + bb) > c.t0) return false;
    return true;
}

// Lazy masking (inference pipelines): the next layer reads, of the rows at or beyond an utterance's length L, only row L
// (the bottom halo of its last valid row; after a 2x2 ceil pool: row 2*ceil(L/2) of this layer, i.e. pooled row ceil(L/2)).
// Whatever lies further down only feeds output rows that are masked themselves.  A tile that straddles L zeroes its rows
// >= L anyway; an ALL-masked tile therefore needs zeros only if that one row is its first row -- otherwise it is skipped
// and the memory behind it stays unwritten.
DASV_DEVICE bool conv_tile_needs_zeros(const ConvParams& p, const ConvTile& c) {
    if (!p.lazy_mask) return true;
    for (int bb = 0; bb < p.BB; ++bb) {
        const int L = conv_len(p, c.b0 + bb);
        if (c.b0 + bb < p.B && (p.pool ? ((L + 1) & ~1) : L) == c.t0) return true;
    }
    return false;
}

// ---- balanced schedule for ragged batches.  A group = the BB utterances of one patch column; its first vt t-tiles
// hold valid frames, the rest are all-masked.  vpre[g] = number of valid tiles in groups < g (built once per CTA).
constexpr int kConvMaxGroups = 2047;

DASV_DEVICE int conv_group_vt(const ConvParams& p, int g) {
    int Lmax = 0;
    for (int bb = 0; bb < p.BB; ++bb) Lmax = max(Lmax, conv_len(p, g * p.BB + bb));
    return (Lmax + p.BT - 1) / p.BT;
}

struct ConvSched {
    const int* vpre;         // [n_bt + 1] (balanced mode)
    int n_pass0, n_pass1;    // tiles of the main pass (valid ones when balanced, else all) and of the zero-fill pass
    int per_t;               // tiles per (group, t-tile): n_mt_eff * n_ft
};

// Tile q of pass `pass`.  Unbalanced mode (no lengths, or too many groups): the plain order, masked tiles flagged.
DASV_DEVICE ConvTile conv_tile_at(const ConvParams& p, const ConvSched& sc, int q, int pass, int n_mt_eff, int rank, bool& masked) {
    if (!p.balanced) {
        const ConvTile c = conv_decode_tile(p, q, n_mt_eff, rank);
        masked = conv_tile_masked(p, c);
        return c;
    }
    const int per_g = sc.per_t * p.n_tt;
    int lo = 0, hi = p.n_bt;             // the group g with pre(g) <= q < pre(g + 1)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        const int pre = pass == 0 ? sc.vpre[mid] : mid * per_g - sc.vpre[mid];
        if (pre <= q) lo = mid; else hi = mid;
    }
    const int g = lo;
    int local = q - (pass == 0 ? sc.vpre[g] : g * per_g - sc.vpre[g]);
    const int vt = (sc.vpre[g + 1] - sc.vpre[g]) / sc.per_t;
    ConvTile c;
    c.split = 0;
    c.m = local % n_mt_eff;
    if (p.pair) c.m = c.m * 2 + rank;
    local /= n_mt_eff;
    c.f0 = (local % p.n_ft) * p.BF;
    c.t0 = (local / p.n_ft + (pass == 0 ? 0 : vt)) * p.BT;
    c.b0 = g * p.BB;
    masked = pass != 0;
    return c;
}

DASV_DEVICE void conv_trace(const ConvParams& p, int slot) {
    if (p.trace != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.trace[static_cast<size_t>(blockIdx.x) * 8 + slot] = t;
    }
}

DASV_DEVICE void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
DASV_DEVICE void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
DASV_DEVICE void tmem_ld_x1(uint32_t taddr, uint32_t& r0) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(taddr) : "memory");
}
DASV_DEVICE void tmem_ld_x2(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
}

// 0xFFFF per half where the packed bf16 value is > 0
DASV_DEVICE uint32_t bf16x2_positive_mask(uint32_t v) {
    const uint32_t lo = ((v & 0x8000u) == 0u && (v & 0x7FFFu) != 0u) ? 0x0000FFFFu : 0u;
    const uint32_t hi = ((v & 0x80000000u) == 0u && (v & 0x7FFF0000u) != 0u) ? 0xFFFF0000u : 0u;
    return lo | hi;
}

template <bool F32, int ACT>
DASV_DEVICE void conv_store(void* y, size_t idx, float v) {
    if (F32) static_cast<float*>(y)[idx] = v;
    else static_cast<uint16_t*>(y)[idx] = cvt16_bits<ACT == 3 ? 1 : ACT>(v);
}

// DGRAD = false: the forward layer (bias + ReLU (+ pool)).  DGRAD = true: the input-gradient pass (linear epilogue, optional
// ReLU-backward mask); a separate instantiation so that the forward's epilogue carries none of its branches.
// ACT = format of a 16-bit output: 1 = bf16, 2 = fp16 (saturating), 3 = split bf16 (the fp32x3 mode: an fp32 value v is
// stored as hi = bf16(v) in channel c and lo = bf16(v - hi) in channel Cout + c of a 2*Cout-channel tensor).
template <bool PAIR, bool DGRAD = false, int ACT = 1, bool FUSE11 = false>
__global__ void __launch_bounds__(kConvThreads + (FUSE11 ? kConvFuseThreads : 0), 1)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles: align the dynamic window by hand.
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* ring = smem;                                                       // ring 1
    unsigned char* ring_a = smem + static_cast<size_t>(p.stages) * p.stage_bytes;     // ring 2: weight tiles (reuse mode)
    unsigned char* stage = ring_a + static_cast<size_t>(p.sa) * kConvABytes;          // epilogue staging, 2 halves x 2 x kConvEpiBytes
    uint64_t* full = reinterpret_cast<uint64_t*>(stage + 4 * kConvEpiBytes);
    uint64_t* empty = full + p.stages;
    uint64_t* afull = empty + p.stages;
    uint64_t* aempty = afull + p.sa;
    uint64_t* acc_full = aempty + p.sa;         // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int* vpre = reinterpret_cast<int*>(tmem_slot + 4);     // [n_bt + 1], balanced mode only (the host sized the window for it)
    uint64_t* s_full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(full) + p.fuse_off);   // [2] FUSE11: scratch patch written
    uint64_t* s_free = s_full + 2;                                                                          // [2] FUSE11: its MMAs have retired
    float* x_sm = reinterpret_cast<float*>(s_free + 2);                                                    // FUSE11: [(BT+4)][(BF+4)] input patch

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;                  // 0 = leader of the CTA pair
    const int n_mt_eff = PAIR ? p.n_mt / 2 : p.n_mt;
    const int n_tiles = n_mt_eff * p.n_ft * p.n_tt * p.n_bt * (p.splitk > 1 ? p.splitk : 1);
    const int tile0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int ksteps = 9 * p.kchunks;

    if (threadIdx.x == 0) {
        conv_trace(p, 0);                                   // CTA start
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < p.sa; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], PAIR ? 16 : 8); }
        if (FUSE11) for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 1); }
        fence_mbar_init();
    }
    if (threadIdx.x == 32) griddep_launch();   // the stream's next kernel may begin its own prologue (it waits for this grid below)
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 2) {
        if (PAIR) { tmem_alloc_2sm(tmem_slot, p.tmem_cols); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();       // the peer's barriers must exist before remote arrivals / TMA completions reach them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above overlapped the previous kernel's tail (programmatic dependent launch); x, the mask and the
    // memory behind y belong to earlier kernels of the stream from here on
    if (threadIdx.x == 0) conv_trace(p, 1);                 // set-up done (barriers, TMEM, descriptors)
    griddep_wait();
    if (threadIdx.x == 0) conv_trace(p, 2);                 // the previous kernel of the stream has finished
    ConvSched sched{vpre, n_tiles, 0, n_mt_eff * p.n_ft};
    if (p.balanced) {
        if (warp == 3) {                                    // the spare role warp scans the groups' valid-tile counts
            int carry = 0;
            for (int g0 = 0; g0 < p.n_bt; g0 += 32) {
                const int g = g0 + lane;
                int v = g < p.n_bt ? conv_group_vt(p, g) * sched.per_t : 0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, v, off);
                    if (lane >= off) v += u;
                }
                if (g < p.n_bt) vpre[g + 1] = carry + v;
                carry += __shfl_sync(0xffffffffu, v, 31);
            }
            if (lane == 0) vpre[0] = 0;
        }
        __syncthreads();
        sched.n_pass0 = vpre[p.n_bt];
        sched.n_pass1 = n_tiles - sched.n_pass0;
    }

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0 && p.reuse) {
            // tap-row reuse: per (64-channel slice, dx) ONE activation patch with a +-1 frame halo, then the three
            // weight tiles of that tap column (dy = -1, 0, +1); 3x less activation traffic than one box per tap.
            uint32_t sb = 0, bph = 0, sa = 0, aph = 0, fit = 0;
            for (int tile = tile0; tile < sched.n_pass0; tile += tile_step) {
                bool tile_masked;
                const ConvTile c = conv_tile_at(p, sched, tile, 0, n_mt_eff, static_cast<int>(rank), tile_masked);
                if (tile_masked) continue;
                const int slot = static_cast<int>(blockIdx.x) * 2 + static_cast<int>(fit & 1u);    // FUSE11: this tile's scratch patch
                if (FUSE11) { mbar_wait(&s_full[fit & 1u], (fit >> 1) & 1u); ++fit; }
                const int kc0 = p.splitk > 1 ? c.split * p.kpc : 0;
                const int kc1 = p.splitk > 1 ? min(p.kchunks, kc0 + p.kpc) : p.kchunks;
                for (int kc = kc0; kc < kc1; ++kc) {
                    const int kcx = kc >= p.kcx_wrap ? kc - p.kcx_wrap : kc;     // activation slice of this K slice
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        mbar_wait(&empty[sb], bph ^ 1u);
                        if (PAIR) {
                            // each CTA fetches ITS half of the patch (+ halo) into its own SMEM; both loads complete on
                            // the leader's barrier, which the leader arms for the bytes of both
                            const int tq = c.t0 + (p.split_t ? static_cast<int>(rank) * p.RT : 0);
                            const int bq = c.b0 + (p.split_t ? 0 : static_cast<int>(rank));
                            if (rank == 0) mbar_arrive_expect_tx(&full[sb], 2u * p.b_bytes);
                            tma_load_4d_2sm(ring + static_cast<size_t>(sb) * p.stage_bytes, &tmB, &full[sb], kcx * kConvKC, c.f0 + dxi - 1, tq - 1, bq);
                        } else {
                            mbar_arrive_expect_tx(&full[sb], p.b_bytes);
                            if (FUSE11)      // the scratch patch holds frames t0-1 .. t0+BT and bins f0-1 .. f0+BF, zeros outside the image
                                tma_load_4d(ring + static_cast<size_t>(sb) * p.stage_bytes, &tmB, &full[sb], kcx * kConvKC, dxi, 0, slot);
                            else
                                tma_load_4d(ring + static_cast<size_t>(sb) * p.stage_bytes, &tmB, &full[sb], kcx * kConvKC, c.f0 + dxi - 1, c.t0 - 1, c.b0);
                        }
                        if (++sb == static_cast<uint32_t>(p.stages)) { sb = 0; bph ^= 1u; }
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            mbar_wait(&aempty[sa], aph ^ 1u);
                            if (PAIR) {
                                if (rank == 0) mbar_arrive_expect_tx(&afull[sa], 2u * kConvABytes);
                                tma_load_2d_2sm(ring_a + static_cast<size_t>(sa) * kConvABytes, &tmA, &afull[sa],
                                                (dyi * 3 + dxi) * p.a_tap_stride + kc * kConvKC, c.m * kConvTileM);
                            } else {
                                mbar_arrive_expect_tx(&afull[sa], kConvABytes);
                                tma_load_2d(ring_a + static_cast<size_t>(sa) * kConvABytes, &tmA, &afull[sa],
                                            (dyi * 3 + dxi) * p.a_tap_stride + kc * kConvKC, c.m * kConvTileM);
                            }
                            if (++sa == static_cast<uint32_t>(p.sa)) { sa = 0; aph ^= 1u; }
                        }
                    }
                }
            }
        } else if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int tile = tile0; tile < sched.n_pass0; tile += tile_step) {
                bool tile_masked;
                const ConvTile c = conv_tile_at(p, sched, tile, 0, n_mt_eff, static_cast<int>(rank), tile_masked);
                if (tile_masked) continue;
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&empty[st], ph ^ 1u);
                        unsigned char* a_sm = ring + static_cast<size_t>(st) * p.stage_bytes;
                        mbar_arrive_expect_tx(&full[st], kConvABytes + p.b_bytes);
                        tma_load_2d(a_sm, &tmA, &full[st], tap * p.a_tap_stride + kc * kConvKC, c.m * kConvTileM);
                        tma_load_4d(a_sm + kConvABytes, &tmB, &full[st], (kc >= p.kcx_wrap ? kc - p.kcx_wrap : kc) * kConvKC, c.f0 + dx, c.t0 + dy, c.b0);
                        if (++st == static_cast<uint32_t>(p.stages)) { st = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0 && rank == 0) {           // in pair mode only the leader issues (for both CTAs)
            const uint32_t idesc = umma_idesc_f16kind(PAIR ? 2 * kConvTileM : kConvTileM, static_cast<uint32_t>(p.Npad), p.w_f16 != 0, p.x_f16 != 0);
            uint32_t st = 0, ph = 0, sa = 0, aph = 0, acc_it = 0;
            for (int tile = tile0; tile < sched.n_pass0; tile += tile_step) {
                bool tile_masked;
                const ConvTile c = conv_tile_at(p, sched, tile, 0, n_mt_eff, static_cast<int>(rank), tile_masked);
                if (tile_masked) continue;
                // split mode: the two accumulators of a tile fill TMEM (2 x 256 columns), so tiles are not double-buffered there
                const uint32_t as = ACT == 3 ? 0u : (acc_it & 1u), accph = ACT == 3 ? (acc_it & 1u) : ((acc_it >> 1) & 1u);
                mbar_wait(&acc_empty[as], accph ^ 1u);
                tc_fence_after();
                // Split (fp32x3) mode keeps TWO accumulators per tile: [0] the hi*hi products, [1] the two cross terms, which are
                // 2^-8 smaller.  The tensor core adds into its fp32 accumulator with truncation, a bias of up to one ulp OF THE
                // ACCUMULATOR per MMA; with the cross terms elsewhere the large accumulator sees a third of the additions and the
                // small one's ulps are negligible (measured: embeddings 1.4e-4 -> within the 1e-4 bar).  Summed in the epilogue.
                constexpr bool X3 = ACT == 3;
                const uint32_t d_tmem0 = tmem_base + as * static_cast<uint32_t>(p.Npad);
                const int n_hi = p.kcx_wrap >> 1;                       // split mode: K slices [0, n_hi) are hi*hi
                uint32_t firstj[2] = {1u, 1u};
                if (p.reuse) {
                    const int n_groups = 3 * (p.splitk > 1 ? min(p.kchunks, (c.split + 1) * p.kpc) - c.split * p.kpc : p.kchunks);
                    for (int g = 0; g < n_groups; ++g) {                // (slice, dx) groups
                        const uint32_t jacc = (X3 && g / 3 >= n_hi) ? 1u : 0u;
                        const uint32_t d_tmem = d_tmem0 + jacc * static_cast<uint32_t>(p.Npad);
                        uint32_t first = firstj[jacc];
                        firstj[jacc] = 0u;
                        mbar_wait(&full[st], ph);
                        const uint32_t b_addr = smem_u32(ring + static_cast<size_t>(st) * p.stage_bytes);
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            mbar_wait(&afull[sa], aph);
                            tc_fence_after();
                            if (g == 0 && dyi == 0 && acc_it == 0) conv_trace(p, 3);          // first operands have landed
                            const uint64_t a_desc = umma_desc_k128(smem_u32(ring_a + static_cast<size_t>(sa) * kConvABytes));
                            // view of the patch shifted by dyi frames: BF rows of 128 B per frame
                            // (the 128-byte swizzle is a function of absolute SMEM address bits, so a view may start at any row:
                            //  measured on B200 -- the descriptor's base-offset field must stay 0 for this)
                            const uint64_t b_desc = umma_desc_k128(b_addr + static_cast<uint32_t>(dyi * p.BF) * 128u);
#pragma unroll
                            for (int k = 0; k < kConvKC / 16; ++k) {
                                if (PAIR) umma_bf16_2sm(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                                                        (first && k == 0) ? 0u : 1u);
                                else umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                                               (first && k == 0) ? 0u : 1u);
                            }
                            first = 0u;
                            if (PAIR) umma_commit_2sm(&aempty[sa], 3); else umma_commit(&aempty[sa]);
                            if (++sa == static_cast<uint32_t>(p.sa)) { sa = 0; aph ^= 1u; }
                        }
                        if (PAIR) umma_commit_2sm(&empty[st], 3); else umma_commit(&empty[st]);   // patch free once its three tap rows have retired
                        if (++st == static_cast<uint32_t>(p.stages)) { st = 0; ph ^= 1u; }
                    }
                } else {
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t jacc = (X3 && ks % p.kchunks >= n_hi) ? 1u : 0u;
                        const uint32_t d_tmem = d_tmem0 + jacc * static_cast<uint32_t>(p.Npad);
                        const uint32_t first = firstj[jacc];
                        firstj[jacc] = 0u;
                        mbar_wait(&full[st], ph);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(ring + static_cast<size_t>(st) * p.stage_bytes);
                        const uint64_t a_desc = umma_desc_k128(a_addr);
                        const uint64_t b_desc = umma_desc_k128(a_addr + kConvABytes);
#pragma unroll
                        for (int k = 0; k < kConvKC / 16; ++k)     // +32 B per 16-element K step inside the swizzle atom
                            umma_bf16(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                                      (first && k == 0) ? 0u : 1u);
                        umma_commit(&empty[st]);                   // frees the ring slot when these MMAs retire
                        if (++st == static_cast<uint32_t>(p.stages)) { st = 0; ph ^= 1u; }
                    }
                }
                if (PAIR) umma_commit_2sm(&acc_full[as], 3); else umma_commit(&acc_full[as]);   // accumulator complete -> epilogue(s)
                if (FUSE11) umma_commit(&s_free[acc_it & 1u]);      // ... and the tile's scratch patch may be overwritten
                ++acc_it;
            }
            conv_trace(p, 4);                                       // last MMA issued
        }
    } else if (FUSE11 && warp >= 12) {
        // ------------------------------------------------------------ fused first layer: conv11 + bias + ReLU of the tile's
        // haloed region, one tile ahead of the MMAs.  A thread owns 8 output channels (weights in registers, packed for
        // fma.rn.f32x2) and every PL-th pixel; the arithmetic and its order are conv11_direct_kernel's, so the values are
        // bit-identical to the unfused path.
        const int tid11 = static_cast<int>(threadIdx.x) - kConvThreads;
        const int C1 = p.Cin, CG = C1 / 4, PL = kConvFuseThreads / CG;          // a thread owns 4 channels (two packed pairs)
        const int cg = tid11 % CG, pl = tid11 / CG;
        const int PW = p.BF + 2, PH = p.BT + 2, XW = p.BF + 4, XH = p.BT + 4;
        uint64_t wr2[9][2], br2[2];
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            br2[e >> 1] = pack_f32x2(p.b11[cg * 4 + e], p.b11[cg * 4 + e + 1]);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
                wr2[tap][e >> 1] = pack_f32x2(p.w11[(cg * 4 + e) * 9 + tap], p.w11[(cg * 4 + e + 1) * 9 + tap]);
        }
        // trip table (once per CTA): pixel pair -> {offset of its 3x4 input window in x_sm | tt << 16 | ff << 24, offset of its
        // two output pixels in the scratch patch}, so that a trip costs one LDS.64 instead of divisions and multiplies
        const int HW = PW >> 1, n_pairs = PH * HW;
        uint2* trip_sm = reinterpret_cast<uint2*>(x_sm + ((XH * XW + 3) & ~3));
        for (int i = tid11; i < n_pairs; i += kConvFuseThreads) {
            const int tt = i / HW, ff = 2 * (i - tt * HW);
            trip_sm[i] = make_uint2(static_cast<uint32_t>(tt * XW + ff) | (static_cast<uint32_t>(tt) << 16) | (static_cast<uint32_t>(ff) << 24),
                                    static_cast<uint32_t>((tt * PW + ff) * C1));
        }
        uint32_t fit = 0;
        for (int tile = tile0; tile < sched.n_pass0; tile += tile_step) {
            bool tile_masked;
            const ConvTile c = conv_tile_at(p, sched, tile, 0, n_mt_eff, static_cast<int>(rank), tile_masked);
            if (tile_masked) continue;
            const uint32_t buf = fit & 1u;
            if (fit >= 2) mbar_wait(&s_free[buf], ((fit >> 1) - 1u) & 1u);      // the MMAs of the tile that used this patch are done
            const int L = conv_len(p, c.b0);
            named_bar_sync(3, kConvFuseThreads);                                 // everybody is done with the previous input patch
            for (int i = tid11; i < XH * XW; i += kConvFuseThreads) {
                const int r = i / XW, q = i - r * XW;
                const int t = c.t0 - 2 + r, f = c.f0 - 2 + q;
                x_sm[i] = (t >= 0 && t < L && f >= 0 && f < p.F) ? p.x0[(static_cast<size_t>(c.b0) * p.T + t) * p.F + f] : 0.f;
            }
            named_bar_sync(3, kConvFuseThreads);
            uint16_t* sp = static_cast<uint16_t*>(p.scratch) + (static_cast<size_t>(blockIdx.x) * 2 + buf) * PH * PW * C1 + cg * 4;
            // two horizontally adjacent pixels per trip (they share 12 of their 18 input taps); PW is even.  A tile whose whole
            // haloed region lies inside the image and the utterance (most of them) skips the per-pixel checks.
            const bool interior = c.t0 >= 1 && c.t0 + p.BT < L && c.f0 >= 1 && c.f0 + p.BF < p.F;
            for (int pp = pl; pp < n_pairs; pp += PL) {
                const uint2 tr = trip_sm[pp];
                bool ok0 = true, ok1 = true;
                if (!interior) {
                    const int t = c.t0 - 1 + static_cast<int>((tr.x >> 16) & 0xffu), f = c.f0 - 1 + static_cast<int>(tr.x >> 24);
                    const bool row_ok = t >= 0 && t < L;
                    ok0 = row_ok && f >= 0 && f < p.F;
                    ok1 = row_ok && f + 1 >= 0 && f + 1 < p.F;
                }
                uint2 v0 = make_uint2(0u, 0u), v1 = v0;                          // outside: conv12's zero padding / masked rows
                if (ok0 || ok1) {
                    const float* xw = x_sm + (tr.x & 0xffffu);
                    float xv[3][4];
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int j = 0; j < 4; ++j) xv[dy][j] = xw[dy * XW + j];
                    uint64_t accA[2], accB[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) { accA[e] = br2[e]; accB[e] = br2[e]; }
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t xa = pack_f32x2(xv[dy][dx], xv[dy][dx]);
                            const uint64_t xb = pack_f32x2(xv[dy][dx + 1], xv[dy][dx + 1]);
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                accA[e] = fma_f32x2(xa, wr2[dy * 3 + dx][e], accA[e]);
                                accB[e] = fma_f32x2(xb, wr2[dy * 3 + dx][e], accB[e]);
                            }
                        }
                    float a0[4], a1[4];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        unpack_f32x2(accA[e], a0[2 * e], a0[2 * e + 1]);
                        unpack_f32x2(accB[e], a1[2 * e], a1[2 * e + 1]);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) { a0[e] = fmaxf(a0[e], 0.f); a1[e] = fmaxf(a1[e], 0.f); }
                    constexpr int A16 = ACT == 2 ? 2 : 1;
                    if (ok0) v0 = make_uint2(pack16<A16>(a0[0], a0[1]), pack16<A16>(a0[2], a0[3]));
                    if (ok1) v1 = make_uint2(pack16<A16>(a1[0], a1[1]), pack16<A16>(a1[2], a1[3]));
                }
                uint16_t* dst = sp + tr.y;
                *reinterpret_cast<uint2*>(dst) = v0;
                *reinterpret_cast<uint2*>(dst + C1) = v1;
            }
            __threadfence();                                                     // the patch is in L2 ...
            asm volatile("fence.proxy.async;" ::: "memory");                     // ... and ordered before the TMA (async proxy) reads of it
            named_bar_sync(3, kConvFuseThreads);
            if (tid11 == 0) mbar_arrive(&s_full[buf]);
            ++fit;
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (TMEM -> regs -> SMEM -> HBM)
        // 8 warps = 2 halves x 4 TMEM lane quarters.  A thread owns one output channel (its TMEM lane); the two
        // halves take alternate chunks of kConvEpiChunk output pixels.  NHWC outputs go through a per-half,
        // double-buffered [32 pixels][128 ch] bf16 staging tile so the global stores are 16-byte, fully coalesced
        // runs of 256 B per pixel; the final layer's [B,T',C*F'] output is contiguous per thread and stored directly.
        const int q = warp & 3;                                 // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;                       // 0 or 1
        const int et = (threadIdx.x - 128) & 127;               // thread index within the half
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int T = p.T, F = p.F, Cout = p.Cout, BF = p.BF, BT = p.BT;
        const int T2 = (T + 1) / 2, F2 = F / 2;
        const int OBF = p.pool ? BF / 2 : BF, OBT = p.pool ? BT / 2 : BT;   // output patch
        const int OT = p.pool ? T2 : T, OF = p.pool ? F2 : F;
        const int OPP = OBT * OBF;                              // output pixels per utterance of the patch
        const int UC = BT * BF + p.gap_cols;                    // accumulator columns per utterance of the patch (incl. gap)
        const int NO = p.BB * OPP;                              // output pixels per tile
        const float inv_obf = 1.0f / static_cast<float>(OBF), inv_opp = 1.0f / static_cast<float>(OPP);
        const int ch = q * 32 + lane;                           // channel within the 128-wide tile
        unsigned char* my_stage = stage + half * (2 * kConvEpiBytes);
        uint32_t acc_it = 0, chunk_it = 0;
        for (int pass = 0; pass < 2; ++pass)                    // pass 1 (balanced mode): the all-masked tiles, zeros only
        for (int tile = tile0; tile < (pass == 0 ? sched.n_pass0 : sched.n_pass1); tile += tile_step) {
            bool masked;
            const ConvTile c = conv_tile_at(p, sched, tile, pass, n_mt_eff, static_cast<int>(rank), masked);
            if (masked && !conv_tile_needs_zeros(p, c)) continue;           // lazy masking: nobody reads this tile
            const int n = c.m * kConvTileM + ch;
            const bool n_ok = n < Cout;
            const int ot0 = p.pool ? (c.t0 >> 1) : c.t0, of0 = p.pool ? (c.f0 >> 1) : c.f0;
            uint32_t tcol = 0;
            if (!masked) {
                const uint32_t as = ACT == 3 ? 0u : (acc_it & 1u), aph = ACT == 3 ? (acc_it & 1u) : ((acc_it >> 1) & 1u);
                mbar_wait(&acc_full[as], aph);
                tc_fence_after();
                if (threadIdx.x == 128 && acc_it == 0) conv_trace(p, 5);        // first accumulator complete
                tcol = tmem_base + lane_addr + as * static_cast<uint32_t>(p.Npad);
            }
            const float bias = (n_ok && p.bias != nullptr) ? p.bias[n] : 0.f;

            if (p.splitk > 1) {
                // split-K: this CTA's partial sums go to the workspace as they are ([pixel][Cout] fp32: a warp's 32 channels
                // are one 128-byte run); the finishing kernel adds the splits.  The halves take alternate 16-column groups.
                // The column -> (utterance, frame, bin) decomposition is carried along incrementally (one division pair per 16
                // columns, not per column: the divisions made this loop 10 us long, profiles/r2_b1_trace.txt) and the next
                // group's TMEM load is in flight while this one is stored.
                float* wsp = p.ws + static_cast<size_t>(c.split) * p.B * T * F * Cout + n;
                // (split-K runs in tap-row reuse mode without CTA pairs: the gap between two utterances of a patch is two whole
                //  rows of BF columns, so an utterance owns RPU rows of the accumulator)
                const int RPU = UC / BF;
                // 32-bit element offsets (the host only splits along K when the workspace holds < 2^31 values): a column costs
                // a predicated store and a handful of integer instructions
                auto row_off = [&](int bb, int tl, bool& ok) -> uint32_t {
                    const int b = c.b0 + bb, t = c.t0 + tl;
                    ok = bb < p.BB && tl < BT && b < p.B && t < T && n_ok;
                    return ((static_cast<uint32_t>(b) * T + t) * F + c.f0) * Cout;
                };
                auto store16 = [&](const uint32_t (&r)[16], int cg0) {
                    const int row = cg0 / BF;
                    int fl = cg0 - row * BF, bb = row / RPU, tl = row - bb * RPU;
                    bool ok;
                    uint32_t off = row_off(bb, tl, ok) + static_cast<uint32_t>(fl) * Cout;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (ok) wsp[off] = __uint_as_float(r[j]);
                        off += Cout;
                        if (++fl == BF) {                       // next accumulator row
                            fl = 0;
                            if (++tl == RPU) { tl = 0; ++bb; }
                            off = row_off(bb, tl, ok);
                        }
                    }
                };
                uint32_t ra[16], rb[16];
                int cg = half * 16;
                if (cg < p.Npad) tmem_ld_x16(tcol + cg, ra);
                while (cg < p.Npad) {
                    tc_wait_ld();
                    if (cg + 32 < p.Npad) tmem_ld_x16(tcol + cg + 32, rb);
                    store16(ra, cg);
                    cg += 32;
                    if (cg >= p.Npad) break;
                    tc_wait_ld();
                    if (cg + 32 < p.Npad) tmem_ld_x16(tcol + cg + 32, ra);
                    store16(rb, cg);
                    cg += 32;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc_it & 1u]);
                ++acc_it;
                continue;
            }

            if (p.ref_layout) {
                // pooled[b, t2, n*F2 + f2] = relu(max over the valid 2x2 window + bias)   (feature = c*F' + f, CNNs.py:88-89)
                for (int rp = half; rp < p.BB * (BT / 2); rp += 2) {
                    const int bb = rp / (BT / 2), tp = rp - bb * (BT / 2);
                    const int b = c.b0 + bb;
                    const int Lb = conv_len(p, b);
                    const int t = c.t0 + 2 * tp;
                    if (b >= p.B || t >= T) continue;           // warp-uniform
                    const bool r0_ok = !masked && t < Lb, r1_ok = !masked && (t + 1) < Lb;
                    const uint32_t col0 = tcol + static_cast<uint32_t>(bb * UC + 2 * tp * BF);
                    const size_t row = (static_cast<size_t>(b) * T2 + (t >> 1)) * (static_cast<size_t>(Cout) * F2) +
                                       static_cast<size_t>(n) * F2 + (c.f0 >> 1);
                    for (int fp0 = 0; fp0 < BF / 2; fp0 += 4) {
                        uint32_t v[4][4], v2[ACT == 3 ? 4 : 1][4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (!masked && fp0 + u < BF / 2) {  // warp-uniform
                                tmem_ld_x2(col0 + 2 * (fp0 + u), v[u][0], v[u][1]);
                                tmem_ld_x2(col0 + BF + 2 * (fp0 + u), v[u][2], v[u][3]);
                                if (ACT == 3) {                 // split mode: the cross-term accumulator
                                    tmem_ld_x2(col0 + p.Npad + 2 * (fp0 + u), v2[ACT == 3 ? u : 0][0], v2[ACT == 3 ? u : 0][1]);
                                    tmem_ld_x2(col0 + p.Npad + BF + 2 * (fp0 + u), v2[ACT == 3 ? u : 0][2], v2[ACT == 3 ? u : 0][3]);
                                }
                            }
                        }
                        if (!masked) tc_wait_ld();
                        if (ACT == 3 && !masked) {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int e = 0; e < 4; ++e) v[u][e] = __float_as_uint(__uint_as_float(v[u][e]) + __uint_as_float(v2[ACT == 3 ? u : 0][e]));
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (fp0 + u < BF / 2 && n_ok) {
                                float m = 0.f;
                                if (r0_ok) {
                                    m = fmaxf(__uint_as_float(v[u][0]), __uint_as_float(v[u][1]));
                                    if (r1_ok) m = fmaxf(m, fmaxf(__uint_as_float(v[u][2]), __uint_as_float(v[u][3])));
                                    m = fmaxf(m + bias, 0.f);
                                }
                                if (p.y_f32) conv_store<true, ACT>(p.y, row + fp0 + u, m); else conv_store<false, ACT>(p.y, row + fp0 + u, m);
                            }
                        }
                    }
                }
                if (!masked) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_remote(&acc_empty[ACT == 3 ? 0u : (acc_it & 1u)], 0); else mbar_arrive(&acc_empty[ACT == 3 ? 0u : (acc_it & 1u)]); }
                    ++acc_it;
                }
                continue;
            }

            // ---------------- NHWC bf16 output: this half's chunks of kConvEpiChunk output pixels
            const int n_chunks = (NO + kConvEpiChunk - 1) / kConvEpiChunk;
            constexpr int kParts = ACT == 3 ? 2 : 1;                // split output: the chunk is staged and stored twice (hi, lo)
            auto cvt_out = [](float v, int part) -> uint16_t {
                if (ACT != 3) return cvt16_bits<ACT == 3 ? 1 : ACT>(v);
                const uint16_t hi = cvt_bf16_bits(v);
                return part == 0 ? hi : cvt_bf16_bits(v - __uint_as_float(static_cast<uint32_t>(hi) << 16));
            };
            for (int ck = half; ck < n_chunks; ck += 2)
            for (int part = 0; part < kParts; ++part) {
                const int o0 = ck * kConvEpiChunk;
                const int cnt = min(kConvEpiChunk, NO - o0);
                unsigned char* buf = my_stage + (chunk_it & 1u) * kConvEpiBytes;
                if (!masked) {
                    // phase 1: this thread's channel of `cnt` output pixels -> staging[pixel][ch]
                    uint16_t* dst = reinterpret_cast<uint16_t*>(buf) + ch;
                    if (!p.pool) {
                        // (Tried: both 16-column loads of a chunk issued before one wait -- conv21, the layer whose epilogue is
                        //  as long as its K = 1152 main loop, went from 1228 to 1128 TFLOP/s: the two warps of a scheduler then
                        //  wait and convert in lockstep instead of covering each other.)
                        // output o of utterance bb sits in accumulator column o + bb * gap_cols
#pragma unroll
                        for (int g16 = 0; g16 < kConvEpiChunk / 16; ++g16) {
                            const int oa = o0 + g16 * 16;
                            if (g16 * 16 < cnt) {                                   // warp-uniform
                                uint32_t r[16];
                                const int bba = __float2int_rz((static_cast<float>(oa) + 0.5f) * inv_opp);
                                const int bbz = __float2int_rz((static_cast<float>(min(oa + 15, NO - 1)) + 0.5f) * inv_opp);
                                const int cola = oa + bba * p.gap_cols;
                                if (bba == bbz && cola + 16 <= p.Npad) {
                                    tmem_ld_x16(tcol + cola, r);
                                } else {                        // group straddles an utterance boundary (or the accumulator's end)
#pragma unroll
                                    for (int j = 0; j < 16; ++j) {
                                        const int bbj = __float2int_rz((static_cast<float>(min(oa + j, NO - 1)) + 0.5f) * inv_opp);
                                        tmem_ld_x1(tcol + min(oa + j + bbj * p.gap_cols, p.Npad - 1), r[j]);
                                    }
                                }
                                if (ACT == 3) {                 // split mode: add the cross-term accumulator
                                    uint32_t r2[16];
                                    if (bba == bbz && cola + 16 <= p.Npad) {
                                        tmem_ld_x16(tcol + p.Npad + cola, r2);
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 16; ++j) {
                                            const int bbj = __float2int_rz((static_cast<float>(min(oa + j, NO - 1)) + 0.5f) * inv_opp);
                                            tmem_ld_x1(tcol + p.Npad + min(oa + j + bbj * p.gap_cols, p.Npad - 1), r2[j]);
                                        }
                                    }
                                    tc_wait_ld();
#pragma unroll
                                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                                }
                                tc_wait_ld();
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (g16 * 16 + j < cnt)
                                        dst[(g16 * 16 + j) * kConvTileM] = cvt_out(DGRAD ? __uint_as_float(r[j]) + bias : fmaxf(__uint_as_float(r[j]) + bias, 0.f), part);
                            }
                        }
                    } else {
                        // output o -> (bb, tp, fp); window columns (bb*BT + 2tp)*BF + 2fp (+1, +BF, +BF+1)
                        int bb = __float2int_rz((static_cast<float>(o0) + 0.5f) * inv_opp);
                        int rem = o0 - bb * OPP;
                        int tp = __float2int_rz((static_cast<float>(rem) + 0.5f) * inv_obf), fp = rem - tp * OBF;
                        int Lcur = conv_len(p, c.b0 + bb);
                        // Wide path: when a pooled row of the patch holds a multiple of 8 (4) outputs, the two accumulator rows of
                        // 8 (4) consecutive windows are 16 (8) consecutive columns each -- two TMEM loads instead of sixteen
                        // (eight), and one index decomposition per group.  (Short-K layers are epilogue-bound: conv12.)
                        constexpr bool kWide = ACT != 3 && !FUSE11;      // (those variants have no registers to spare)
                        const int wide = (!kWide || !p.wide_epi) ? 0 : ((OBF & 7) == 0 ? 8 : ((OBF & 3) == 0 ? 4 : 0));
                        if (kWide && wide == 8) {
                            for (int j0 = 0; j0 < cnt; j0 += 8) {       // o0 and cnt are multiples of 8 here (OPP is)
                                const uint32_t col = tcol + static_cast<uint32_t>(bb * UC + 2 * tp * BF + 2 * fp);
                                const bool r1 = (c.t0 + 2 * tp + 1) < Lcur;
                                uint32_t ra[16], rb[16];
                                tmem_ld_x16(col, ra);
                                tmem_ld_x16(col + BF, rb);
                                tc_wait_ld();
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    float m = fmaxf(__uint_as_float(ra[2 * u]), __uint_as_float(ra[2 * u + 1]));
                                    if (r1) m = fmaxf(m, fmaxf(__uint_as_float(rb[2 * u]), __uint_as_float(rb[2 * u + 1])));
                                    dst[(j0 + u) * kConvTileM] = cvt_out(fmaxf(m + bias, 0.f), part);
                                }
                                fp += 8;
                                if (fp == OBF) { fp = 0; if (++tp == OBT) { tp = 0; ++bb; Lcur = conv_len(p, c.b0 + bb); } }
                            }
                        } else if (kWide && wide == 4) {
                            for (int j0 = 0; j0 < cnt; j0 += 4) {
                                const uint32_t col = tcol + static_cast<uint32_t>(bb * UC + 2 * tp * BF + 2 * fp);
                                const bool r1 = (c.t0 + 2 * tp + 1) < Lcur;
                                uint32_t ra[8], rb[8];
                                tmem_ld_x8(col, ra);
                                tmem_ld_x8(col + BF, rb);
                                tc_wait_ld();
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    float m = fmaxf(__uint_as_float(ra[2 * u]), __uint_as_float(ra[2 * u + 1]));
                                    if (r1) m = fmaxf(m, fmaxf(__uint_as_float(rb[2 * u]), __uint_as_float(rb[2 * u + 1])));
                                    dst[(j0 + u) * kConvTileM] = cvt_out(fmaxf(m + bias, 0.f), part);
                                }
                                fp += 4;
                                if (fp == OBF) { fp = 0; if (++tp == OBT) { tp = 0; ++bb; Lcur = conv_len(p, c.b0 + bb); } }
                            }
                        } else
                        for (int j0 = 0; j0 < cnt; j0 += 4) {
                            uint32_t v[4][4], v2[ACT == 3 ? 4 : 1][4];
                            bool r1[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (j0 + u < cnt) {             // warp-uniform
                                    const uint32_t col = tcol + static_cast<uint32_t>(bb * UC + 2 * tp * BF + 2 * fp);
                                    r1[u] = (c.t0 + 2 * tp + 1) < Lcur;     // ceil-mode / masked second row
                                    tmem_ld_x2(col, v[u][0], v[u][1]);
                                    tmem_ld_x2(col + BF, v[u][2], v[u][3]);
                                    if (ACT == 3) {             // split mode: the cross-term accumulator
                                        tmem_ld_x2(col + p.Npad, v2[ACT == 3 ? u : 0][0], v2[ACT == 3 ? u : 0][1]);
                                        tmem_ld_x2(col + p.Npad + BF, v2[ACT == 3 ? u : 0][2], v2[ACT == 3 ? u : 0][3]);
                                    }
                                    if (++fp == OBF) { fp = 0; if (++tp == OBT) { tp = 0; ++bb; Lcur = conv_len(p, c.b0 + bb); } }
                                }
                            }
                            tc_wait_ld();
                            if (ACT == 3) {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    if (j0 + u < cnt)
#pragma unroll
                                        for (int e = 0; e < 4; ++e) v[u][e] = __float_as_uint(__uint_as_float(v[u][e]) + __uint_as_float(v2[ACT == 3 ? u : 0][e]));
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (j0 + u < cnt) {
                                    float m = fmaxf(__uint_as_float(v[u][0]), __uint_as_float(v[u][1]));
                                    if (r1[u]) m = fmaxf(m, fmaxf(__uint_as_float(v[u][2]), __uint_as_float(v[u][3])));
                                    dst[(j0 + u) * kConvTileM] = cvt_out(fmaxf(m + bias, 0.f), part);   // max and +bias/ReLU commute
                                }
                            }
                        }
                    }
                    named_bar_sync(1 + half, 128);
                }
                // phase 2: 16 threads per pixel copy 256 B runs to y (zeros for masked frames)
                {
                    const int seg = et & 15;
                    const bool seg_ok = c.m * kConvTileM + seg * 8 < Cout;
#pragma unroll
                    for (int i = 0; i < kConvEpiChunk / 8; ++i) {
                        const int po = (et >> 4) + 8 * i;
                        if (po < cnt) {
                            const int o = o0 + po;
                            const int bb = __float2int_rz((static_cast<float>(o) + 0.5f) * inv_opp);
                            const int rem = o - bb * OPP;
                            const int tl = __float2int_rz((static_cast<float>(rem) + 0.5f) * inv_obf), fl = rem - tl * OBF;
                            const int b = c.b0 + bb, to = ot0 + tl, fo = of0 + fl;
                            if (b < p.B && to < OT && seg_ok) {
                                const int t_in = p.pool ? 2 * to : to;
                                uint4 val = make_uint4(0u, 0u, 0u, 0u);
                                if (!masked && t_in < conv_len(p, b))
                                    val = *reinterpret_cast<const uint4*>(buf + po * (kConvTileM * 2) + seg * 16);
                                const size_t yoff = ((static_cast<size_t>(b) * OT + to) * OF + fo) * (kParts * Cout) + part * Cout + c.m * kConvTileM + seg * 8;
                                if (DGRAD && p.mask != nullptr) {    // fused ReLU backward: keep the gradient where the activation was > 0
                                    const uint4 mv = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.mask) + yoff);
                                    val.x &= bf16x2_positive_mask(mv.x); val.y &= bf16x2_positive_mask(mv.y);
                                    val.z &= bf16x2_positive_mask(mv.z); val.w &= bf16x2_positive_mask(mv.w);
                                }
                                *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y) + yoff) = val;
                            }
                        }
                    }
                }
                if (!masked) ++chunk_it;
            }
            if (!masked) {                                      // all of this warp's TMEM reads of the tile are done
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_remote(&acc_empty[ACT == 3 ? 0u : (acc_it & 1u)], 0); else mbar_arrive(&acc_empty[ACT == 3 ? 0u : (acc_it & 1u)]); }
                ++acc_it;
            }
        }
    }
    if (threadIdx.x == 128) conv_trace(p, 6);               // epilogue (first warp) done
    tc_fence_before();
    if (PAIR) cluster_sync_all();       // nobody leaves (or frees TMEM) while the peer may still signal into this CTA
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_2sm(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (threadIdx.x == 0) conv_trace(p, 7);                 // CTA end
}

// Second half of a split-K launch: y = format(relu(max over the pool window (sum over splits of ws) + bias)).
// ws [S][B,T,F,Cout] fp32.  A lane owns 4 channels of ONE input pixel: with pooling the four positions of an output's 2x2 window
// sit in four neighbouring lanes and meet through two shuffles, and the S loads of a lane are issued four at a time -- the
// kernel is a handful of dependent L2 round trips long (it was 16 of them, ~8 us, with one thread per output and a loop over
// window x splits; profiles/r2_b1_trace.txt).  The sum over the splits runs in split order: deterministic.
// ACT: 0 = fp32, 1 = bf16, 2 = fp16 output.
template <int ACT>
__global__ void __launch_bounds__(256)
conv_splitk_finish_kernel(const float* __restrict__ ws, int S, const float* __restrict__ bias, void* __restrict__ y,
                          int B, int T, int F, int Cout, int pool, int ref, int relu) {
    griddep_launch();
    griddep_wait();
    const int OT = pool ? (T + 1) / 2 : T, OF = pool ? F / 2 : F, C4 = Cout / 4;
    const int W = pool ? 4 : 1;
    const size_t total = static_cast<size_t>(B) * OT * OF * C4 * W;          // a multiple of 4 when pooling: quads never straddle the end
    const size_t split_stride = static_cast<size_t>(B) * T * F * Cout;
    const int lane = threadIdx.x & 31;
    for (size_t base = static_cast<size_t>(blockIdx.x) * blockDim.x + (threadIdx.x - lane); base < total; base += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t i = base + lane;
        const bool live = i < total;
        const int w = pool ? static_cast<int>(i & 3) : 0;
        size_t r = live ? (pool ? i >> 2 : i) : 0;
        const int c4 = static_cast<int>(r % C4); r /= C4;
        const int fo = static_cast<int>(r % OF); r /= OF;
        const int to = static_cast<int>(r % OT);
        const int b = static_cast<int>(r / OT);
        const int t = pool ? 2 * to + (w >> 1) : to, f = pool ? 2 * fo + (w & 1) : fo;
        float4 a = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (live && t < T) {                                                  // ceil mode: the window's second row may not exist
            const float* src = ws + ((static_cast<size_t>(b) * T + t) * F + f) * Cout + c4 * 4;
            a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s0 = 0; s0 < S; s0 += 4) {
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (s0 + k < S) v[k] = *reinterpret_cast<const float4*>(src + (s0 + k) * split_stride);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (s0 + k < S) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
            }
        }
        if (pool) {
#pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                a.x = fmaxf(a.x, __shfl_xor_sync(0xffffffffu, a.x, off));
                a.y = fmaxf(a.y, __shfl_xor_sync(0xffffffffu, a.y, off));
                a.z = fmaxf(a.z, __shfl_xor_sync(0xffffffffu, a.z, off));
                a.w = fmaxf(a.w, __shfl_xor_sync(0xffffffffu, a.w, off));
            }
        }
        if (!live || w != 0) continue;
        const float4 bv = *reinterpret_cast<const float4*>(bias + c4 * 4);
        float o[4] = {a.x + bv.x, a.y + bv.y, a.z + bv.z, a.w + bv.w};
        if (relu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
        }
        if (ref) {                                                            // [B,T',C*F'] with feature = c*F' + f (CNNs.py:88-89)
            const size_t row = (static_cast<size_t>(b) * OT + to) * (static_cast<size_t>(Cout) * OF);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const size_t idx = row + static_cast<size_t>(c4 * 4 + e) * OF + fo;
                if (ACT == 0) static_cast<float*>(y)[idx] = o[e];
                else static_cast<uint16_t*>(y)[idx] = cvt16_bits<ACT == 0 ? 1 : ACT>(o[e]);
            }
        } else {
            const size_t idx = ((static_cast<size_t>(b) * OT + to) * OF + fo) * Cout + c4 * 4;
            if (ACT == 0) *reinterpret_cast<float4*>(static_cast<float*>(y) + idx) = make_float4(o[0], o[1], o[2], o[3]);
            else *reinterpret_cast<uint2*>(static_cast<uint16_t*>(y) + idx) =
                     make_uint2(pack16<ACT == 0 ? 1 : ACT>(o[0], o[1]), pack16<ACT == 0 ? 1 : ACT>(o[2], o[3]));
        }
    }
}

// ---------------------------------------------------------------------------------- host side
struct ConvPlan {
    int BF, BT, BB, N, Npad;
    double cost;
};

// Pick the patch shape that minimises the modelled time of the whole layer.  Per 16-deep K step a 128 x Npad
// MMA costs Npad/2 tensor cycles; the SM can ingest ~64 B/clk from L2 (measured: every layer plateaus at
// ~15 TB/s chip-wide; layers needing > 55 B/clk already lose tensor time), so a step also costs its operand bytes / 52.  `halo` = 2 in tap-row reuse mode (patches carry
// +-1 frame and utterances inside a patch are separated by 2 halo rows of accumulator columns).
static ConvPlan conv_plan(int B, int T, int F, int Cin, int Cout, bool pool, int halo, bool pair, int sms, bool ragged, int nmax = 256, int bbmax = 1 << 30,
                          int ksplit = 1) {
    // Cin here = the contraction depth per tap (3 * Cin in split mode)
    ConvPlan best{0, 0, 0, 0, 0, 1e300};
    const double ksteps = 9.0 * Cin / 16.0;
    for (int BF = 2; BF <= F && BF <= 256; BF += 2) {
        if (F % BF != 0) continue;
        const int bt_max = 256 / BF;
        for (int BT = pool ? 2 : 1; BT <= bt_max; BT += pool ? 2 : 1) {
            if (BT > T + 1 && BT > 2) break;
            const int n_tt = (T + BT - 1) / BT;
            const int bb_max = 256 / (BF * BT);          // several utterances per patch when one utterance's rows leave room
            for (int BB = 1; BB <= bb_max && BB <= B && BB <= bbmax; ++BB) {
                int N = (BB - 1) * (BT + halo) * BF + BT * BF;            // accumulator columns incl. halo gaps
                double b_rows = halo ? BB * (BT + 2.0) * BF / 3.0 : BB * BT * BF;   // activation rows fetched per tap (per CTA)
                if (pair) {
                    // CTA pairs: the patch is cut in two equal halves (along t, or one utterance each); every half
                    // carries its own halo, so there are no gap columns, and N must split on an 8-row boundary
                    if (!((BB == 1 && BT % 2 == 0) || BB == 2)) continue;
                    if (BB == 1) { N = BT * BF; if (N % 16 != 0) continue; }      // halves split along t: no room for padding
                    else N = 2 * ((BT * BF + 7) / 8 * 8);                          // one utterance per CTA, each half padded to 8 rows
                    const int RT = BB == 1 ? BT / 2 : BT;
                    b_rows = (RT + 2.0) * BF / 3.0;
                }
                const int Npad = (N + 15) / 16 * 16;
                if (Npad > nmax) continue;                   // 2 (double buffer) x accumulators per tile x Npad <= 512 TMEM columns
                // tiles are dealt to the SMs (or SM pairs) in whole waves: with few tiles (small batches) a smaller patch
                // that fills more SMs wins even though each of its MMAs is less efficient
                const int n_mt = (Cout + kConvTileM - 1) / kConvTileM;
                const double tiles = static_cast<double>(F / BF) * n_tt * ((B + BB - 1) / BB) * (pair ? n_mt / 2 : n_mt) * ksplit;
                const double units = pair ? sms / 2 : sms;
                const double waves = ceil(tiles / units);
                const double ingest = (128.0 + b_rows) * 32.0 / 52.0;                   // bytes per 16-deep step / ~52 B/clk effective
                double step = Npad / 2.0 > ingest ? Npad / 2.0 : ingest;
                if (step < 40.0) step = 40.0;                                           // issue + operand-fetch floor of one MMA
                double cost = waves * (step * ksteps / ksplit + 700.0);
                // ragged batches: the tile that straddles an utterance's end computes rows that are masked afterwards
                // (half a tile height per utterance on average, against ~3/4 of the padded length in valid rows)
                if (ragged) cost *= 1.0 + (0.5 * BT) / (0.75 * T + 1.0);
                if (cost < best.cost) best = ConvPlan{BF, BT, BB, N, Npad, cost};
            }
        }
    }
    return best;
}

}  // namespace dasv

using namespace dasv;

// ---- launch cache: the patch plan, both tensor maps (they embed the x / packed-weight addresses) and the launch shape of
// every (addresses, shape, flags, tuning knobs) seen, so that a steady-state call is a lookup + one launch.  PyTorch's
// caching allocator hands the same activation buffers back step after step, so a model's steps hit after the first one.
struct ConvKey {
    const void* x; const void* wp;
    int B, T, F, Cin, Cout, flags, y_dtype, dgrad, dev;   // flags include the operand-format bits
    int env_reuse, env_pair, env_sb, ragged;   // ragged: lengths given (the plan then prefers low tiles)
    int env_nosplit;                           // DASV_CONV_NOSPLITK=1: never split along K (bit-identical results at every batch size)
    int env_nowide;                            // DASV_CONV_NOWIDE=1: pooled epilogue with the narrow TMEM loads (A/B runs)
    int env_split;                             // DASV_CONV_SPLITK=n: force this split factor where the launch may split (tuning)
    char env_plan[16];
};
struct ConvEntry {
    ConvKey key;
    CUtensorMap tmA, tmB;
    ConvParams p;
    size_t smem, ws_bytes;
    int grid, pair;
    unsigned long long stamp;
};
static std::mutex g_conv_mu;
static unsigned long long* g_conv_trace = nullptr;   // dasv_debug_conv_trace
static ConvEntry g_conv_cache[64];
static int g_conv_n = 0;
static unsigned long long g_conv_clock = 0;

static bool conv_key_eq(const ConvKey& a, const ConvKey& b) { return memcmp(&a, &b, sizeof(ConvKey)) == 0; }

// The dynamic shared memory limit of a kernel variant is raised once per device (to the architectural 227 KB).
typedef void (*ConvKernelFn)(const CUtensorMap, const CUtensorMap, const ConvParams);
// variant = pair * 4 + {0: forward bf16 out, 1: input-gradient pass (bf16), 2: forward fp16 out, 3: forward split-bf16 out}
static ConvKernelFn conv_kernel_variant(int variant) {
    switch (variant) {
        case 0: return conv3x3_igemm_kernel<false, false, 1>;
        case 1: return conv3x3_igemm_kernel<false, true, 1>;
        case 2: return conv3x3_igemm_kernel<false, false, 2>;
        case 3: return conv3x3_igemm_kernel<false, false, 3>;
        case 4: return conv3x3_igemm_kernel<true, false, 1>;
        case 5: return conv3x3_igemm_kernel<true, true, 1>;
        case 6: return conv3x3_igemm_kernel<true, false, 2>;
        case 7: return conv3x3_igemm_kernel<true, false, 3>;
        case 8: return conv3x3_igemm_kernel<false, false, 1, true>;     // conv11 fused in front (bf16 / fp16 activations)
        default: return conv3x3_igemm_kernel<false, false, 2, true>;
    }
}
static int conv_raise_smem(int variant, int dev) {
    static bool done[10][64] = {};
    if (dev >= 0 && dev < 64 && done[variant][dev]) return 0;
    const int kMax = 227 * 1024;
    cudaError_t e = cudaFuncSetAttribute(conv_kernel_variant(variant), cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
    if (e != cudaSuccess) { set_error("conv3x3_igemm_bf16: smem attribute: %s", cudaGetErrorString(e)); return 1; }
    if (dev >= 0 && dev < 64) done[variant][dev] = true;
    return 0;
}

// Plan + tensor maps of one (x, wp, shape, flags) combination.  Returns 0 and fills `en` (everything except the per-call
// pointers bias / lengths / mask / y), or 1 with the error set.
static int conv_build_entry(ConvEntry& en, const ConvKey& k) {
    const void* x = k.x; const void* wp = k.wp;
    const int B = k.B, T = k.T, F = k.F, Cin = k.Cin, Cout = k.Cout, flags = k.flags;
    const bool pool = (flags & 2) != 0, ref = (flags & 4) != 0;
    EncodeTiledFn encode = get_encode_tiled();
    if (!encode) { set_error("conv3x3_igemm_bf16: cuTensorMapEncodeTiled is not available from the CUDA driver"); return 1; }

    // tap-row reuse is the default; DASV_CONV_REUSE=0 selects one TMA box per tap (A/B comparisons, debugging)
    const int reuse = k.env_reuse;
    const int cout_pad = (Cout + kConvTileM - 1) / kConvTileM * kConvTileM;
    // CTA pairs (cta_group::2, 256 channels per pair) need an even number of 128-channel tiles; DASV_CONV_PAIR=0/1 overrides
    int pair = (flags & 8) != 0;                                 // DASV_CONV_PAIR
    if (k.env_pair >= 0) pair = k.env_pair;
    if (!reuse || cout_pad % (2 * kConvTileM) != 0) pair = 0;
    const int sms = sm_count();
    const bool x3 = (flags & 64) != 0;                           // split operands (fp32x3 mode)
    const int Kc = x3 ? 3 * Cin : Cin;                           // contraction depth per tap
    const int Cx = x3 ? 2 * Cin : Cin;                           // channels of the x tensor
    const int nmax = 256;                                        // split mode: two accumulators per tile, tiles not double-buffered
    const bool fused = (flags & 128) != 0;                       // conv11 computed in front by the kernel itself: one utterance per patch
    if (fused) pair = 0;
    ConvPlan pl = conv_plan(B, T, F, Kc, Cout, pool, reuse ? 2 : 0, pair != 0, sms, k.ragged != 0, nmax, fused ? 1 : (1 << 30));
    if (pair && pl.N == 0) { pair = 0; pl = conv_plan(B, T, F, Kc, Cout, pool, 2, false, sms, k.ragged != 0, nmax); }
    if (k.env_plan[0]) {                                         // "BF,BT,BB" tuning override (scripts/bench_conv_layers.py)
        int bf = 0, bt = 0, bb = 0;
        if (sscanf(k.env_plan, "%d,%d,%d", &bf, &bt, &bb) == 3 && bf > 0 && F % bf == 0 && bf % 2 == 0 && bt > 0 && (!pool || bt % 2 == 0) && bb > 0) {
            int n = (bb - 1) * (bt + (reuse ? 2 : 0)) * bf + bt * bf;
            bool ok = true;
            if (pair) { ok = (bb == 1 && bt % 2 == 0 && (bt * bf) % 16 == 0) || bb == 2; n = bb == 1 ? bt * bf : 2 * ((bt * bf + 7) / 8 * 8); }
            if (fused && bb != 1) ok = false;
            if (ok && (n + 15) / 16 * 16 <= nmax) pl = ConvPlan{bf, bt, bb, n, (n + 15) / 16 * 16, 0.0};
        }
    }
    // split-K (small batches): when the best plan leaves most SMs idle, deal the K slices of every tile to several CTAs and
    // finish in a second, streaming kernel.  Compared through the same cost model (+ the finishing pass).
    int splitk = 1, kpc = Kc / kConvKC;
    if (reuse && !x3 && !fused && !k.ragged && (flags & 1) && !k.env_plan[0] && Cout % 4 == 0 && !k.env_nosplit && pl.N != 0) {
        const int kch = Kc / kConvKC;
        // the cheapest split by the model (+ its finishing pass), taken when it beats the unsplit launch (single-CTA tiles) by 10 %;
        // threshold and the per-split term fitted to scripts/ubench/splitk_sweep.py (B = 1..8, profiles/r2_splitk_sweep.txt)
        const ConvPlan base = pair ? conv_plan(B, T, F, Kc, Cout, pool, 2, false, sms, false, nmax) : pl;
        double best = 0.9 * (base.N ? (base.cost < pl.cost ? base.cost : pl.cost) : pl.cost);
        for (int sreq = 2; sreq <= 16 && sreq <= kch; sreq *= 2) {
            const int kp = (kch + sreq - 1) / sreq, seff = (kch + kp - 1) / kp;
            if (seff < 2 || (k.env_split > 1 && seff != k.env_split)) continue;
            ConvPlan ps = conv_plan(B, T, F, Kc, Cout, pool, 2, false, sms, false, nmax, 1 << 30, seff);
            if (ps.N == 0) continue;
            // finishing pass: the partials are written and read once (L2-resident at these sizes), ~64 B/clk per SM, + a launch
            const double fin = 4000.0 + 500.0 * seff + 2.0 * seff * static_cast<double>(B) * T * F * Cout * 4.0 / (64.0 * sms);
            if (static_cast<double>(seff) * B * T * F * Cout >= 2147483648.0) continue;   // the raw epilogue indexes the workspace with 32 bits
            if (ps.cost + fin < best || (k.env_split > 1 && splitk == 1)) { best = ps.cost + fin; pl = ps; splitk = seff; kpc = kp; pair = 0; }
        }
    }
    if (pl.N == 0) { set_error("conv3x3_igemm_bf16: no patch shape for T=%d F=%d", T, F); return 1; }
    const int pair_rt = pl.BB == 1 ? pl.BT / 2 : pl.BT;           // frames per CTA half in pair mode
    const int box_t = pair ? pair_rt + 2 : pl.BT + (reuse ? 2 : 0);
    const int box_b = pair ? 1 : pl.BB;

    {
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(9) * Kc, static_cast<cuuint64_t>(cout_pad)};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(9) * Kc * 2};
        const cuuint32_t box[2] = {kConvKC, kConvTileM};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&en.tmA, (flags & 16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_igemm_bf16: weight tensor map encode failed (%d)", static_cast<int>(r)); return 1; }
    }
    {
        // fused first layer: the "activation tensor" is the scratch, [slots][BT+2][BF+2][Cin] (k.x points at it)
        const cuuint64_t dF = fused ? pl.BF + 2 : F, dT = fused ? pl.BT + 2 : T, dB = fused ? 2 * static_cast<cuuint64_t>(sms) : B;
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(Cx), dF, dT, dB};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(Cx) * 2, dF * Cx * 2, dT * dF * Cx * 2};
        const cuuint32_t box[4] = {kConvKC, static_cast<cuuint32_t>(pl.BF), static_cast<cuuint32_t>(box_t),
                                   static_cast<cuuint32_t>(box_b)};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&en.tmB, (flags & 32) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_igemm_bf16: activation tensor map encode failed (%d)", static_cast<int>(r)); return 1; }
    }

    ConvParams p{};
    p.B = B; p.T = T; p.F = F; p.Cin = Cin; p.Cout = Cout;
    p.BF = pl.BF; p.BT = pl.BT; p.BB = pl.BB; p.N = pl.N; p.Npad = pl.Npad;
    p.n_ft = F / pl.BF; p.n_tt = (T + pl.BT - 1) / pl.BT; p.n_bt = (B + pl.BB - 1) / pl.BB; p.n_mt = cout_pad / kConvTileM;
    p.kchunks = Kc / kConvKC;
    p.a_tap_stride = Kc;
    p.kcx_wrap = x3 ? 2 * Cin / kConvKC : 0x7fffffff;
    p.pool = pool; p.ref_layout = ref; p.y_f32 = (k.y_dtype == 0);
    p.relu = (flags & 1) ? 1 : 0;
    p.lazy_mask = (flags & 256) ? 1 : 0;
    p.wide_epi = k.env_nowide ? 0 : 1;
    p.w_f16 = (flags & 16) ? 1 : 0; p.x_f16 = (flags & 32) ? 1 : 0;
    p.balanced = (k.ragged && p.n_bt <= kConvMaxGroups && !getenv("DASV_CONV_UNBALANCED")) ? 1 : 0;
    p.splitk = splitk; p.kpc = kpc;
    en.ws_bytes = splitk > 1 ? static_cast<size_t>(splitk) * B * T * F * Cout * sizeof(float) : 0;
    p.reuse = reuse;
    p.gap_cols = pair ? (pl.BB == 2 ? pl.Npad / 2 - pl.BT * pl.BF : 0) : (reuse ? 2 * pl.BF : 0);
    p.pair = pair; p.RT = pair_rt; p.split_t = pl.BB == 1;
    p.b_bytes = static_cast<uint32_t>(box_b) * box_t * pl.BF * 128u;
    // staging + alignment slack + barriers (+ the valid-tile prefix table of a ragged batch) (+ the fused first layer's
    // barriers and input patch)
    const uint32_t kTable = p.balanced ? ((static_cast<uint32_t>(p.n_bt) + 1u) * 4u + 15u) / 16u * 16u : 0u;
    const uint32_t kFuse = fused ? (32u + ((static_cast<uint32_t>(pl.BT + 4) * (pl.BF + 4) + 3u) & ~3u) * 4u      // barriers + input patch
                                    + static_cast<uint32_t>(pl.BT + 2) * ((pl.BF + 2) / 2) * 8u + 15u) / 16u * 16u : 0u;   // + trip table
    p.fuse_off = 512u + kTable;
    const uint32_t kFixed = 4 * kConvEpiBytes + 1024 + 512 + kTable + kFuse;
    const uint32_t kAvail = 227u * 1024u - kFixed;
    if (reuse) {
        // ring 1 = activation patches (+ the rows a 16-padded, 2-frame-shifted MMA view may touch), ring 2 = weight tiles
        const uint32_t view_rows = pair ? static_cast<uint32_t>(pl.Npad) / 2u : static_cast<uint32_t>(pl.Npad);   // rows one CTA's MMA view spans
        p.stage_bytes = ((view_rows + 2u * pl.BF) * 128u + 1023u) & ~1023u;
        if (p.stage_bytes < p.b_bytes) p.stage_bytes = (p.b_bytes + 1023u) & ~1023u;
        int sb = 3;
        int sa = (static_cast<int>(kAvail) - sb * static_cast<int>(p.stage_bytes)) / static_cast<int>(kConvABytes);
        if (sa < 5) { sb = 2; sa = (static_cast<int>(kAvail) - sb * static_cast<int>(p.stage_bytes)) / static_cast<int>(kConvABytes); }
        if (sa > 12) sa = 12;
        if (sa < 3) { set_error("conv3x3_igemm_bf16: rings do not fit shared memory"); return 1; }
        if (k.env_sb >= 2 && k.env_sb <= 4) { sb = k.env_sb; sa = (static_cast<int>(kAvail) - sb * static_cast<int>(p.stage_bytes)) / static_cast<int>(kConvABytes); if (sa > 12) sa = 12; }
        p.stages = sb; p.sa = sa;
    } else {
        p.stage_bytes = kConvABytes + ((static_cast<uint32_t>(pl.Npad) * 128u + 1023u) & ~1023u);
        int stages = static_cast<int>(kAvail / p.stage_bytes);
        if (stages > 8) stages = 8;
        if (stages < 2) { set_error("conv3x3_igemm_bf16: ring does not fit shared memory"); return 1; }
        p.stages = stages; p.sa = 0;
    }
    uint32_t cols = 32;
    while (cols < 2u * pl.Npad) cols <<= 1;
    p.tmem_cols = cols;
    if (getenv("DASV_CONV_DEBUG"))
        fprintf(stderr, "conv plan: B=%d T=%d F=%d Cin=%d Cout=%d pool=%d reuse=%d pair=%d BF=%d BT=%d BB=%d N=%d Npad=%d stages=%d sa=%d stage_bytes=%u b_bytes=%u splitk=%d\n",
                B, T, F, Cin, Cout, (int)pool, reuse, pair, pl.BF, pl.BT, pl.BB, pl.N, pl.Npad, p.stages, p.sa, p.stage_bytes, p.b_bytes, splitk);
    const long long n_tiles = static_cast<long long>(pair ? p.n_mt / 2 : p.n_mt) * p.n_ft * p.n_tt * p.n_bt * splitk;
    if (n_tiles > 0x7fffffffLL) { set_error("conv3x3_igemm_bf16: too many tiles"); return 1; }
    en.smem = static_cast<size_t>(p.stages) * p.stage_bytes + static_cast<size_t>(p.sa) * kConvABytes + kFixed;
    en.grid = pair ? 2 * static_cast<int>(n_tiles < sms / 2 ? n_tiles : sms / 2) : static_cast<int>(n_tiles < sms ? n_tiles : sms);
    en.pair = pair;
    en.p = p;
    en.key = k;
    return 0;
}

struct ConvFront {            // the fused first layer's own operands (flags bit 128)
    const float* x0; const float* w11; const float* b11;
};

static int conv_igemm_launch(const void* x, const void* wp, const float* bias, const int32_t* lengths, const void* mask,
                             void* y, int y_dtype, int flags,
                             int B, int T, int F, int Cin, int Cout, void* stream, const ConvFront* front = nullptr,
                             void* workspace = nullptr, size_t* ws_query = nullptr) {
    if (!ws_query && (!x || !wp || !y)) { set_error("conv3x3_igemm_bf16: null argument"); return 1; }
    if (Cin % kConvKC != 0 || Cin <= 0) { set_error("conv3x3_igemm_bf16: Cin=%d must be a positive multiple of 64", Cin); return 1; }
    if (Cout <= 0 || Cout % 8 != 0) { set_error("conv3x3_igemm_bf16: Cout=%d must be a positive multiple of 8", Cout); return 1; }
    if (F <= 0 || F % 2 != 0 || F > 256) { set_error("conv3x3_igemm_bf16: F=%d must be even and <= 256", F); return 1; }
    const bool pool = (flags & 2) != 0, ref = (flags & 4) != 0;
    if (!(flags & 1) && pool) { set_error("conv3x3_igemm_bf16: the pooled epilogue always applies ReLU (drop DASV_CONV_POOL or set DASV_CONV_RELU)"); return 1; }
    if (ref && !pool) { set_error("conv3x3_igemm_bf16: REF_LAYOUT requires POOL"); return 1; }
    const int act = (flags & 32) ? 2 : 1;                        // 16-bit activation format: the output's follows the input's
    if (!ref && y_dtype != act) { set_error("conv3x3_igemm_bf16: NHWC output must have the input's 16-bit format (dtype %d)", act); return 1; }
    if (y_dtype != 0 && y_dtype != act) { set_error("conv3x3_igemm_bf16: bad y dtype %d", y_dtype); return 1; }
    if (((flags & 16) != 0) != ((flags & 32) != 0)) {            // measured on B200: a mixed pair faults (illegal instruction)
        set_error("conv3x3_igemm_bf16: tcgen05 kind::f16 needs both operands in the same format (set both or neither of DASV_CONV_W_F16, DASV_CONV_X_F16)"); return 1;
    }
    if ((flags & 128) && ((flags & 64) || !(flags & 1) || (Cin != 64 && Cin != 128))) {
        set_error("conv3x3_igemm_bf16: the fused first layer needs a ReLU forward pass in bf16/fp16 and Cin in {64, 128}"); return 1;
    }
    if ((flags & 64) && ((flags & 48) || !(flags & 1))) { set_error("conv3x3_igemm_bf16: the split (fp32x3) mode is bf16, forward only"); return 1; }
    if (!(flags & 1) && (flags & 48)) { set_error("conv3x3_igemm_bf16: the input-gradient pass is bf16 only"); return 1; }
    if (B <= 0 || T <= 0) return 0;
    if (!ws_query && ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(wp) & 15))) {
        set_error("conv3x3_igemm_bf16: x and wp must be 16-byte aligned"); return 1;
    }
    ConvKey k;
    memset(&k, 0, sizeof(k));
    k.x = x; k.wp = wp; k.B = B; k.T = T; k.F = F; k.Cin = Cin; k.Cout = Cout; k.flags = flags; k.y_dtype = y_dtype;
    if (cudaGetDevice(&k.dev) != cudaSuccess) { set_error("conv3x3_igemm_bf16: no current device"); cudaGetLastError(); return 1; }
    k.env_reuse = 1; k.env_pair = -1; k.env_sb = 0; k.ragged = lengths != nullptr;
    if (const char* e = getenv("DASV_CONV_NOSPLITK")) k.env_nosplit = atoi(e) != 0;
    if (const char* e = getenv("DASV_CONV_SPLITK")) k.env_split = atoi(e);
    if (const char* e = getenv("DASV_CONV_NOWIDE")) k.env_nowide = atoi(e) != 0;
    if (const char* e = getenv("DASV_CONV_REUSE")) k.env_reuse = atoi(e) != 0;
    if (const char* e = getenv("DASV_CONV_PAIR")) k.env_pair = atoi(e) != 0;
    if (const char* e = getenv("DASV_CONV_SB")) k.env_sb = atoi(e);
    if (const char* e = getenv("DASV_CONV_PLAN")) strncpy(k.env_plan, e, sizeof(k.env_plan) - 1);

    if (ws_query) {                                              // workspace size of this shape: the plan alone (dummy addresses for the maps)
        static __align__(128) unsigned char dummy[128];
        k.x = dummy; k.wp = dummy;
        ConvEntry en;
        if (conv_build_entry(en, k)) return 1;
        *ws_query = en.ws_bytes;
        return 0;
    }
    CUtensorMap tmA, tmB;
    ConvParams p;
    size_t smem, ws_bytes; int grid, pair;
    {
        std::lock_guard<std::mutex> lock(g_conv_mu);
        ConvEntry* hit = nullptr;
        for (int i = 0; i < g_conv_n; ++i)
            if (conv_key_eq(g_conv_cache[i].key, k)) { hit = &g_conv_cache[i]; break; }
        if (!hit) {
            ConvEntry en;
            if (conv_build_entry(en, k)) return 1;
            int slot = g_conv_n;
            if (g_conv_n < 64) ++g_conv_n;
            else {                                              // evict the least recently used entry
                slot = 0;
                for (int i = 1; i < 64; ++i) if (g_conv_cache[i].stamp < g_conv_cache[slot].stamp) slot = i;
            }
            g_conv_cache[slot] = en;
            hit = &g_conv_cache[slot];
        }
        hit->stamp = ++g_conv_clock;
        tmA = hit->tmA; tmB = hit->tmB; p = hit->p; smem = hit->smem; grid = hit->grid; pair = hit->pair; ws_bytes = hit->ws_bytes;
    }
    if (ws_bytes && !workspace) {
        set_error("conv3x3_igemm_bf16: this shape runs split-K and needs a workspace of %zu bytes (dasv_conv3x3_igemm_workspace_bytes)", ws_bytes);
        return 1;
    }
    p.ws = static_cast<float*>(workspace);
    p.bias = bias; p.lengths = lengths; p.y = y; p.mask = mask;
    p.trace = g_conv_trace;
    if (front) { p.x0 = front->x0; p.w11 = front->w11; p.b11 = front->b11; p.scratch = const_cast<void*>(x); }

    const int variant = (flags & 128) ? (act == 2 ? 9 : 8) : (pair ? 4 : 0) + (p.relu ? ((flags & 64) ? 3 : (act == 2 ? 2 : 0)) : 1);
    if (conv_raise_smem(variant, k.dev)) return 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kConvThreads + ((flags & 128) ? kConvFuseThreads : 0));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_kernel_variant(variant), tmA, tmB, p);
    if (e != cudaSuccess) { set_error("conv3x3_igemm_bf16: launch failed: %s", cudaGetErrorString(e)); return 1; }
    if (p.splitk > 1) {                                          // second half: sum the splits, bias, ReLU, pool, format
        const size_t total = static_cast<size_t>(B) * (p.pool ? (T + 1) / 2 : T) * (p.pool ? F / 2 : F) * (Cout / 4) * (p.pool ? 4 : 1);   // a lane per input pixel
        size_t blocks = (total + 255) / 256;
        const size_t cap = static_cast<size_t>(sm_count()) * 8;
        if (blocks > cap) blocks = cap;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int relu = p.relu, pool = p.pool, ref = p.ref_layout;
        if (p.y_f32) e = launch_pdl(conv_splitk_finish_kernel<0>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st,
                                    static_cast<const float*>(p.ws), p.splitk, bias, y, B, T, F, Cout, pool, ref, relu);
        else if (act == 2) e = launch_pdl(conv_splitk_finish_kernel<2>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st,
                                          static_cast<const float*>(p.ws), p.splitk, bias, y, B, T, F, Cout, pool, ref, relu);
        else e = launch_pdl(conv_splitk_finish_kernel<1>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st,
                            static_cast<const float*>(p.ws), p.splitk, bias, y, B, T, F, Cout, pool, ref, relu);
        if (e != cudaSuccess) { set_error("conv3x3_igemm_bf16: finishing launch failed: %s", cudaGetErrorString(e)); return 1; }
    }
    return check_launch("conv3x3_igemm_bf16");
}

// Debugging aid (scripts/b1_trace.py): every conv3x3_igemm launch that follows writes, per CTA, eight %globaltimer stamps (ns)
// into buf[blockIdx.x * 8 + i]: 0 CTA start, 1 set-up done, 2 previous kernel finished, 3 first operands landed, 4 last MMA
// issued, 5 first accumulator complete, 6 epilogue done, 7 CTA end.  buf: device memory for >= 8 * grid values; NULL switches
// the stamps off (the default; the kernels then only test the pointer).
extern "C" int dasv_debug_conv_trace(unsigned long long* buf) {
    g_conv_trace = buf;
    return 0;
}

extern "C" int dasv_conv3x3_igemm_bf16(const void* x, const void* wp, const float* bias, const int32_t* lengths,
                                       void* y, int y_dtype, int flags,
                                       int B, int T, int F, int Cin, int Cout, void* workspace, void* stream) {
    if (!bias) { set_error("conv3x3_igemm_bf16: null bias"); return 1; }
    return conv_igemm_launch(x, wp, bias, lengths, nullptr, y, y_dtype, flags, B, T, F, Cin, Cout, stream, nullptr, workspace);
}

// Bytes of `workspace` a call with these arguments needs: 0 except for shapes with so few tiles that the launch is split
// along K (small batches), where the partial sums of the splits pass through it.
extern "C" size_t dasv_conv3x3_igemm_workspace_bytes(int y_dtype, int flags, int has_lengths, int B, int T, int F, int Cin, int Cout) {
    size_t bytes = 0;
    static const int32_t some_lengths = 0;
    if (B <= 0 || T <= 0) return 0;
    cudaFree(nullptr);                                          // the plan needs the driver (tensor-map encoder): make sure the runtime is up
    if (conv_igemm_launch(nullptr, nullptr, nullptr, has_lengths ? &some_lengths : nullptr, nullptr, nullptr, y_dtype, flags,
                          B, T, F, Cin, Cout, nullptr, nullptr, nullptr, &bytes))
        return 0;
    return bytes;
}

extern "C" int dasv_conv3x3_dgrad_bf16(const void* g, const void* wp_rot, const void* relu_mask, const int32_t* lengths,
                                       void* dx, int B, int T, int F, int Cg, int Cx, void* stream) {
    return conv_igemm_launch(g, wp_rot, nullptr, lengths, relu_mask, dx, 1, 0, B, T, F, Cg, Cx, stream);
}

// conv11 + conv12 in one kernel (scripts/CNNs.py:72-74): relu(conv11(x0) + b11) never goes to HBM as a tensor.  Four extra warps
// of the implicit-GEMM kernel compute it per tile (haloed region, zeros outside the image and beyond the utterance) into a
// per-CTA, double-buffered scratch patch that stays in L2, one tile ahead of the MMAs; the TMA producer reads the patches
// from there.  Same values as dasv_conv11_direct + dasv_conv3x3_igemm_bf16 (bit-identical), without the 2 x 2.1 GB round trip
// of the 128-channel tensor.  `scratch`: dasv_conv12_fused_workspace_bytes(C1) bytes, contents irrelevant.
extern "C" size_t dasv_conv12_fused_workspace_bytes(int C1) {
    // [2 * SMs] patches of at most (BT+2)(BF+2) = BT*BF + 2(BT+BF) + 4 <= 256 + 2*257 + 4 = 774 pixels of C1 16-bit channels
    return static_cast<size_t>(2) * sm_count() * 776 * (C1 > 0 ? C1 : 0) * 2;
}

extern "C" int dasv_conv12_fused_bf16(const float* x0, const float* w11, const float* b11, const void* wp, const float* bias,
                                      const int32_t* lengths, void* y, void* scratch, int y_dtype, int flags,
                                      int B, int T, int F, int C1, int Cout, void* stream) {
    if (!x0 || !w11 || !b11 || !bias || !scratch) { set_error("conv12_fused: null argument"); return 1; }
    if (flags & (4 | 8 | 64 | 128)) { set_error("conv12_fused: flags may hold RELU, POOL and the operand formats only"); return 1; }
    if ((reinterpret_cast<uintptr_t>(scratch) & 127) != 0) { set_error("conv12_fused: scratch must be 128-byte aligned"); return 1; }
    const ConvFront front{x0, w11, b11};
    return conv_igemm_launch(scratch, wp, bias, lengths, nullptr, y, y_dtype, flags | 128, B, T, F, C1, Cout, stream, &front);
}
