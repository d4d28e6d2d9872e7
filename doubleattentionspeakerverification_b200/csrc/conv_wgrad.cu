// Weight gradient of the VGG front-end's 3x3 convolutions on the tensor cores (sm_100a): the training-side companion of
// conv_igemm.cu (the reference gets it from cuDNN through autograd, scripts/CNNs.py:59-66 + train.py:201).
//
//   dW[co][ci][ky][kx] = sum_{b,t,f} G[b,t,f,co] * X[b,t+ky-1,f+kx-1,ci]
//
// G = gradient at the conv output (after the ReLU / pool backward, zero for frames past an utterance), X = the layer's
// input, both bf16 NHWC.  This is a GEMM whose contraction index is the PIXEL, so with channels contiguous in memory both
// operands are MN-major: a 128-byte shared-memory row holds 64 channels of one pixel, rows are pixels, and tcgen05.mma
// reads them through MN-major SWIZZLE_128B descriptors (instruction-descriptor bits 15/16) exactly as TMA wrote them.
//
// One CTA owns a [128 co] x [N ci] tile of a GROUP of taps and a range of (utterance, frame tile) items (split-K).
// The taps' accumulators live side by side in TMEM (taps x N <= 512 columns): N = 128 with three groups of three taps
// (default), 64 with 5 + 4, or 32 with all nine.  Per item TMA loads
//   G tile  [BT frames][F+1 bins] x 128 co        (two 64-channel boxes; bin -1 is out of bounds = zero filled)
//   X patch [BT+2 frames][F+1 bins] x N ci        (one or two boxes, +-1 frame halo, the same zero-filled border column)
// One zero column per frame serves as the right border of frame t and the left border of frame t+1 (row pitch P = F+1).
// The contraction runs LINEARLY over the G tile's rows r = t*P + f + 1: the X row that pairs with G row r for tap
// (ky,kx) is r + ky*P + kx - 1, a constant row offset, so a tap is just a different start address of the same X
// patch.  The zero border columns of G make the wrap-around rows harmless, rows beyond the tile are zero because the
// whole shared memory is cleared once and TMA never writes them.
// Partial sums go to a workspace [split][tap][co][ci] (fp32) and a second kernel adds the splits in fixed order
// (deterministic) into the reference layout [Cout][Cin][3][3].
//
// Roofline: tensor.  The A tile is re-read from shared memory for every MMA, so operand bytes per FLOP fall with N:
// measured 470 TFLOP/s (N = 32), 606 (N = 64), 1040 (N = 128), 1159 (N = 128 on CTA pairs), 1248 (+ shared border column
// and tile-height selection), 1340 (+ no frame halo when a tap group is one kernel row, taller tiles) over the
// exampleModel's seven layers at batch 256 = 0.95 of the sustained cuBLAS bf16 rate; splitting the taps over CTAs costs
// extra tile loads (3x at N = 128) which stay below the SM's L2 ingest rate.
#include "common.cuh"
#include "tmap.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <stdio.h>

namespace dasv {

constexpr int kWgThreads = 256;      // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warp 3 idle, warps 4-7 epilogue
constexpr int kWgM = 128;            // output channels per tile (TMEM lanes)
constexpr uint32_t kWgTmemCols = 512;

struct WgradParams {
    float* ws;                       // [splits][9][Cout][Cin]
    float* ws_bias;                  // [splits][Cout] or NULL: column sums of G (the bias gradient), from one extra MMA per K step
    int B, T, F, Cin, Cout;
    int BT, n_tt, n_items, items_per_split, splits;
    int n_mt, n_nt;
    int N, tg, ng;                   // input channels per tile, taps per group, tap groups
    int halo;                        // frames of +-halo in the X patch: 1, or 0 when a tap group is one kernel row (tg == 3)
    int rowsG, K16, rowsX;
    uint32_t g_box_bytes;            // bytes one G box delivers (rowsG * 128)
    uint32_t g_alloc;                // K16 * 128
    uint32_t x_box_bytes, x_off;     // X patch inside a stage: x_off = 2 * g_alloc + 1024 (a guard row precedes the patch)
    uint32_t x_stride;               // N = 128: distance between the two 64-channel X boxes
    uint32_t stage_bytes;
    int stages;
};

// MN-major operand, 128-byte swizzle: a row = 64 channels (128 B) of one K index; 8 consecutive K rows form the 1024-byte
// swizzle atom (SBO between atoms along K); the next 64 channels live `lbo` bytes further (LBO).
DASV_DEVICE uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// PAIR: a 2-CTA cluster owns 256 output channels (tcgen05 cta_group::2): each CTA loads the G tile of its own 128 channels
// and HALF of the X patch (64 of the 128 input channels), so the operand bytes an SM feeds per MMA drop from 8 KB to 6 KB.
template <int TG, bool PAIR>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* ring = smem;
    unsigned char* ones = ring + static_cast<size_t>(p.stages) * p.stage_bytes;        // [16 rows][64 ch] of bf16 1.0
    uint64_t* full = reinterpret_cast<uint64_t*>(ones + 2048);
    uint64_t* empty = full + p.stages;
    uint64_t* acc_full = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;                 // 0 = the leader, which issues the MMAs
    const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int n_mu = PAIR ? p.n_mt / 2 : p.n_mt;                          // m units (tiles, or tile pairs)
    const int pairs = n_mu * p.n_nt;
    const int pair = unit % pairs;
    const int grp = (unit / pairs) % p.ng, split = unit / (pairs * p.ng);
    const int m = PAIR ? 2 * (pair % n_mu) + static_cast<int>(rank) : pair % n_mu, n = pair / n_mu;
    const int tap0 = grp * p.tg, tap1 = min(tap0 + p.tg, 9);
    const int N = p.N;
    const int xc0 = PAIR ? n * N + static_cast<int>(rank) * 64 : ((N == 32) ? (n >> 1) * 64 : n * N);   // first channel of this CTA's X box(es)
    const uint32_t xhalf = (!PAIR && N == 32) ? static_cast<uint32_t>(n & 1) * 64u : 0u;
    const int xboxes = (!PAIR && N == 128) ? 2 : 1;
    const bool do_bias = p.ws_bias != nullptr && n == 0 && grp == 0;        // this CTA also sums G over its items
    const uint32_t bias_col = static_cast<uint32_t>(p.tg * N);
    const int item0 = split * p.items_per_split;
    const int item1 = min(item0 + p.items_per_split, p.n_items);

    // clear the ring once: rows that TMA never writes (K padding, guard rows) must read as finite zeros
    {
        uint4* z = reinterpret_cast<uint4*>(ring);
        const size_t n16 = static_cast<size_t>(p.stages) * p.stage_bytes / 16;
        for (size_t i = threadIdx.x; i < n16; i += kWgThreads) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < 2048 / 16; i += kWgThreads)
            reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmX); }
    if (warp == 2) {
        if (PAIR) { tmem_alloc_2sm(tmem_slot, kWgTmemCols); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, kWgTmemCols); tmem_relinquish(); }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();           // the peer's barriers must exist before TMA completions / commits reach them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                            // TMA producer
            int st = 0;
            uint32_t ph = 0;
            for (int it = item0; it < item1; ++it) {
                const int b = it / p.n_tt, t0 = (it % p.n_tt) * p.BT;
                // with a halo the patch starts one frame early; without, the group IS kernel row ky = grp: frames t0 + ky - 1 ...
                const int tx = p.halo ? t0 - 1 : t0 + grp - 1;
                mbar_wait(&empty[st], ph ^ 1u);
                unsigned char* sb = ring + static_cast<size_t>(st) * p.stage_bytes;
                if (PAIR) {
                    // both CTAs load into their own shared memory; all bytes complete on the leader's barrier
                    if (rank == 0) mbar_arrive_expect_tx(&full[st], 2u * (2 * p.g_box_bytes + p.x_box_bytes));
                    tma_load_4d_2sm(sb, &tmG, &full[st], m * kWgM, -1, t0, b);
                    tma_load_4d_2sm(sb + p.g_alloc, &tmG, &full[st], m * kWgM + 64, -1, t0, b);
                    tma_load_4d_2sm(sb + p.x_off, &tmX, &full[st], xc0, -1, tx, b);
                } else {
                    mbar_arrive_expect_tx(&full[st], 2 * p.g_box_bytes + xboxes * p.x_box_bytes);
                    tma_load_4d(sb, &tmG, &full[st], m * kWgM, -1, t0, b);
                    tma_load_4d(sb + p.g_alloc, &tmG, &full[st], m * kWgM + 64, -1, t0, b);
                    tma_load_4d(sb + p.x_off, &tmX, &full[st], xc0, -1, tx, b);
                    if (xboxes == 2) tma_load_4d(sb + p.x_off + p.x_stride, &tmX, &full[st], xc0 + 64, -1, tx, b);
                }
                if (++st == p.stages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {               // MMA issuer (the pair's leader)
            const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * kWgM : kWgM, static_cast<uint32_t>(N)) | (1u << 15) | (1u << 16);   // A and B MN-major
            // one thread issues every MMA, so the loop must cost only a few instructions per MMA: descriptors are advanced
            // by adding row offsets (in 16-byte units) to the 64-bit value, taps are unrolled
            const int frow = p.F + 1;               // row pitch: F bins + the shared zero border column
            long long toff[TG];
#pragma unroll
            for (int j = 0; j < TG; ++j) {
                const int tap = min(tap0 + j, 8);
                const int ky = p.halo ? tap / 3 : 0;         // without halo the patch already starts at the group's kernel row
                toff[j] = static_cast<long long>(ky * frow + (tap % 3) - 1) * 8;                  // rows * 128 B >> 4; -1 lands on the guard row
            }
            const int ntaps = tap1 - tap0;
            const uint32_t idesc_ones = umma_idesc_bf16(PAIR ? 2 * kWgM : kWgM, PAIR ? 32 : 16) | (1u << 15) | (1u << 16);
            const uint64_t ones_desc = umma_desc_mn128(smem_u32(ones), 16);
            int st = 0;
            uint32_t ph = 0;
            uint32_t acc = 0;
            for (int it = item0; it < item1; ++it) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(ring + static_cast<size_t>(st) * p.stage_bytes);
                uint64_t a_desc = umma_desc_mn128(sb, p.g_alloc);
                uint64_t b_desc = umma_desc_mn128(sb + p.x_off + xhalf, (!PAIR && N == 128) ? p.x_stride : 16u);
                for (int k = 0; k < p.K16; k += 16) {
#pragma unroll
                    for (int j = 0; j < TG; ++j)
                        if (j < ntaps) {
                            if (PAIR) umma_bf16_2sm(tmem_base + static_cast<uint32_t>(j * N), a_desc, b_desc + static_cast<uint64_t>(toff[j]), idesc, acc);
                            else umma_bf16(tmem_base + static_cast<uint32_t>(j * N), a_desc, b_desc + static_cast<uint64_t>(toff[j]), idesc, acc);
                        }
                    acc = 1u;
                    a_desc += 16 * 8;                   // 16 rows of 128 B
                    b_desc += 16 * 8;
                }
                if (do_bias) {                      // column sums of G: G^T . 1 against a tile of ones
                    uint64_t g_desc = umma_desc_mn128(sb, p.g_alloc);
#pragma unroll 1
                    for (int k = 0; k < p.K16; k += 16) {
                        if (PAIR) umma_bf16_2sm(tmem_base + bias_col, g_desc, ones_desc, idesc_ones, (it > item0 || k > 0) ? 1u : 0u);
                        else umma_bf16(tmem_base + bias_col, g_desc, ones_desc, idesc_ones, (it > item0 || k > 0) ? 1u : 0u);
                        g_desc += 16 * 8;
                    }
                }
                if (PAIR) umma_commit_2sm(&empty[st], 3); else umma_commit(&empty[st]);   // frees the stage (in both CTAs) when the MMAs above have read it
                if (++st == p.stages) { st = 0; ph ^= 1u; }
            }
            if (PAIR) umma_commit_2sm(acc_full, 3); else umma_commit(acc_full);
        }
    } else if (warp >= 4) {
        // epilogue: lane = output channel, 32 input channels per tap
        const int q = warp & 3;
        const int co = m * kWgM + q * 32 + lane;
        if (item1 > item0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int tap = tap0; tap < tap1; ++tap) {
#pragma unroll 1
            for (int c32 = 0; c32 < N; c32 += 32) {
                uint32_t r[32];
                if (item1 > item0) {
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>((tap - tap0) * N + c32), r);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                if (co < p.Cout) {
                    float4* dst = reinterpret_cast<float4*>(p.ws + ((static_cast<size_t>(split) * 9 + tap) * p.Cout + co) * p.Cin + n * N + c32);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                             __uint_as_float(r[4 * j + 3]));
                }
            }
        }
        if (do_bias) {
            uint32_t r[32];
            float v = 0.f;
            if (item1 > item0) {
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + bias_col, r);
                tc_wait_ld();
                v = __uint_as_float(r[0]);
            }
            if (co < p.Cout) p.ws_bias[static_cast<size_t>(split) * p.Cout + co] = v;
        }
        tc_fence_before();
    }
    if (PAIR) cluster_sync_all();           // nobody leaves (or frees TMEM) while the peer may still signal into this CTA
    else __syncthreads();
    if (warp == 2) { tc_fence_after(); if (PAIR) tmem_dealloc_2sm(tmem_base, kWgTmemCols); else tmem_dealloc(tmem_base, kWgTmemCols); }
}

// dW[co][ci][tap] (+)= sum over splits, fixed order
__global__ void conv_wgrad_reduce_kernel(const float* ws, float* dw, const float* ws_bias, float* db, int splits, int Cout, int Cin, int accumulate) {
    const size_t n = static_cast<size_t>(Cout) * Cin;
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // (co, ci)
    if (db != nullptr && i < static_cast<size_t>(Cout)) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += ws_bias[static_cast<size_t>(sp) * Cout + i];
        db[i] = accumulate ? db[i] + s : s;
    }
    if (i >= n) return;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += ws[(static_cast<size_t>(sp) * 9 + tap) * n + i];
        float* d = dw + i * 9 + tap;
        *d = accumulate ? *d + s : s;
    }
}

struct WgradPlan {
    int ok, BT, n_tt, n_items, splits, items_per_split, n_mt, n_nt, N, tg, ng, rowsG, K16, rowsX, stages, pair, halo;
    uint32_t g_alloc, x_off, x_stride, stage_bytes;
    size_t smem;
};

static WgradPlan wgrad_plan(int B, int T, int F, int Cin, int Cout, int sms) {
    WgradPlan pl{};
    if (Cin % 64 != 0 || Cout % 8 != 0 || F < 1 || F + 1 > 256 || T < 1 || B < 1) return pl;
    // input channels per tile: the operand feed from shared memory (MN-major reads measured at ~64 B/clk) bounds the MMA
    // rate, and bytes per FLOP fall with N, but 9 taps x N accumulator columns must fit the 512 TMEM columns, so the taps
    // are split into groups handled by different CTAs (each loads the tiles again): N = 128 / 3 groups, 64 / 2, 32 / 1
    pl.N = (Cin % 128 == 0) ? 128 : 64;
    if (const char* e = getenv("DASV_WGRAD_N")) { const int v = atoi(e); if ((v == 32 || v == 64 || v == 128) && Cin % v == 0) pl.N = v; }
    pl.tg = pl.N == 128 ? 3 : (pl.N == 64 ? 5 : 9);
    pl.ng = pl.N == 128 ? 3 : (pl.N == 64 ? 2 : 1);
    // three groups of three taps = one kernel row each: the X patch of a group needs no frame halo (only its own row of
    // frames, shifted by ky - 1), which cuts its bytes by (BT + 2) / BT and lets taller tiles fit
    pl.halo = pl.tg == 3 ? 0 : 1;
    if (getenv("DASV_WGRAD_HALO")) pl.halo = 1;
    // CTA pairs (DASV_WGRAD_PAIR=0 disables): 256 output channels per cluster, each CTA holds half of the X patch;
    // measured 1141 vs 1040 TFLOP/s over the exampleModel's seven layers (conv22: 1426 vs 1260)
    pl.pair = (pl.N == 128 && ((Cout + kWgM - 1) / kWgM) % 2 == 0) ? 1 : 0;
    if (const char* e = getenv("DASV_WGRAD_PAIR")) { if (atoi(e) == 0) pl.pair = 0; }
    const int xboxes = (pl.N == 128 && !pl.pair) ? 2 : 1;
    const uint32_t avail = 227u * 1024u - 2048u - 1024u - 256u;
    // frames per tile: at most `row_cap` contraction rows per stage, and among those the height that wastes the fewest MMA rows
    // on the 16-row padding of a tile and on the partly empty last tile of an utterance
    {
        int row_cap = 352;                           // measured: 176 -> 1280, 256 -> 1323, 352+ -> 1330-1350 TFLOP/s (two stages must still fit)
        if (const char* e = getenv("DASV_WGRAD_ROWS")) { const int v = atoi(e); if (v >= 16 && v <= 1024) row_cap = v; }
        const int bt_max = max(1, min(min(row_cap / (F + 1), T), 254));
        double best = -1.0;
        pl.BT = bt_max;
        for (int bt = bt_max; bt >= max(1, bt_max / 2); --bt) {
            const int k16 = (bt * (F + 1) + 15) / 16 * 16;
            const int ntt = (T + bt - 1) / bt;
            const double useful = static_cast<double>(T) * F / (static_cast<double>(ntt) * k16);
            if (useful > best + 1e-9) { best = useful; pl.BT = bt; }
        }
    }
    for (;; --pl.BT) {                               // largest frame tile that leaves room for two stages
        pl.rowsG = pl.BT * (F + 1);
        pl.K16 = (pl.rowsG + 15) / 16 * 16;
        pl.rowsX = (pl.BT + 2 * pl.halo) * (F + 1);
        pl.g_alloc = (static_cast<uint32_t>(pl.K16) * 128u + 1023u) & ~1023u;
        pl.x_off = 2 * pl.g_alloc + 1024u;
        // X rows an MMA view may touch: -1 .. K16 - 1 + 2 halo (F + 1) + 1
        pl.x_stride = ((static_cast<uint32_t>(pl.K16 + 2 * pl.halo * (F + 1) + 2)) * 128u + 1023u) & ~1023u;
        pl.stage_bytes = pl.x_off + xboxes * pl.x_stride;
        if (2 * pl.stage_bytes <= avail || pl.BT == 1) break;
    }
    pl.stages = static_cast<int>(avail / pl.stage_bytes);
    if (pl.stages > 4) pl.stages = 4;
    if (pl.stages < 1) return pl;
    pl.smem = static_cast<size_t>(pl.stages) * pl.stage_bytes + 2048 + 1024 + 256;
    pl.n_tt = (T + pl.BT - 1) / pl.BT;
    pl.n_items = B * pl.n_tt;
    pl.n_mt = (Cout + kWgM - 1) / kWgM;
    pl.n_nt = Cin / pl.N;
    const int pairs = pl.n_mt * pl.n_nt * pl.ng;       // CTAs per split (a CTA pair counts as two)
    // split-K factor: fill the SMs in whole waves, keep >= 2 items per CTA, bound the workspace
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64; ++s) {
        if (s > 1 && (pl.n_items + s - 1) / s < 2) break;
        const double ctas = static_cast<double>(pairs) * s;
        const double waves = (ctas + sms - 1) / sms;
        const double eff = ctas / (static_cast<double>(static_cast<long long>(waves)) * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    if (const char* e = getenv("DASV_WGRAD_SPLITS")) { const int v = atoi(e); if (v >= 1 && v <= 256) best = v; }
    pl.splits = best;
    pl.items_per_split = (pl.n_items + best - 1) / best;
    pl.ok = 1;
    return pl;
}

}  // namespace dasv

using namespace dasv;

extern "C" size_t dasv_conv3x3_wgrad_workspace_bytes(int B, int T, int F, int Cin, int Cout) {
    const int sms = sm_count();
    const WgradPlan pl = wgrad_plan(B, T, F, Cin, Cout, sms);
    if (!pl.ok) return 0;
    return static_cast<size_t>(pl.splits) * (static_cast<size_t>(9) * Cout * Cin + Cout) * sizeof(float);
}

extern "C" int dasv_conv3x3_wgrad_bf16(const void* x, const void* g, float* dw, float* db, void* workspace, int accumulate,
                                       int B, int T, int F, int Cin, int Cout, void* stream) {
    if (B < 0 || T < 0) { set_error("conv3x3_wgrad_bf16: negative shape"); return 1; }
    if (!x || !g || !dw || !workspace) { set_error("conv3x3_wgrad_bf16: null pointer"); return 1; }
    if (Cout % 64 != 0) { set_error("conv3x3_wgrad_bf16: Cout %d must be a multiple of 64", Cout); return 1; }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (B == 0 || T == 0) {
        if (!accumulate) {
            cudaMemsetAsync(dw, 0, static_cast<size_t>(Cout) * Cin * 9 * sizeof(float), s);
            if (db) cudaMemsetAsync(db, 0, static_cast<size_t>(Cout) * sizeof(float), s);
        }
        return check_launch("conv3x3_wgrad_bf16");
    }
    const WgradPlan pl = wgrad_plan(B, T, F, Cin, Cout, sms);
    if (!pl.ok) { set_error("conv3x3_wgrad_bf16: unsupported shape T=%d F=%d Cin=%d Cout=%d (need Cin %% 64 == 0, F <= 254)", T, F, Cin, Cout); return 1; }
    EncodeTiledFn encode = get_encode_tiled();
    if (!encode) { set_error("conv3x3_wgrad_bf16: cuTensorMapEncodeTiled is not available from the CUDA driver"); return 1; }
    CUtensorMap tmG, tmX;
    for (int which = 0; which < 2; ++which) {
        const int C = which ? Cin : Cout;
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(F), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(F) * C * 2, static_cast<cuuint64_t>(T) * F * C * 2};
        const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(F + 1), static_cast<cuuint32_t>(which ? pl.BT + 2 * pl.halo : pl.BT), 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(which ? &tmX : &tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(which ? x : g), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3_wgrad_bf16: tensor map encode failed (%d)", static_cast<int>(r)); return 1; }
    }
    WgradParams p{};
    p.ws = static_cast<float*>(workspace);
    p.ws_bias = db ? p.ws + static_cast<size_t>(pl.splits) * 9 * Cout * Cin : nullptr;
    p.B = B; p.T = T; p.F = F; p.Cin = Cin; p.Cout = Cout;
    p.BT = pl.BT; p.n_tt = pl.n_tt; p.n_items = pl.n_items; p.items_per_split = pl.items_per_split; p.splits = pl.splits;
    p.n_mt = pl.n_mt; p.n_nt = pl.n_nt; p.N = pl.N; p.tg = pl.tg; p.ng = pl.ng; p.halo = pl.halo; p.rowsG = pl.rowsG; p.K16 = pl.K16; p.rowsX = pl.rowsX;
    p.g_box_bytes = static_cast<uint32_t>(pl.rowsG) * 128u; p.g_alloc = pl.g_alloc;
    p.x_box_bytes = static_cast<uint32_t>(pl.rowsX) * 128u; p.x_off = pl.x_off; p.x_stride = pl.x_stride;
    p.stage_bytes = pl.stage_bytes; p.stages = pl.stages;
    if (getenv("DASV_CONV_DEBUG"))
        fprintf(stderr, "wgrad plan: B=%d T=%d F=%d Cin=%d Cout=%d BT=%d rowsG=%d K16=%d rowsX=%d stages=%d stage_bytes=%u items=%d splits=%d N=%d pair=%d\n",
                B, T, F, Cin, Cout, pl.BT, pl.rowsG, pl.K16, pl.rowsX, pl.stages, pl.stage_bytes, pl.n_items, pl.splits, pl.N, pl.pair);
    auto kern = pl.pair ? conv_wgrad_kernel<3, true>
                        : (pl.tg == 3 ? conv_wgrad_kernel<3, false> : (pl.tg == 5 ? conv_wgrad_kernel<5, false> : conv_wgrad_kernel<9, false>));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
    if (e != cudaSuccess) { set_error("conv3x3_wgrad_bf16: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e)); return 1; }
    const int grid = pl.n_mt * pl.n_nt * pl.ng * pl.splits;
    if (pl.pair) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(static_cast<unsigned>(grid));       // n_mt is even: consecutive CTAs form the pairs
        cfg.blockDim = dim3(kWgThreads);
        cfg.dynamicSmemBytes = pl.smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tmG, tmX, p);
        if (e != cudaSuccess) { set_error("conv3x3_wgrad_bf16: pair launch failed: %s", cudaGetErrorString(e)); return 1; }
    } else {
        kern<<<grid, kWgThreads, pl.smem, s>>>(tmG, tmX, p);
    }
    if (check_launch("conv3x3_wgrad_bf16")) return 1;
    const size_t n = static_cast<size_t>(Cout) * Cin;
    conv_wgrad_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p.ws, dw, p.ws_bias, db, pl.splits, Cout, Cin, accumulate);
    return check_launch("conv3x3_wgrad_reduce");
}
