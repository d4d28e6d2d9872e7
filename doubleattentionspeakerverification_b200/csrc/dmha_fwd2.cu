// Fused DoubleMHA pooling forward, v2 mapping (the default path of dasv_dmha_fwd).
// Replaces scripts/poolings.py:73-80, :100-109, :45-51, :61-71, :126-129 with ONE pass over x.
//
// Roofline: HBM.  Algorithmic bytes = sum_b L_b * D * sizeof(x) (+ outputs).  What ncu showed on the way here:
//   * one 16-byte vector per lane (v1, dmha_fwd.cu) is instruction-issue bound (0.74 warp instructions per input
//     float): the per-row shuffle reduction and the redundant softmax bookkeeping dominate.  Here a lane owns NV
//     vectors (16-24 elements) of a (frame, head) row, G = 2..32 lanes per row, so a row costs log2(G) shuffles; the
//     running sums are rescaled lazily (only when the running max grows by > 2^kLazy); the multiply-adds are packed
//     fma.rn.f32x2 (sm_100), two elements per instruction.
//   * Row groups of one LDS.128 phase start at different vectors (rot) so that rows whose byte size is a multiple
//     of 128 do not collide on the same banks.
//   * Utterances are handed out DYNAMICALLY: the producer thread of a CTA takes the next utterance from an atomic
//     counter (in the caller's workspace) whenever it is ready to stream a new one and tells the consumers through a
//     per-stage descriptor (utterance, first frame, frame count, last-stage flag) written before the stage's mbarrier
//     is armed.  With ragged lengths this removes the makespan penalty of a static round-robin deal (a CTA that drew
//     two long utterances no longer decides the kernel time).  Without a workspace the deal is static.
//   * Tried and dropped (round 2): a tenth warp that runs the whole end-of-utterance stage while the consumers move on
//     (partials handed over through a pair of mbarriers, lanes mapped (head, part of dh) so that no shuffle sits in the
//     merge loop): bit-identical, but slower -- bf16 full 43.3 -> 46.7 us, ragged sorted 35.1 -> 41.3 us: one warp takes
//     2-3 us for what eight warps do in well under 1 us, and that latency lands on the end of every CTA's last utterance,
//     which is where this kernel loses its time (B=2960 utterances: 0.92 of the HBM peak for bf16, 1.07 of the copy figure
//     for fp32; B=512: a fixed ~10 us of launch, ramp and tail on top of 33 us of streaming).
//   * Tried and dropped (round 2): one more ring stage per CTA (80 instead of 64 KB in flight) paid for with a single copy of
//     the merge scratch and a trailing barrier per utterance: bf16 full 43.1 -> 43.6 us, ragged 38.4 -> 38.9 us.
//   * Tried and dropped (round 2): ONE 16-warp CTA per SM with a 9-stage ring (all of an SM's bytes in flight for one utterance):
//     bf16 full 43.1 -> 59.7 us, fp32 68.7 -> 73.9 us -- the end-of-utterance stage idles the whole SM without a second CTA.
//   * Tried and dropped: cutting the flattened (utterance, frame) stream into equal per-CTA ranges with partial
//     states + tickets in the workspace (bit-exact, but the per-segment finish -- partial write, fence, ticket, merge,
//     ~3-4 us -- cost more than the 13 % tail it removed: fp32 83.8 us vs 76.1 us at B=512,T=200,D=1024,H=16).
#include "dmha_common.cuh"
#include "dmha_finish.cuh"
#include <math.h>
#include <stdlib.h>

namespace dasv {

constexpr float kDmhaLazy = 8.0f;
constexpr int kDmha2MaxGrid = 1024;

DASV_DEVICE uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
DASV_DEVICE uint64_t pack_u32x2(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}

// one 16-byte vector of a row as packed fp32 pairs: 2 pairs (fp32 input) or 4 pairs (bf16 input)
template <int VE, bool BF16>
DASV_DEVICE void load_row_pairs(const unsigned char* p, uint64_t (&x2)[VE / 2]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    if constexpr (BF16) {
        x2[0] = pack_u32x2(v.x << 16, v.x & 0xFFFF0000u);
        x2[1] = pack_u32x2(v.y << 16, v.y & 0xFFFF0000u);
        x2[2] = pack_u32x2(v.z << 16, v.z & 0xFFFF0000u);
        x2[3] = pack_u32x2(v.w << 16, v.w & 0xFFFF0000u);
    } else {
        x2[0] = pack_u32x2(v.x, v.y);
        x2[1] = pack_u32x2(v.z, v.w);
    }
}

struct Dmha2Smem {
    uint32_t ring, q, a, pacc, pm, pl, u, w, meta, bars, total;
};
__host__ __device__ inline Dmha2Smem dmha2_smem(int D, int H, int dh, int S, int stages, uint32_t stage_bytes) {
    Dmha2Smem s;
    uint32_t o = 0;
    s.ring = o; o += stages * stage_bytes;
    s.q = o;    o += D * 4;
    s.a = o;    o += dh * 4;
    s.pacc = o; o += 2 * S * D * 4;                  // scratch is double-buffered by utterance parity
    s.pm = o;   o += 2 * H * S * 4;
    s.pl = o;   o += 2 * H * S * 4;
    s.u = o;    o += 2 * H * 4;
    s.w = o;    o += 2 * H * 4;
    o = (o + 15u) & ~15u;
    s.meta = o; o += stages * 16;
    s.bars = o; o += 2 * stages * 8;
    s.total = o;
    return s;
}

// What a ring stage holds: frames [t0, t0 + nf) of utterance b, whose (clamped) length is Lb.  b < 0 ends the CTA's
// work.  Written by the producer before it arms the stage's barrier, read by the consumers after their wait.
struct __align__(16) Dmha2Stage {
    int b, t0, nf, Lb;
};

template <bool BF16, int NV>
constexpr int dmha_fwd2_min_ctas() { return (NV * (BF16 ? 8 : 4) <= 12) ? 3 : 2; }

template <bool BF16, int G, int NV, int FB, bool RAGGED>
__global__ void __launch_bounds__(kDmhaThreads, dmha_fwd2_min_ctas<BF16, NV>()) dmha_fwd2_kernel(const DmhaFwdParams p) {
    constexpr int VE = BF16 ? 8 : 4;
    constexpr int VP = VE / 2;                                  // packed pairs per vector
    constexpr uint32_t ES = BF16 ? 2u : 4u;
    constexpr int RPW = 32 / G;                                 // (frame, head) rows per warp
    constexpr int GPP = (G >= 8) ? 1 : 8 / G;                   // row groups per 8-lane LDS.128 phase
    extern __shared__ __align__(128) unsigned char smem[];

    const int D = p.D, H = p.H, dh = p.dh, S = p.S, T = p.T;
    const uint32_t frame_bytes = static_cast<uint32_t>(D) * ES;
    const uint32_t stage_bytes = p.fps * frame_bytes;
    const Dmha2Smem L = dmha2_smem(D, H, dh, S, p.stages, stage_bytes);
    unsigned char* ring = smem + L.ring;
    float* q_sm = reinterpret_cast<float*>(smem + L.q);         // [H][dh]  (query transposed)
    float* a_sm = reinterpret_cast<float*>(smem + L.a);         // [dh]
    float* pacc = reinterpret_cast<float*>(smem + L.pacc);      // [H*S][dh] partial weighted sums -> ctx
    float* pm = reinterpret_cast<float*>(smem + L.pm);          // [H*S] reference max (log2 units)
    float* pl = reinterpret_cast<float*>(smem + L.pl);          // [H*S] running sum
    float* u_sm = reinterpret_cast<float*>(smem + L.u);         // [H] head scores
    float* w_sm = reinterpret_cast<float*>(smem + L.w);         // [H] head weights
    Dmha2Stage* meta = reinterpret_cast<Dmha2Stage*>(smem + L.meta);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kDmhaConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) griddep_launch();
    // Programmatic dependent launch: x, the deal counter and the memory behind the outputs belong to earlier kernels of the
    // stream.  The producer waits for them before its first claim; the consumers first stage the query and att (module
    // parameters, never written by a kernel that lets its dependents start early) and wait afterwards.

    if (warp == kDmhaConsumerWarps) {
        if (lane == 0) {                            // producer: HBM -> SMEM ring, one linear bulk copy per stage
            // the first utterance of a CTA is its own index (no round trip to the counter, and its length -- an input of the
            // call, never written by a kernel that lets its dependents start early -- is fetched while the previous kernel
            // still runs); the counter hands out the utterances from gridDim.x on
            const int hs = p.hs > 1 ? p.hs : 1;     // head split: pseudo-utterance b = head group b % hs of utterance b / hs
            const size_t row_bytes = hs > 1 ? static_cast<size_t>(p.ldx) * ES : frame_bytes;
            int b = static_cast<int>(blockIdx.x);
            int Lnext = (p.lengths && b < p.B) ? p.lengths[b / hs] : T;
            griddep_wait();
            int st = 0;
            uint32_t ph = 0;
            while (b < p.B) {
                int Lb = max(0, min(Lnext, T));
                const unsigned char* xb = p.x + static_cast<size_t>(b / hs) * T * row_bytes + static_cast<size_t>(b % hs) * frame_bytes;
                int f0 = 0;
                do {                                // an empty utterance still gets one (empty) stage so that it is finished
                    const int nf = min(p.fps, Lb - f0);
                    mbar_wait(&empty[st], ph ^ 1u);
                    meta[st] = Dmha2Stage{b, f0, nf, Lb};
                    if (nf > 0) {
                        const uint32_t bytes = static_cast<uint32_t>(nf) * frame_bytes;
                        mbar_arrive_expect_tx(&full[st], bytes);
                        if (hs == 1) {
                            bulk_g2s(ring + st * stage_bytes, xb + static_cast<size_t>(f0) * frame_bytes, bytes, &full[st]);
                        } else {                        // this head group's slice of every frame: one copy per row
                            for (int r = 0; r < nf; ++r)
                                bulk_g2s(ring + st * stage_bytes + static_cast<uint32_t>(r) * frame_bytes,
                                         xb + static_cast<size_t>(f0 + r) * row_bytes, frame_bytes, &full[st]);
                        }
                    } else {
                        mbar_arrive(&full[st]);
                    }
                    if (++st == p.stages) { st = 0; ph ^= 1u; }
                    f0 += p.fps;
                } while (f0 < Lb);
                b = p.ws_cnt ? static_cast<int>(gridDim.x) + atomicAdd(p.ws_cnt, 1) : b + static_cast<int>(gridDim.x);
                if (b < p.B) Lnext = p.lengths ? p.lengths[b / hs] : T;
            }
            mbar_wait(&empty[st], ph ^ 1u);         // terminator stage
            meta[st] = Dmha2Stage{-1, 0, 0, 0};
            mbar_arrive(&full[st]);
            dmha_release_counter(p.ws_cnt);
        }
        return;
    }

    // ---------------------------------------------------------------- consumers (the producer is already streaming)
    for (int i = tid; i < D; i += kDmhaConsumerThreads) {
        const int h = i / dh, d = i - h * dh;
        q_sm[i] = p.hs > 1 ? p.query[d * p.Hq + static_cast<int>(blockIdx.x % p.hs) * H + h]   // this CTA's head group (static deal)
                           : p.query[d * H + h];    // reference layout [dh, H] (poolings.py:90)
    }
    if (p.att != nullptr)
        for (int i = tid; i < dh; i += kDmhaConsumerThreads) a_sm[i] = p.att[i];
    named_bar_sync(1, kDmhaConsumerThreads);
    griddep_wait();

    const int grp = warp * RPW + lane / G, lig = lane % G;
    const int head = grp % H, slot = grp / H;       // this group's head and frame slot (frames f = slot mod S)
    const bool active = slot < S;
    const int rot = (lane / G) % GPP;
    uint32_t voff[NV];
    bool vok[NV];
    uint64_t q2[NV][VP];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int idx = ((v + rot) % NV) * G + lig;             // 16-byte vector of the row held in slot v
        vok[v] = active && idx * VE < dh;
        voff[v] = static_cast<uint32_t>(head) * dh * ES + static_cast<uint32_t>(idx) * 16u;
#pragma unroll
        for (int e = 0; e < VP; ++e)
            q2[v][e] = vok[v] ? pack_f32x2(q_sm[head * dh + idx * VE + 2 * e], q_sm[head * dh + idx * VE + 2 * e + 1])
                              : pack_f32x2(0.f, 0.f);
    }

    int st = 0, par = 0;
    uint32_t ph = 0;
    float m = -INFINITY, l = 0.f;
    uint64_t acc2[NV][VP];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int e = 0; e < VP; ++e) acc2[v][e] = pack_f32x2(0.f, 0.f);

    while (true) {
        mbar_wait(&full[st], ph);
        const Dmha2Stage sg = meta[st];             // read before the stage is released
        if (sg.b < 0) break;
        const int b = sg.b, nf = sg.nf;
        const unsigned char* sbase = ring + st * stage_bytes;
        // FB rows per trip: rows fb+slot and fb+S+slot are independent until the softmax update, which gives
        // the LDS -> FMA -> shuffle -> exp2 chain a second row to overlap with.
        for (int fb = 0; fb < nf; fb += FB * S) {               // warp-uniform trip count (fps is a multiple of S)
            uint64_t xs[FB][NV][VP];
            float sc[FB];
            bool valid[FB];
#pragma unroll
            for (int r = 0; r < FB; ++r) {
                const int f = fb + r * S + slot;
                valid[r] = active && f < nf;
                const unsigned char* row = sbase + static_cast<uint32_t>(valid[r] ? f : 0) * frame_bytes;   // safe address when idle
                uint64_t s2a = pack_f32x2(0.f, 0.f), s2b = s2a;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if (RAGGED && !vok[v]) {
#pragma unroll
                        for (int e = 0; e < VP; ++e) xs[r][v][e] = pack_f32x2(0.f, 0.f);
                    } else {
                        load_row_pairs<VE, BF16>(row + voff[v], xs[r][v]);
                    }
#pragma unroll
                    for (int e = 0; e < VP; e += 2) {
                        s2a = fma_f32x2(xs[r][v][e], q2[v][e], s2a);
                        s2b = fma_f32x2(xs[r][v][e + 1], q2[v][e + 1], s2b);
                    }
                }
                float a0, a1, b0, b1;
                unpack_f32x2(s2a, a0, a1);
                unpack_f32x2(s2b, b0, b1);
                sc[r] = (a0 + a1) + (b0 + b1);
            }
#pragma unroll
            for (int r = 0; r < FB; ++r) sc[r] = group_sum<G>(sc[r]);
            float mx = -INFINITY;
#pragma unroll
            for (int r = 0; r < FB; ++r) {
                sc[r] = valid[r] ? sc[r] * p.scale_log2 : -INFINITY;     // log2-unit score of this (frame, head)
                mx = fmaxf(mx, sc[r]);
                if (p.align != nullptr && valid[r] && lig == 0)
                    p.align[(static_cast<size_t>(b) * T + sg.t0 + fb + r * S + slot) * H + head] = sc[r];   // raw score, normalised at the end
            }
            if (mx > m + kDmhaLazy) {                            // lazy rescale; first frame: m = -inf -> corr = 0
                const float corr = fast_exp2(m - mx);
                const uint64_t corr2 = pack_f32x2(corr, corr);
                l *= corr;
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int e = 0; e < VP; ++e) acc2[v][e] = mul_f32x2(acc2[v][e], corr2);
                m = mx;
            }
            const float mref = (m == -INFINITY) ? 0.f : m;      // idle group: exp2(-inf - 0) = 0
#pragma unroll
            for (int r = 0; r < FB; ++r) {
                const float pr = fast_exp2(sc[r] - mref);
                const uint64_t pr2 = pack_f32x2(pr, pr);
                l += pr;
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int e = 0; e < VP; ++e) acc2[v][e] = fma_f32x2(pr2, xs[r][v][e], acc2[v][e]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
        if (sg.t0 + sg.nf < sg.Lb) continue;

        // ============================================================ end of utterance b
        float* pacc_b = pacc + par * (S * D);
        float* pm_b = pm + par * (S * H);
        float* pl_b = pl + par * (S * H);
        // ------------------------------------------------------------ merge the S frame slots per head
        if (active) {
            const int sl = head * S + slot;
            if (lig == 0) { pm_b[sl] = m; pl_b[sl] = l; }
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (vok[v]) {
                    const int idx = ((v + rot) % NV) * G + lig;
#pragma unroll
                    for (int e = 0; e < VP; ++e) {
                        float lo, hi;
                        unpack_f32x2(acc2[v][e], lo, hi);
                        pacc_b[sl * dh + idx * VE + 2 * e] = lo;
                        pacc_b[sl * dh + idx * VE + 2 * e + 1] = hi;
                    }
                }
        }
        m = -INFINITY; l = 0.f;                                  // state for the next utterance
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < VP; ++e) acc2[v][e] = pack_f32x2(0.f, 0.f);
        dmha_finish_utterance2<kDmhaConsumerWarps>(p, b, sg.Lb, S, pacc_b, pm_b, pl_b, u_sm + par * H, w_sm + par * H, a_sm, tid, warp, lane);
        par ^= 1;
    }
}

// The attention over heads (poolings.py:45-51, :61-71) as its own small kernel, for launches that split the heads of an
// utterance over several CTAs: u[h] = <ctx[b,h,:], att> (-inf where keep == 0), headw = softmax_h(u), out = sum_h headw ctx.
// One CTA per utterance; the operation order is dmha_finish_utterance2's.
__global__ void __launch_bounds__(256) dmha_heads_kernel(const float* __restrict__ ctx, const float* __restrict__ att,
                                                         const uint8_t* __restrict__ keep, float* __restrict__ out,
                                                         float* __restrict__ headw, int H, int dh) {
    extern __shared__ float hs_sm[];              // u[H], w[H]
    float* u_sm = hs_sm;
    float* w_sm = hs_sm + H;
    griddep_launch();
    griddep_wait();
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* cb = ctx + static_cast<size_t>(b) * H * dh;
    for (int h = warp; h < H; h += 8) {
        float dot = 0.f;
        for (int d = lane; d < dh; d += 32) dot = fmaf(cb[h * dh + d], att[d], dot);
        dot = warp_sum(dot);
        if (lane == 0) {
            const bool kept = keep == nullptr || keep[static_cast<size_t>(b) * H + h] != 0;       // poolings.py:42
            u_sm[h] = kept ? dot : -INFINITY;
        }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int h = lane; h < H; h += 32) mx = fmaxf(mx, u_sm[h]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int h = lane; h < H; h += 32) sum += expf(u_sm[h] - mx);          // all heads dropped -> NaN, as in the reference
    sum = warp_sum(sum);
    if (warp == 0)
        for (int h = lane; h < H; h += 32) {
            const float w = expf(u_sm[h] - mx) / sum;
            w_sm[h] = w;
            if (headw != nullptr) headw[static_cast<size_t>(b) * H + h] = w;
        }
    __syncthreads();
    if (out != nullptr)
        for (int d = tid; d < dh; d += 256) {
            float o = 0.f;
            for (int h = 0; h < H; ++h) o = fmaf(w_sm[h], cb[h * dh + d], o);                     // poolings.py:68-69
            out[static_cast<size_t>(b) * dh + d] = o;
        }
}

// ---------------------------------------------------------------------------------- host side
struct DmhaPlan2 { int ok, G, NV, S, fps, stages, FB, ragged; };

static DmhaPlan2 dmha_make_plan2(int x_dtype, int T, int D, int H) {
    DmhaPlan2 pl{};
    const bool bf16 = x_dtype == 1;
    const int VE = bf16 ? 8 : 4, nvmax = bf16 ? 3 : 5;        // <= 20-24 elements of a row per lane (96 registers, 2 CTAs/SM)
    if (H <= 0 || D <= 0 || D % H != 0) return pl;
    const int dh = D / H;
    if (dh % VE != 0) return pl;
    const int nvec = dh / VE;
    int G = 2;
    while (G <= 32 && (nvec + G - 1) / G > nvmax) G <<= 1;
    if (G > 32) return pl;
    const int ngrp = kDmhaConsumerThreads / G;
    if (H > ngrp) return pl;
    int S = ngrp / H;
    if (S > 8) S = 8;
    const size_t frame_bytes = static_cast<size_t>(D) * (bf16 ? 2 : 4);
    int fps = static_cast<int>((16 * 1024) / frame_bytes) / S * S;
    if (fps < S) fps = S;
    const int tcap = (T + S - 1) / S * S;
    if (fps > tcap) fps = tcap > 0 ? tcap : S;
    pl.ok = 1; pl.G = G; pl.NV = (nvec + G - 1) / G; pl.S = S; pl.fps = fps; pl.stages = 4;
    // tuning overrides for sweeps (scripts/sweep_dmha.py); unset in production
    if (const char* e = getenv("DASV_DMHA_FPS")) { const int v = atoi(e); if (v > 0) pl.fps = (v + S - 1) / S * S; }
    if (const char* e = getenv("DASV_DMHA_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= 16) pl.stages = v; }
    pl.ragged = (pl.G * pl.NV != nvec);                  // some lanes' vector slots fall outside the row
    pl.FB = (!pl.ragged && pl.fps % (2 * S) == 0) ? 2 : 1;
    if (const char* e = getenv("DASV_DMHA_FB")) { if (atoi(e) == 1) pl.FB = 1; }
    return pl;
}

template <typename Kern>
static int launch_fwd2_kernel(Kern kern, DmhaFwdParams& p, size_t smem, void* workspace, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("dmha_fwd: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return 1; }
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDmhaThreads, smem);
    if (occ < 1) { set_error("dmha_fwd: kernel does not fit on an SM (smem %zu B)", smem); return 1; }
    int grid = sms * occ;
    if (grid > kDmha2MaxGrid) grid = kDmha2MaxGrid;
    if (grid > p.B) grid = p.B;
    p.ws_cnt = nullptr;
    if (workspace != nullptr && !getenv("DASV_DMHA_STATIC")) {
        // dynamic deal: every utterance is claimed through the counter, which starts at zero (the caller zeroes the
        // workspace once; the last CTA to leave puts it back to zero, see dmha_release_counter)
        p.ws_cnt = static_cast<int*>(workspace);
    }
    e = launch_pdl(kern, dim3(grid), dim3(kDmhaThreads), smem, stream, p);
    if (e != cudaSuccess) { set_error("dmha_fwd: launch failed: %s", cudaGetErrorString(e)); return 1; }
    return check_launch("dmha_fwd");
}

template <bool BF16>
static int dispatch_fwd2(const DmhaPlan2& pl, DmhaFwdParams& p, size_t smem, void* ws, cudaStream_t s) {
#define DASV_CASE2(g, nv) \
    if (pl.G == g && pl.NV == nv) { \
        if (pl.ragged) return launch_fwd2_kernel(dmha_fwd2_kernel<BF16, g, nv, 1, true>, p, smem, ws, s); \
        if (pl.FB == 2) return launch_fwd2_kernel(dmha_fwd2_kernel<BF16, g, nv, 2, false>, p, smem, ws, s); \
        return launch_fwd2_kernel(dmha_fwd2_kernel<BF16, g, nv, 1, false>, p, smem, ws, s); \
    }
#define DASV_ROW2(g) DASV_CASE2(g, 1) DASV_CASE2(g, 2) DASV_CASE2(g, 3) \
    if constexpr (!BF16) { DASV_CASE2(g, 4) DASV_CASE2(g, 5) }
    DASV_ROW2(2) DASV_ROW2(4) DASV_ROW2(8) DASV_ROW2(16) DASV_ROW2(32)
#undef DASV_ROW2
#undef DASV_CASE2
    set_error("dmha_fwd: no v2 kernel for G=%d NV=%d", pl.G, pl.NV);
    return 1;
}

size_t dmha_fwd2_workspace_bytes(int B, int D, int H) {
    (void)D; (void)H;
    return B > 0 ? 256 : 0;          // the utterance counter (padded)
}

static int dmha_fwd2_launch_plan(DmhaFwdParams p, int x_dtype, void* workspace, cudaStream_t stream);

int dmha_fwd2_launch(DmhaFwdParams p, int x_dtype, void* workspace, cudaStream_t stream) {
    const bool bf16 = x_dtype == 1;
    p.hs = 1; p.ldx = p.D; p.Hq = p.H;
    // Small batches: one CTA streams its utterance at what ONE SM can ingest (~90 GB/s: a 4 s utterance of the exampleModel,
    // 512 KB, took 17 us at batch 1).  The heads are independent until the attention over heads, so an utterance is cut into hs
    // head groups that stream on hs SMs; the attention over heads follows as a small second kernel over ctx.
    if (p.ctx != nullptr && p.align == nullptr && !getenv("DASV_DMHA_NOSPLIT")) {
        const size_t utt_bytes = static_cast<size_t>(p.T) * p.D * (bf16 ? 2 : 4);
        int hs = 8;
        while (hs > 1 && (p.H % hs != 0 || static_cast<long long>(p.B) * hs > sm_count() || utt_bytes / hs < 32 * 1024 ||
                          !dmha_make_plan2(x_dtype, p.T, p.D / hs, p.H / hs).ok)) hs >>= 1;
        if (hs > 1 && utt_bytes >= 128 * 1024) {
            DmhaFwdParams q = p;
            q.hs = hs; q.ldx = p.D; q.Hq = p.H;
            q.B = p.B * hs; q.H = p.H / hs; q.D = p.D / hs;
            q.att = nullptr; q.keep = nullptr; q.out = nullptr; q.headw = nullptr;      // the kernel stops at ctx / lse
            const int r = dmha_fwd2_launch_plan(q, x_dtype, nullptr, stream);          // static deal: CTA i owns pseudo-utterance i
            if (r != 0) return r;
            if (p.att == nullptr) return 0;
            cudaError_t e = launch_pdl(dmha_heads_kernel, dim3(p.B), dim3(256), static_cast<size_t>(2 * p.H) * sizeof(float), stream,
                                       static_cast<const float*>(p.ctx), p.att, p.keep, p.out, p.headw, p.H, p.dh);
            if (e != cudaSuccess) { set_error("dmha_fwd: heads kernel launch failed: %s", cudaGetErrorString(e)); return 1; }
            return check_launch("dmha_fwd");
        }
    }
    return dmha_fwd2_launch_plan(p, x_dtype, workspace, stream);
}

static int dmha_fwd2_launch_plan(DmhaFwdParams p, int x_dtype, void* workspace, cudaStream_t stream) {
    DmhaPlan2 p2 = dmha_make_plan2(x_dtype, p.T, p.D, p.H);
    if (!p2.ok) return -1;
    const bool bf16 = x_dtype == 1;
    const uint32_t stage_bytes = static_cast<uint32_t>(p2.fps) * p.D * (bf16 ? 2 : 4);
    size_t smem = dmha2_smem(p.D, p.H, p.dh, p2.S, p2.stages, stage_bytes).total;
    // keep two CTAs per SM when a shallower ring allows it -- unless there are no more utterances than SMs: then a CTA has the SM
    // to itself and the ring is as deep as shared memory allows (a lone CTA streams at bytes-in-flight / latency)
    if (!getenv("DASV_DMHA_STAGES")) {
        if (p.B <= sm_count()) {
            p2.stages = 8;
            smem = dmha2_smem(p.D, p.H, p.dh, p2.S, p2.stages, stage_bytes).total;
            while (smem > 220 * 1024 && p2.stages > 3) smem = dmha2_smem(p.D, p.H, p.dh, p2.S, --p2.stages, stage_bytes).total;
        }
        else while (smem > 113 * 1024 && p2.stages > 3) smem = dmha2_smem(p.D, p.H, p.dh, p2.S, --p2.stages, stage_bytes).total;
    }
    while (smem > 227 * 1024 && p2.stages > 2) smem = dmha2_smem(p.D, p.H, p.dh, p2.S, --p2.stages, stage_bytes).total;
    if (smem > 227 * 1024) return -1;
    p.fps = p2.fps; p.stages = p2.stages; p.S = p2.S;
    return bf16 ? dispatch_fwd2<true>(p2, p, smem, workspace, stream) : dispatch_fwd2<false>(p2, p, smem, workspace, stream);
}

}  // namespace dasv

extern "C" size_t dasv_dmha_fwd_workspace_bytes(int B, int T, int D, int H) {
    (void)T;
    return dasv::dmha_fwd2_workspace_bytes(B, D, H);
}
