// NOT part of libdasv_b200.so since round 2: the warp-MMA (mma.sync) bf16 pooling forward measured equal to the CUDA-core kernel
// (43.4 vs 43.5 us at the microbench shape: both sit on the same launch / ramp / tail floor).  Kept as a worked example of
// ldmatrix + HMMA with hi/lo split fp32 operands.  To build it again: add dmha_fwd3_launch back to dmha_common.cuh / dmha_fwd.cu.
// Fused DoubleMHA pooling forward for bf16 features, v3: the per-row dot products and weighted sums go through the
// warp-level tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32) so that the kernel is bound by HBM, not by
// instruction issue.  Same semantics and outputs as dmha_fwd2.cu (scripts/poolings.py:73-80, :100-109, :45-51, :61-71,
// :126-129), same producer ring / dynamic utterance deal; only the consumer math differs.
//
// Why: with bf16 features the v2 kernel spends one ALU instruction per element on the bf16 -> fp32 conversion on top of
// the packed FMAs (ncu: 56 % issue-active, 0.66 of the HBM peak at B=512,T=200,D=1024,H=16).  Here a warp owns a head
// pair and handles a tile of 16 frames with
//   scores   S[16 frames x 8]  = X[16 frames x dh] * Qs[dh x 8]     two columns of Qs = bf16 hi / lo parts of the fp32
//                                                                   query (q = hi + lo to 2^-17), rest 0
//   context  C[dh x 8]        += X^T[dh x 16 frames] * Ps[16 x 8]   two columns of Ps = bf16 hi / lo parts of the
//                                                                   softmax weights (p = hi + lo to 2^-17)
// i.e. two ldmatrix.x4 (plain for the scores, .trans for the context) and two MMAs per 16x16 block of x.  The first
// version (one head per warp, results in lanes tig == 0 only) executed 215 warp instructions per (tile, head) -- no
// better than the CUDA-core kernel's 240 per 1024 elements -- and ran a serial chain; sharing the MMA columns between
// the two heads of a pair halves the softmax bookkeeping per head.  Products of bf16 pairs are exact in fp32 and accumulation is fp32, so
// the result differs from the CUDA-core kernel only by summation order and the 2^-17 split of p.
// The tensor cores are used as a wide dot-product unit here (7/8 of every MMA is padding); this is not a dense
// contraction and the roofline stays HBM.
//
// Stage layout: frames are copied one bulk copy per frame into rows of pitch D*2 + pad bytes (pitch = 16 mod 128) so
// that the 8 row addresses of an ldmatrix phase fall into different banks.
#include "dmha_common.cuh"
#include "dmha_finish.cuh"
#include <math.h>
#include <stdlib.h>

namespace dasv {

constexpr float kDmha3Lazy = 8.0f;

DASV_DEVICE void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
DASV_DEVICE void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
DASV_DEVICE void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
DASV_DEVICE float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct Dmha3Smem {
    uint32_t ring, a, pacc, pm, pl, u, w, meta, bars, total;
};
__host__ __device__ inline Dmha3Smem dmha3_smem(int H, int dh, int S, int stages, uint32_t stage_bytes) {
    Dmha3Smem s;
    uint32_t o = 0;
    s.ring = o; o += stages * stage_bytes;
    s.a = o;    o += dh * 4;
    s.pacc = o; o += 2 * S * H * dh * 4;             // scratch is double-buffered by utterance parity
    s.pm = o;   o += 2 * S * H * 4;
    s.pl = o;   o += 2 * S * H * 4;
    s.u = o;    o += 2 * H * 4;
    s.w = o;    o += 2 * H * 4;
    o = (o + 15u) & ~15u;
    s.meta = o; o += stages * 16;
    s.bars = o; o += 2 * stages * 8;
    s.total = o;
    return s;
}
__host__ __device__ inline uint32_t dmha3_pitch(int D) {
    const uint32_t row = static_cast<uint32_t>(D) * 2u;
    return row + (144u - row % 128u) % 128u;        // = 16 (mod 128)
}

struct __align__(16) Dmha3Stage {
    int b, t0, nf, Lb;                               // frames [t0, t0 + nf) of utterance b, whose length is Lb
};

// NCW consumer warps.  A warp owns a PAIR of adjacent heads (2 pr, 2 pr + 1) and a frame slot s: of every stage it
// handles the 16-frame tiles j = s, s + S, ...  The two heads share one set of MMA columns: head u's split query sits in
// columns 4u (hi) and 4u + 1 (lo) of its B operand, so after the score MMAs lane (g, tig) with tig = 2u holds the scores
// of frames g and g + 8 for head u, and the softmax bookkeeping of both heads runs in the same instructions.
// KS = dh / 16.
template <int NCW, int KS>
__global__ void __launch_bounds__((NCW + 1) * 32, NCW <= 8 ? 2 : 1) dmha_fwd3_kernel(const DmhaFwdParams p) {
    constexpr int kThreads = NCW * 32;
    extern __shared__ __align__(128) unsigned char smem[];

    const int D = p.D, H = p.H, dh = p.dh, T = p.T, S = p.S;
    const uint32_t frame_bytes = static_cast<uint32_t>(D) * 2u;
    const uint32_t pitch = p.pitch3;
    const uint32_t stage_bytes = p.fps * pitch;
    const Dmha3Smem L = dmha3_smem(H, dh, S, p.stages, stage_bytes);
    unsigned char* ring = smem + L.ring;
    float* a_sm = reinterpret_cast<float*>(smem + L.a);
    float* pacc = reinterpret_cast<float*>(smem + L.pacc);      // [H*S][dh]
    float* pm = reinterpret_cast<float*>(smem + L.pm);
    float* pl = reinterpret_cast<float*>(smem + L.pl);
    float* u_sm = reinterpret_cast<float*>(smem + L.u);
    float* w_sm = reinterpret_cast<float*>(smem + L.w);
    Dmha3Stage* meta = reinterpret_cast<Dmha3Stage*>(smem + L.meta);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // producer warp: lane 0 owns the barriers and the deal, all lanes issue the per-frame bulk copies
        int st = 0;
        uint32_t ph = 0;
        int b = 0;
        if (lane == 0) b = p.ws_cnt ? atomicAdd(p.ws_cnt, 1) : static_cast<int>(blockIdx.x);
        b = __shfl_sync(0xffffffffu, b, 0);
        while (b < p.B) {
            int Lb = p.lengths ? p.lengths[b] : T;
            Lb = max(0, min(Lb, T));
            const unsigned char* xb = p.x + static_cast<size_t>(b) * T * frame_bytes;
            int f0 = 0;
            do {                                    // an empty utterance still gets one (empty) stage so that it is finished
                const int nf = min(p.fps, Lb - f0);
                if (lane == 0) {
                    mbar_wait(&empty[st], ph ^ 1u);
                    meta[st] = Dmha3Stage{b, f0, nf, Lb};
                    if (nf > 0) mbar_arrive_expect_tx(&full[st], static_cast<uint32_t>(nf) * frame_bytes);
                    else mbar_arrive(&full[st]);
                }
                __syncwarp();
                unsigned char* dst = ring + st * stage_bytes;
                if (p.rowcopy3 & 1) {
                    for (int f = lane; f < nf; f += 32)
                        bulk_g2s(dst + f * pitch, xb + static_cast<size_t>(f0 + f) * frame_bytes, frame_bytes, &full[st]);
                } else if (lane == 0 && nf > 0) {
                    bulk_g2s(dst, xb + static_cast<size_t>(f0) * frame_bytes, static_cast<uint32_t>(nf) * frame_bytes, &full[st]);
                }
                if (++st == p.stages) { st = 0; ph ^= 1u; }
                f0 += p.fps;
            } while (f0 < Lb);
            if (lane == 0) b = p.ws_cnt ? atomicAdd(p.ws_cnt, 1) : b + static_cast<int>(gridDim.x);
            b = __shfl_sync(0xffffffffu, b, 0);
        }
        if (lane == 0) {
            mbar_wait(&empty[st], ph ^ 1u);         // terminator stage
            meta[st] = Dmha3Stage{-1, 0, 0, 0};
            mbar_arrive(&full[st]);
            dmha_release_counter(p.ws_cnt);
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    if (p.att != nullptr)
        for (int i = tid; i < dh; i += kThreads) a_sm[i] = p.att[i];

    const int NP = H >> 1;                          // head pairs (H is even on this path)
    const int g = lane >> 2, tig = lane & 3;
    const int pr = warp % NP, slot = warp / NP;
    const bool working = slot < S;                  // warps beyond NP * S only keep the barriers company
    const int ul = tig >> 1;                        // the head of the pair this lane does the softmax for
    const bool lane_active = (tig & 1) == 0;        // the even lane of a head does the bookkeeping (max, sum, alignment)
    const int my_head = 2 * pr + ul;

    // stage the query through shared memory (coalesced global read; the scratch area is free until the first utterance ends)
    for (int i = tid; i < D; i += kThreads) pacc[(i % H) * dh + i / H] = p.query[i];   // [dh, H] (poolings.py:90) -> [H][dh]
    named_bar_sync(1, kThreads);
    // B fragments of the split query of head u: columns n = g in 4u .. 4u + 3 hold hi, lo, hi, lo (the copy lets the
    // lanes tig = 2u + 1 see the same scores as tig = 2u: one of them prepares the hi parts of the weights, the other
    // the lo parts); rows k = 2 tig, 2 tig + 1 (reg 0) and + 8, + 9 (reg 1)
    uint32_t qb[2][KS][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int head = 2 * pr + u;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            float part[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = ks * 16 + 2 * tig + (e & 1) + (e >> 1) * 8;
                const bool mine = (g >> 2) == u;                                  // columns 4u .. 4u + 3 = hi, lo, hi, lo
                const float q = mine ? pacc[head * dh + d] : 0.f;
                const float hi = bf16_round(q);
                part[e] = (g & 1) ? q - hi : hi;
            }
            qb[u][ks][0] = pack_bf16(part[0], part[1]);
            qb[u][ks][1] = pack_bf16(part[2], part[3]);
        }
    }
    named_bar_sync(1, kThreads);

    float m = -INFINITY, l = 0.f;                   // this lane's head: reference max (log2 units), sum of weights
    float acc[2][KS][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[u][ks][e] = 0.f;

    // per-lane byte offsets inside a 16-frame tile for the two ldmatrix flavours
    const uint32_t pair_col = static_cast<uint32_t>(2 * pr * dh) * 2u;
    const uint32_t off_plain = static_cast<uint32_t>(lane & 15) * pitch + static_cast<uint32_t>(lane >> 4) * 16u + pair_col;
    const uint32_t off_trans = static_cast<uint32_t>((lane & 7) + 8 * (lane >> 4)) * pitch + static_cast<uint32_t>((lane >> 3) & 1) * 16u + pair_col;
    const uint32_t head_step = static_cast<uint32_t>(dh) * 2u;
    const uint32_t ring_u32 = smem_u32(ring);
    // where the weights of my B-operand column come from: column g' belongs to head g' >> 2; frame f of that head's tile
    // lives in lane 4 f + 2 (g' >> 2)
    const int src_e = 8 * tig + 2 * (g >> 2) + (g & 1), src_o = src_e + 4;
    const bool col_used = (g & 2) == 0;                                           // g in {0, 1, 4, 5}

    int st = 0, par = 0;
    uint32_t ph = 0;
    while (true) {
        mbar_wait(&full[st], ph);
        const Dmha3Stage sg = meta[st];             // read before the stage is released
        if (sg.b < 0) break;
        const int b = sg.b;
        bool wrote = false;
        if (working && !(p.rowcopy3 & 4)) {
            for (int t0 = slot * 16; t0 < sg.nf; t0 += S * 16) {
                const int nv = min(16, sg.nf - t0);     // valid frames of this tile
                const uint32_t tile_off = st * stage_bytes + t0 * pitch;
                if (nv < 16) {
                    // rows past the utterance hold stale bytes: zero this warp's columns (0 * NaN would poison the MMA)
                    for (int r = nv; r < 16; ++r)
                        for (int c = lane * 8; c < 2 * dh; c += 256)
                            *reinterpret_cast<uint4*>(ring + tile_off + r * pitch + pair_col + c * 2) = make_uint4(0, 0, 0, 0);
                    wrote = true;
                    __syncwarp();
                }
                const uint32_t tb = ring_u32 + tile_off;
                // ---- scores of 16 frames for both heads
                float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t a0[4], a1[4];
                    ldsm_x4(tb + off_plain + ks * 32, a0);
                    ldsm_x4(tb + off_plain + head_step + ks * 32, a1);
                    mma_bf16_16816(c0, a0, qb[0][ks][0], qb[0][ks][1]);
                    mma_bf16_16816(c1, a1, qb[1][ks][0], qb[1][ks][1]);
                }
                float sa = ul ? c1[0] + c1[1] : c0[0] + c0[1];                    // hi + lo parts; frame g
                float sb = ul ? c1[2] + c1[3] : c0[2] + c0[3];                    // frame g + 8
                const bool va = g < nv, vb = g + 8 < nv;
                sa = va ? sa * p.scale_log2 : -INFINITY;                          // log2-unit scores
                sb = vb ? sb * p.scale_log2 : -INFINITY;
                if (p.align != nullptr && lane_active) {
                    float* ab = p.align + (static_cast<size_t>(b) * T + sg.t0 + t0) * H + my_head;
                    if (va) ab[static_cast<size_t>(g) * H] = sa;                  // raw score, normalised at the end
                    if (vb) ab[static_cast<size_t>(g + 8) * H] = sb;
                }
                float mx = fmaxf(sa, sb);                                         // max over the lanes of my head (same tig)
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
                if (mx > m + kDmha3Lazy) {                                        // lazy rescale; first tile: m = -inf -> corr = 0
                    const float corr = fast_exp2(m - mx);
                    l *= corr;
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[u][ks][e] *= corr;    // only my head's columns matter in this lane
                    m = mx;
                }
                const float pa = fast_exp2(sa - m), pb = fast_exp2(sb - m);       // m is finite here: the tile has a valid frame
                l += pa + pb;
                // bf16 hi / lo parts of the weights; low half: frame g, high half: frame g + 8
                const uint32_t hipack = pack_bf16(pa, pb);
                const uint32_t lopack = pack_bf16(pa - bf16_lo(hipack), pb - bf16_hi(hipack));
                const uint32_t mypack = (tig & 1) ? lopack : hipack;
                // B fragment of the split weights (both heads at once): lane (g', tig') holds column g' for frames
                // 2 tig', 2 tig' + 1 (b0) and + 8, + 9 (b1); frame f of head u: hi parts in lane 4 f + 2 u, lo parts in the next
                const uint32_t ev = __shfl_sync(0xffffffffu, mypack, src_e);
                const uint32_t od = __shfl_sync(0xffffffffu, mypack, src_o);
                const uint32_t b0 = col_used ? __byte_perm(ev, od, 0x5410) : 0u;
                const uint32_t b1 = col_used ? __byte_perm(ev, od, 0x7632) : 0u;
                // ---- weighted sums: C_u[d][4u .. 4u + 1] += X_u^T[d][16 frames] * Ps (the other head's columns are ignored)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t a0[4], a1[4];
                    ldsm_x4_trans(tb + off_trans + ks * 32, a0);
                    ldsm_x4_trans(tb + off_trans + head_step + ks * 32, a1);
                    mma_bf16_16816(acc[0][ks], a0, b0, b1);
                    mma_bf16_16816(acc[1][ks], a1, b0, b1);
                }
            }
        }
        if (wrote) fence_proxy_async();             // generic-proxy zero fill before the next bulk copy into this stage
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
        if (sg.t0 + sg.nf < sg.Lb) continue;

        // ============================================================ end of utterance b
        float* pacc_b = pacc + par * (S * H * dh);
        float* pm_b = pm + par * (S * H);
        float* pl_b = pl + par * (S * H);
        if (working) {
            float lsum = l;
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
            if (lane_active) {
                const int sl = my_head * S + slot;
                if (g == 0) { pm_b[sl] = m; pl_b[sl] = lsum; }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    pacc_b[sl * dh + ks * 16 + g] = ul ? acc[1][ks][0] + acc[1][ks][1] : acc[0][ks][0] + acc[0][ks][1];
                    pacc_b[sl * dh + ks * 16 + g + 8] = ul ? acc[1][ks][2] + acc[1][ks][3] : acc[0][ks][2] + acc[0][ks][3];
                }
            }
        }
        m = -INFINITY; l = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[u][ks][e] = 0.f;
        dmha_finish_utterance2<NCW>(p, b, sg.Lb, S, pacc_b, pm_b, pl_b, u_sm + par * H, w_sm + par * H, a_sm, tid, warp, lane);
        par ^= 1;
    }
}

// ---------------------------------------------------------------------------------- host side
template <typename Kern>
static int launch_fwd3_kernel(Kern kern, DmhaFwdParams& p, int threads, size_t smem, void* workspace, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("dmha_fwd3: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return 1; }
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (occ < 1) { set_error("dmha_fwd3: kernel does not fit on an SM (smem %zu B)", smem); return 1; }
    int grid = sms * occ;
    if (grid > p.B) grid = p.B;
    p.ws_cnt = nullptr;
    if (workspace != nullptr && !getenv("DASV_DMHA_STATIC")) {
        p.ws_cnt = static_cast<int*>(workspace);     // zero on entry, zero again on exit (dmha_release_counter)
    }
    kern<<<grid, threads, smem, stream>>>(p);
    return check_launch("dmha_fwd3");
}

// 0 = launched, 1 = error, -1 = shape outside this mapping (the caller falls back to the CUDA-core kernel)
int dmha_fwd3_launch(DmhaFwdParams p, int x_dtype, void* workspace, cudaStream_t stream) {
    // Opt-in (DASV_DMHA_MMA=1): measured on B200 it ties with the CUDA-core kernel (43.4 vs 43.5 us at B=512,T=200,
    // D=1024,H=16) because both sit on the same streaming floor (37.9 us with the math removed), and the CUDA-core
    // kernel keeps plain fp32 arithmetic.
    const char* on = getenv("DASV_DMHA_MMA");
    if (x_dtype != 1 || on == nullptr || atoi(on) == 0) return -1;
    const int H = p.H, D = p.D;
    if (H <= 0 || D <= 0 || D % H != 0 || (H & 1)) return -1;
    const int dh = D / H;
    if (dh % 16 != 0) return -1;
    const int KS = dh / 16;
    if (KS != 1 && KS != 2 && KS != 4 && KS != 8) return -1;
    const int NP = H / 2;
    if (NP > 16) return -1;
    // <= 8 head pairs: 8 consumer warps and two CTAs per SM (one CTA's end-of-utterance stage overlaps the other's
    // streaming); more pairs: 16 warps, one CTA per SM
    int NCW = NP <= 8 ? 8 : 16;
    if (const char* e = getenv("DASV_DMHA3_NCW")) { const int v = atoi(e); if ((v == 8 || v == 16) && NP <= v) NCW = v; }
    int S = NCW / NP;
    if (S > 8) S = 8;
    uint32_t pitch = dmha3_pitch(D);
    p.rowcopy3 = 1;
    int nomath = 0;
    if (const char* e = getenv("DASV_DMHA3_EXP")) {   // experiments: 1 = unpadded rows, 2 = one bulk copy per stage, 4 = consumers skip the math
        const int v = atoi(e);
        if (v & 3) pitch = D * 2;
        if (v & 2) p.rowcopy3 = 0;
        nomath = (v & 4) ? 1 : 0;
    }
    p.pitch3 = static_cast<int>(pitch);
    p.rowcopy3 |= nomath << 2;
    int tiles = static_cast<int>(((NCW == 8 ? 32u : 64u) * 1024u) / (16u * pitch));
    if (tiles < 1) tiles = 1;
    const int tcap = (p.T + 15) / 16;
    if (tiles > tcap) tiles = tcap > 0 ? tcap : 1;
    if (const char* e = getenv("DASV_DMHA3_TILES")) { const int v = atoi(e); if (v > 0) tiles = v; }
    if (S > tiles) S = tiles;                        // no point in more frame slots than tiles per stage
    p.fps = tiles * 16;
    const uint32_t stage_bytes = static_cast<uint32_t>(p.fps) * pitch;
    int stages = NCW == 8 ? 3 : 4;
    const size_t cap = NCW == 8 ? 112 * 1024 : 220 * 1024;
    if (const char* e = getenv("DASV_DMHA3_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= 16) stages = v; }
    size_t smem = dmha3_smem(H, dh, S, stages, stage_bytes).total;
    while (smem > cap && stages > 2) smem = dmha3_smem(H, dh, S, --stages, stage_bytes).total;
    if (smem > cap) return -1;
    p.stages = stages; p.S = S;
    const int threads = (NCW + 1) * 32;
#define DASV_CASE3(ks) \
    if (KS == ks) return NCW == 8 ? launch_fwd3_kernel(dmha_fwd3_kernel<8, ks>, p, threads, smem, workspace, stream) \
                                  : launch_fwd3_kernel(dmha_fwd3_kernel<16, ks>, p, threads, smem, workspace, stream);
    DASV_CASE3(1) DASV_CASE3(2) DASV_CASE3(4) DASV_CASE3(8)
#undef DASV_CASE3
    return -1;
}

}  // namespace dasv
