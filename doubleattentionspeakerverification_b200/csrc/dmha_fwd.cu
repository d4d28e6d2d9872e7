// Fused DoubleMHA pooling forward: one pass over x (see include/dasv_b200.h: dasv_dmha_fwd).
// Replaces scripts/poolings.py:73-80 (innerKeyValueAttention), :100-109 (MultiHeadAttention),
// :45-51,:61-71 (HeadAttention narrow path), :126-129 (DoubleMHA.forward).
//
// Roofline: HBM.  Algorithmic bytes per utterance = L_b * D * sizeof(x) read (+ D/H*4 out).
#include "dmha_common.cuh"
#include <math.h>
#include <stdlib.h>

namespace dasv {

// Per-utterance tail shared by both consumer mappings: merge the S partial (max, sum, weighted sum) states of
// every head, finish ctx / lse, then the attention over heads (poolings.py:45-51,61-71) and the alignment fix-up.
DASV_DEVICE void dmha_fwd_finish(const DmhaFwdParams& p, int b, int Lb, int warp, int lane, int tid,
                                 float* pacc, float* pm, float* pl, float* u_sm, float* w_sm, const float* a_sm) {
    const int H = p.H, dh = p.dh, S = p.S, T = p.T;
        for (int h = warp; h < H; h += kDmhaConsumerWarps) {
            float M = -INFINITY;
            for (int s = 0; s < S; ++s) M = fmaxf(M, pm[h * S + s]);
            const float Mref = (M == -INFINITY) ? 0.f : M;
            float Lsum = 0.f;
            for (int s = 0; s < S; ++s) Lsum += pl[h * S + s] * fast_exp2(pm[h * S + s] - Mref);
            const float inv = Lsum > 0.f ? 1.f / Lsum : 0.f;
            float dot = 0.f;
            for (int d = lane; d < dh; d += 32) {
                float c = 0.f;
                for (int s = 0; s < S; ++s) c = fmaf(pacc[(h * S + s) * dh + d], fast_exp2(pm[h * S + s] - Mref), c);
                c *= inv;
                pacc[(h * S) * dh + d] = c;                    // ctx[b,h,d], kept in smem for the head stage
                if (p.ctx != nullptr) p.ctx[(static_cast<size_t>(b) * H + h) * dh + d] = c;
                if (p.att != nullptr) dot = fmaf(c, a_sm[d], dot);
            }
            dot = warp_sum(dot);
            __syncwarp();
            if (lane == 0) {
                u_sm[h] = dot;                                  // poolings.py:47 (no scale)
                const float lse2 = M + log2f(Lsum);             // log2 units; -inf for an empty utterance
                pm[h * S] = lse2;
                if (p.lse != nullptr) p.lse[static_cast<size_t>(b) * H + h] = lse2 * kLn2;
            }
        }
        named_bar_sync(1, kDmhaConsumerThreads);

        if (p.att != nullptr) {
            if (warp == 0) {
                // softmax over heads (poolings.py:50), with the training-mode keep mask (poolings.py:42)
                float mx = -INFINITY;
                for (int h = lane; h < H; h += 32) {
                    const bool kept = p.keep == nullptr || p.keep[static_cast<size_t>(b) * H + h] != 0;
                    const float u = kept ? u_sm[h] : -INFINITY;
                    u_sm[h] = u;
                    mx = fmaxf(mx, u);
                }
                mx = warp_max(mx);
                float sum = 0.f;
                for (int h = lane; h < H; h += 32) {
                    const float e = expf(u_sm[h] - mx);        // all heads dropped -> NaN, as in the reference
                    w_sm[h] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                for (int h = lane; h < H; h += 32) {
                    const float w = w_sm[h] / sum;
                    w_sm[h] = w;
                    if (p.headw != nullptr) p.headw[static_cast<size_t>(b) * H + h] = w;
                }
            }
            named_bar_sync(1, kDmhaConsumerThreads);
            if (p.out != nullptr) {
                for (int d = tid; d < dh; d += kDmhaConsumerThreads) {
                    float o = 0.f;
                    for (int h = 0; h < H; ++h) o = fmaf(w_sm[h], pacc[(h * S) * dh + d], o);   // poolings.py:68-69
                    p.out[static_cast<size_t>(b) * dh + d] = o;
                }
            }
        }
        if (p.align != nullptr) {
            // alignment = softmax over time (poolings.py:77): exp2(raw - lse); frames >= L are 0
            float* ab = p.align + static_cast<size_t>(b) * T * H;
            for (int i = tid; i < T * H; i += kDmhaConsumerThreads) {
                const int t = i / H, h = i - t * H;
                ab[i] = (t < Lb) ? fast_exp2(ab[i] - pm[h * S]) : 0.f;
            }
        }
        named_bar_sync(1, kDmhaConsumerThreads);   // pacc/pm/u/w are reused by the next utterance
}

// Resident CTAs per SM the register allocator must allow: the light configurations want >= 4
// co-resident utterances per SM so that a batch of a few hundred utterances is one wave.
template <bool BF16, int NV, int HPG>
constexpr int dmha_fwd_min_ctas() { return (NV == 1 && HPG == 1) ? (BF16 ? 3 : 4) : 1; }

template <bool BF16, int G, int NV, int HPG>
__global__ void __launch_bounds__(kDmhaThreads, dmha_fwd_min_ctas<BF16, NV, HPG>()) dmha_fwd_kernel(const DmhaFwdParams p) {
    constexpr int VE = BF16 ? 8 : 4;
    constexpr int FB = kDmhaFB;
    constexpr int NG = kDmhaConsumerThreads / G;
    extern __shared__ __align__(128) unsigned char smem[];

    const int D = p.D, H = p.H, dh = p.dh, S = p.S, T = p.T;
    const uint32_t frame_bytes = static_cast<uint32_t>(D) * (BF16 ? 2u : 4u);
    const uint32_t stage_bytes = p.fps * frame_bytes;
    const DmhaFwdSmem L = dmha_fwd_smem(D, H, dh, S, p.stages, stage_bytes);
    unsigned char* ring = smem + L.ring;
    float* q_sm = reinterpret_cast<float*>(smem + L.q);       // [H][dh]  (query transposed)
    float* a_sm = reinterpret_cast<float*>(smem + L.a);       // [dh]
    float* pacc = reinterpret_cast<float*>(smem + L.pacc);    // [H*S][dh] partial weighted sums -> ctx
    float* pm = reinterpret_cast<float*>(smem + L.pm);        // [H*S] running max (log2 units)
    float* pl = reinterpret_cast<float*>(smem + L.pl);        // [H*S] running sum
    float* u_sm = reinterpret_cast<float*>(smem + L.u);       // [H] head scores
    float* w_sm = reinterpret_cast<float*>(smem + L.w);       // [H] head weights
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
    for (int i = tid; i < D; i += kDmhaThreads) {
        const int h = i / dh, d = i - h * dh;
        q_sm[i] = p.query[d * H + h];               // reference layout [dh, H] (poolings.py:90)
    }
    if (p.att != nullptr)
        for (int i = tid; i < dh; i += kDmhaThreads) a_sm[i] = p.att[i];
    __syncthreads();

    if (warp == kDmhaConsumerWarps) {
        // ------------------------------------------------------------ producer: HBM -> SMEM ring
        if (lane == 0) {
            uint32_t it = 0;
            for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
                int Lb = p.lengths ? p.lengths[b] : T;
                Lb = max(0, min(Lb, T));
                const unsigned char* xb = p.x + static_cast<size_t>(b) * T * frame_bytes;
                for (int f0 = 0; f0 < Lb; f0 += p.fps, ++it) {
                    const int st = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1u;
                    mbar_wait(&empty[st], ph ^ 1u);
                    const uint32_t bytes = static_cast<uint32_t>(min(p.fps, Lb - f0)) * frame_bytes;
                    mbar_arrive_expect_tx(&full[st], bytes);
                    bulk_g2s(ring + st * stage_bytes, xb + static_cast<size_t>(f0) * frame_bytes, bytes, &full[st]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    const int gid = tid / G, lig = tid % G;
    // HPG == 1: head = gid % H, frame split = gid / H (groups beyond H*S idle).  HPG > 1: S == 1.
    const int my_split = (HPG == 1) ? gid / H : 0;
    const bool group_active = (HPG == 1) ? (my_split < S) : true;
    int head_of[HPG];
#pragma unroll
    for (int k = 0; k < HPG; ++k) head_of[k] = (HPG == 1) ? (gid % H) : (gid + k * NG);

    bool vec_ok[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) vec_ok[v] = (lig + v * G) * VE < dh;

    float qreg[NV][VE];
    if constexpr (HPG == 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < VE; ++e)
                qreg[v][e] = vec_ok[v] ? q_sm[head_of[0] * dh + (lig + v * G) * VE + e] : 0.f;
    }

    uint32_t it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        int Lb = p.lengths ? p.lengths[b] : T;
        Lb = max(0, min(Lb, T));

        float m[HPG], l[HPG], acc[HPG][NV][VE];
#pragma unroll
        for (int k = 0; k < HPG; ++k) {
            m[k] = -INFINITY; l[k] = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int e = 0; e < VE; ++e) acc[k][v][e] = 0.f;
        }

        for (int f0 = 0; f0 < Lb; f0 += p.fps, ++it) {
            const int st = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1u;
            mbar_wait(&full[st], ph);
            const int nf = min(p.fps, Lb - f0);
            const unsigned char* sbase = ring + st * stage_bytes;
            // first frame of this stage that belongs to my split: (f0 + f) % S == my_split
            const int fstart = (S == 1) ? 0 : ((my_split - (f0 % S) + S) % S);
            const int nbatches = (nf + FB * S - 1) / (FB * S);     // warp-uniform trip count

#pragma unroll
            for (int k = 0; k < HPG; ++k) {
                const int h = head_of[k];
                const bool head_ok = group_active && (h < H);
                if constexpr (HPG > 1) {
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int e = 0; e < VE; ++e)
                            qreg[v][e] = (head_ok && vec_ok[v]) ? q_sm[h * dh + (lig + v * G) * VE + e] : 0.f;
                }
                for (int bi = 0; bi < nbatches; ++bi) {
                    float xs[FB][NV][VE];
                    float sc[FB];
                    bool ok[FB];
#pragma unroll
                    for (int j = 0; j < FB; ++j) {
                        const int f = fstart + (bi * FB + j) * S;
                        ok[j] = head_ok && (f < nf);
                        float part = 0.f;
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            if (ok[j] && vec_ok[v]) {
                                load_row_vec<VE, BF16>(sbase + static_cast<uint32_t>(f) * frame_bytes +
                                                       (static_cast<uint32_t>(h) * dh + (lig + v * G) * VE) * (BF16 ? 2u : 4u),
                                                       xs[j][v]);
                            } else {
#pragma unroll
                                for (int e = 0; e < VE; ++e) xs[j][v][e] = 0.f;
                            }
#pragma unroll
                            for (int e = 0; e < VE; ++e) part = fmaf(xs[j][v][e], qreg[v][e], part);
                        }
                        sc[j] = part;
                    }
#pragma unroll
                    for (int j = 0; j < FB; ++j) sc[j] = group_sum<G>(sc[j]);
                    float mnew = m[k];
#pragma unroll
                    for (int j = 0; j < FB; ++j) {
                        sc[j] = ok[j] ? sc[j] * p.scale_log2 : -INFINITY;
                        mnew = fmaxf(mnew, sc[j]);
                        if (p.align != nullptr && ok[j] && lig == 0) {
                            const int t = f0 + fstart + (bi * FB + j) * S;
                            p.align[(static_cast<size_t>(b) * T + t) * H + h] = sc[j];   // raw log2-score, rescaled below
                        }
                    }
                    const float mref = (mnew == -INFINITY) ? 0.f : mnew;
                    const float corr = fast_exp2(m[k] - mref);
                    float pj[FB], psum = 0.f;
#pragma unroll
                    for (int j = 0; j < FB; ++j) { pj[j] = fast_exp2(sc[j] - mref); psum += pj[j]; }
                    l[k] = fmaf(l[k], corr, psum);
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int e = 0; e < VE; ++e) {
                            float a = acc[k][v][e] * corr;
#pragma unroll
                            for (int j = 0; j < FB; ++j) a = fmaf(pj[j], xs[j][v][e], a);
                            acc[k][v][e] = a;
                        }
                    m[k] = mnew;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }

        // ------------------------------------------------------------ merge partials, head stage
#pragma unroll
        for (int k = 0; k < HPG; ++k) {
            const int h = head_of[k];
            if (group_active && h < H) {
                const int slot = h * S + my_split;
                if (lig == 0) { pm[slot] = m[k]; pl[slot] = l[k]; }
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (vec_ok[v]) {
#pragma unroll
                        for (int e = 0; e < VE; ++e) pacc[slot * dh + (lig + v * G) * VE + e] = acc[k][v][e];
                    }
            }
        }
        named_bar_sync(1, kDmhaConsumerThreads);

        dmha_fwd_finish(p, b, Lb, warp, lane, tid, pacc, pm, pl, u_sm, w_sm, a_sm);
    }
}

// ---------------------------------------------------------------------------------- host side
DmhaPlan dmha_make_plan(int x_dtype, int T, int D, int H, bool backward) {
    DmhaPlan pl{};
    pl.bf16 = (x_dtype == 1);
    const int VE = pl.bf16 ? 8 : 4;
    if (H <= 0 || D <= 0 || D % H != 0) { pl.err = 1; return pl; }
    const int dh = D / H;
    if (dh % VE != 0) { pl.err = 2; return pl; }
    const int nvec = dh / VE;
    int G = 8;
    while (G < 32 && G < nvec) G <<= 1;
    int NV = (nvec + G - 1) / G;
    if (NV > 4) { pl.err = 3; return pl; }
    if (NV == 3) NV = 4;
    if (NV > 1 && G != 32) { pl.err = 3; return pl; }
    const int NG = kDmhaConsumerThreads / G;
    int HPG = 1, S = 1;
    if (H <= NG) { S = NG / H; if (S > 8) S = 8; }
    else if (H <= 4 * NG) HPG = 4;
    else { pl.err = 4; return pl; }
    pl.G = G; pl.NV = NV; pl.HPG = HPG; pl.S = S;
    const size_t frame_bytes = static_cast<size_t>(D) * (pl.bf16 ? 2 : 4);
    int fps = static_cast<int>((12 * 1024 + frame_bytes - 1) / frame_bytes);
    if (fps < 1) fps = 1;
    if (fps > 16) fps = 16;
    if (fps > T) fps = T > 0 ? T : 1;
    pl.fps = fps;
    pl.stages = 3;
    (void)backward;
    return pl;
}

template <typename Kern>
static int launch_fwd_kernel(Kern kern, const DmhaFwdParams& p, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("dmha_fwd: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return 1; }
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDmhaThreads, smem);
    if (occ < 1) { set_error("dmha_fwd: kernel does not fit on an SM (smem %zu B)", smem); return 1; }
    int grid = sms * occ;
    if (grid > p.B) grid = p.B;
    kern<<<grid, kDmhaThreads, smem, stream>>>(p);
    return check_launch("dmha_fwd");
}

template <bool BF16>
static int dispatch_fwd(const DmhaPlan& pl, const DmhaFwdParams& p, size_t smem, cudaStream_t s) {
#define DASV_CASE(g, nv, hpg) \
    if (pl.G == g && pl.NV == nv && pl.HPG == hpg) return launch_fwd_kernel(dmha_fwd_kernel<BF16, g, nv, hpg>, p, smem, s);
    DASV_CASE(8, 1, 1) DASV_CASE(8, 1, 4)
    DASV_CASE(16, 1, 1) DASV_CASE(16, 1, 4)
    DASV_CASE(32, 1, 1) DASV_CASE(32, 1, 4)
    DASV_CASE(32, 2, 1) DASV_CASE(32, 2, 4)
    DASV_CASE(32, 4, 1) DASV_CASE(32, 4, 4)
#undef DASV_CASE
    set_error("dmha_fwd: no kernel for G=%d NV=%d HPG=%d", pl.G, pl.NV, pl.HPG);
    return 1;
}

}  // namespace dasv

using namespace dasv;

extern "C" int dasv_dmha_fwd(const void* x, int x_dtype, const int32_t* lengths,
                             const float* query, const float* att, const uint8_t* keep,
                             float* out, float* ctx, float* lse, float* headw, float* align, void* workspace,
                             int B, int T, int D, int H, void* stream) {
    if (B < 0 || T < 0) { set_error("dmha_fwd: negative shape"); return 1; }
    if (B == 0) return 0;
    if (x == nullptr || query == nullptr) { set_error("dmha_fwd: null x/query"); return 1; }
    if (x_dtype != 0 && x_dtype != 1) { set_error("dmha_fwd: bad dtype %d", x_dtype); return 1; }
    if (att == nullptr && (out != nullptr || headw != nullptr)) {
        set_error("dmha_fwd: out/headw need att (att==NULL selects the MultiHeadAttention-only mode)");
        return 1;
    }
    DmhaFwdParams p{};
    p.x = static_cast<const unsigned char*>(x);
    p.lengths = lengths; p.query = query; p.att = att; p.keep = keep;
    p.out = out; p.ctx = ctx; p.lse = lse; p.headw = headw; p.align = align;
    p.B = B; p.T = T; p.D = D; p.H = H; p.dh = H > 0 ? D / H : 0;
    p.ws_cnt = nullptr;
    p.scale_log2 = kLog2e / sqrtf(static_cast<float>(H));     // d_k = query.size(-1) = H (poolings.py:75)
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {   // v2 mapping (dmha_fwd2.cu): 0 = launched, 1 = error, -1 = shape outside the mapping
        const int r = dmha_fwd2_launch(p, x_dtype, workspace, s);
        if (r >= 0) return r;
    }
    // shapes outside the v2 mapping (more heads than row groups, very wide rows): v1 mapping
    DmhaPlan pl = dmha_make_plan(x_dtype, T, D, H, false);
    if (pl.err) {
        set_error("dmha_fwd: unsupported shape D=%d H=%d dtype=%d (plan error %d: need D%%H==0, "
                  "head size multiple of %d and <= 512 elements per 32 lanes, H <= 4*groups)",
                  D, H, x_dtype, pl.err, x_dtype ? 8 : 4);
        return 1;
    }
    const uint32_t stage_bytes = static_cast<uint32_t>(pl.fps) * D * (pl.bf16 ? 2 : 4);
    size_t smem = dmha_fwd_smem(D, H, p.dh, pl.S, pl.stages, stage_bytes).total;
    while (smem > 227 * 1024 && pl.stages > 2) smem = dmha_fwd_smem(D, H, p.dh, pl.S, --pl.stages, stage_bytes).total;   // very wide features: shallower ring
    p.fps = pl.fps; p.stages = pl.stages; p.S = pl.S;
    if (smem > 227 * 1024) { set_error("dmha_fwd: D=%d needs %zu B of shared memory (> 227 KB)", D, smem); return 1; }
    return pl.bf16 ? dispatch_fwd<true>(pl, p, smem, s) : dispatch_fwd<false>(pl, p, smem, s);
}
