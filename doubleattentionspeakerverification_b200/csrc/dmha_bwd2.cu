// Fused DoubleMHA pooling backward, v2 mapping (the default path of dasv_dmha_bwd): same work decomposition as
// dmha_fwd2.cu -- a lane owns NV 16-byte vectors of a (frame, head) row, G = 2..32 lanes per row, packed f32x2 math.
// One pass that reads x and writes dx, with p = exp2(s - lse) recomputed from the saved logsumexp:
//   dw = <g,c_h>; du = w (dw - sum w dw); dc_h = w_h g + du_h a (+ g_ctx); datt = sum du_h c_h
//   dp = <dc_h, x>; ds = p (dp - <dc_h,c_h>); dx = p dc_h + ds q_h/sqrt(H); dq_h = sum ds x / sqrt(H)
// Roofline: HBM, 2x the forward's bytes (read x, write dx).  dquery / datt are reduced deterministically: per-CTA
// partials in the workspace, then dmha_bwd_reduce_kernel (dmha_bwd.cu) in fixed order.
#include "dmha_common.cuh"
#include <math.h>

namespace dasv {

DASV_DEVICE uint64_t bwd_mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
DASV_DEVICE uint64_t bwd_pack_u32x2(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
DASV_DEVICE void bwd_unpack_u32x2(uint64_t v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}

template <int VE, bool BF16>
DASV_DEVICE void bwd_load_row_pairs(const unsigned char* p, uint64_t (&x2)[VE / 2]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    if constexpr (BF16) {
        x2[0] = bwd_pack_u32x2(v.x << 16, v.x & 0xFFFF0000u);
        x2[1] = bwd_pack_u32x2(v.y << 16, v.y & 0xFFFF0000u);
        x2[2] = bwd_pack_u32x2(v.z << 16, v.z & 0xFFFF0000u);
        x2[3] = bwd_pack_u32x2(v.w << 16, v.w & 0xFFFF0000u);
    } else {
        x2[0] = bwd_pack_u32x2(v.x, v.y);
        x2[1] = bwd_pack_u32x2(v.z, v.w);
    }
}
template <int VE, bool BF16>
DASV_DEVICE void bwd_store_row_pairs(unsigned char* p, const uint64_t (&o2)[VE / 2]) {
    uint4 v;
    if constexpr (BF16) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) unpack_f32x2(o2[e], f[2 * e], f[2 * e + 1]);
        v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]);
        v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    } else {
        bwd_unpack_u32x2(o2[0], v.x, v.y);
        bwd_unpack_u32x2(o2[1], v.z, v.w);
    }
    *reinterpret_cast<uint4*>(p) = v;
}

struct Dmha2BwdSmem {
    uint32_t ring, q, a, dc, dq, da, dw, du, dcc, lse2, bars, total;
};
__host__ __device__ inline Dmha2BwdSmem dmha2_bwd_smem(int D, int H, int dh, int stages, uint32_t stage_bytes) {
    Dmha2BwdSmem s;
    uint32_t o = 0;
    s.ring = o; o += stages * stage_bytes;
    s.q = o;    o += D * 4;
    s.a = o;    o += dh * 4;
    s.dc = o;   o += D * 4;
    s.dq = o;   o += D * 4;
    s.da = o;   o += dh * 4;
    s.dw = o;   o += H * 4;
    s.du = o;   o += H * 4;
    s.dcc = o;  o += H * 4;
    s.lse2 = o; o += H * 4;
    o = (o + 7u) & ~7u;
    s.bars = o; o += 2 * stages * 8;
    s.total = o;
    return s;
}

template <bool BF16, int G, int NV, bool RAGGED>
__global__ void __launch_bounds__(kDmhaThreads, 2) dmha_bwd2_kernel(const DmhaBwdParams p) {
    constexpr int VE = BF16 ? 8 : 4;
    constexpr int VP = VE / 2;
    constexpr uint32_t ES = BF16 ? 2u : 4u;
    constexpr int RPW = 32 / G;
    constexpr int GPP = (G >= 8) ? 1 : 8 / G;
    extern __shared__ __align__(128) unsigned char smem[];

    const int D = p.D, H = p.H, dh = p.dh, S = p.S, T = p.T;
    const uint32_t frame_bytes = static_cast<uint32_t>(D) * ES;
    const uint32_t stage_bytes = p.fps * frame_bytes;
    const Dmha2BwdSmem L = dmha2_bwd_smem(D, H, dh, p.stages, stage_bytes);
    unsigned char* ring = smem + L.ring;
    float* q_sm = reinterpret_cast<float*>(smem + L.q);
    float* a_sm = reinterpret_cast<float*>(smem + L.a);
    float* dc_sm = reinterpret_cast<float*>(smem + L.dc);
    float* dq_sm = reinterpret_cast<float*>(smem + L.dq);
    float* da_sm = reinterpret_cast<float*>(smem + L.da);
    float* dw_sm = reinterpret_cast<float*>(smem + L.dw);
    float* du_sm = reinterpret_cast<float*>(smem + L.du);
    float* dcc_sm = reinterpret_cast<float*>(smem + L.dcc);
    float* lse2_sm = reinterpret_cast<float*>(smem + L.lse2);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const bool has_head = p.att != nullptr;

    if (tid == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kDmhaConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kDmhaConsumerWarps) {
        if (lane == 0) {                            // producer: HBM -> SMEM ring
            int st = 0;
            uint32_t ph = 0;
            for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
                int Lb = p.lengths ? p.lengths[b] : T;
                Lb = max(0, min(Lb, T));
                const unsigned char* xb = p.x + static_cast<size_t>(b) * T * frame_bytes;
                for (int f0 = 0; f0 < Lb; f0 += p.fps) {
                    mbar_wait(&empty[st], ph ^ 1u);
                    const uint32_t bytes = static_cast<uint32_t>(min(p.fps, Lb - f0)) * frame_bytes;
                    mbar_arrive_expect_tx(&full[st], bytes);
                    bulk_g2s(ring + st * stage_bytes, xb + static_cast<size_t>(f0) * frame_bytes, bytes, &full[st]);
                    if (++st == p.stages) { st = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    for (int i = tid; i < D; i += kDmhaConsumerThreads) {
        const int h = i / dh, d = i - h * dh;
        q_sm[i] = p.query[d * H + h];
        dq_sm[i] = 0.f;
    }
    for (int i = tid; i < dh; i += kDmhaConsumerThreads) {
        a_sm[i] = has_head ? p.att[i] : 0.f;
        da_sm[i] = 0.f;
    }
    named_bar_sync(1, kDmhaConsumerThreads);

    const int grp = warp * RPW + lane / G, lig = lane % G;
    const int head = grp % H, slot = grp / H;
    const bool active = slot < S;
    const int rot = (lane / G) % GPP;
    uint32_t voff[NV];
    int vidx[NV];
    bool vok[NV];
    uint64_t q2[NV][VP], dq2[NV][VP];
    const uint64_t zero2 = pack_f32x2(0.f, 0.f);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int idx = ((v + rot) % NV) * G + lig;
        vidx[v] = idx;
        vok[v] = active && idx * VE < dh;
        voff[v] = static_cast<uint32_t>(head) * dh * ES + static_cast<uint32_t>(idx) * 16u;
#pragma unroll
        for (int e = 0; e < VP; ++e) {
            q2[v][e] = vok[v] ? pack_f32x2(q_sm[head * dh + idx * VE + 2 * e], q_sm[head * dh + idx * VE + 2 * e + 1]) : zero2;
            dq2[v][e] = zero2;
        }
    }

    int st = 0;
    uint32_t ph = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        int Lb = p.lengths ? p.lengths[b] : T;
        Lb = max(0, min(Lb, T));
        const float* cb = p.ctx + static_cast<size_t>(b) * D;           // ctx[b] as [H][dh]

        // ---------------------------------------------------------------- head-stage backward (tiny, per utterance)
        if (has_head) {
            const float* gb = p.g_out + static_cast<size_t>(b) * dh;
            for (int h = warp; h < H; h += kDmhaConsumerWarps) {
                float dot = 0.f;
                for (int d = lane; d < dh; d += 32) dot = fmaf(gb[d], cb[h * dh + d], dot);
                dot = warp_sum(dot);
                if (lane == 0) dw_sm[h] = dot;
            }
            named_bar_sync(1, kDmhaConsumerThreads);
            if (warp == 0) {
                const float* wb = p.headw + static_cast<size_t>(b) * H;
                float sum = 0.f;
                for (int h = lane; h < H; h += 32) sum = fmaf(wb[h], dw_sm[h], sum);
                sum = warp_sum(sum);
                for (int h = lane; h < H; h += 32) du_sm[h] = wb[h] * (dw_sm[h] - sum);   // masked head: w = 0 -> du = 0
            }
            named_bar_sync(1, kDmhaConsumerThreads);
        }
        for (int h = warp; h < H; h += kDmhaConsumerWarps) {
            const float wh = has_head ? p.headw[static_cast<size_t>(b) * H + h] : 0.f;
            const float duh = has_head ? du_sm[h] : 0.f;
            float dot = 0.f;
            for (int d = lane; d < dh; d += 32) {
                float dcv = 0.f;
                if (has_head) dcv = fmaf(wh, p.g_out[static_cast<size_t>(b) * dh + d], duh * a_sm[d]);
                if (p.g_ctx != nullptr) dcv += p.g_ctx[static_cast<size_t>(b) * D + h * dh + d];
                dc_sm[h * dh + d] = dcv;
                dot = fmaf(dcv, cb[h * dh + d], dot);
            }
            dot = warp_sum(dot);
            if (lane == 0) {
                dcc_sm[h] = dot;
                lse2_sm[h] = p.lse[static_cast<size_t>(b) * H + h] * kLog2e;
            }
        }
        if (has_head) {
            for (int d = tid; d < dh; d += kDmhaConsumerThreads) {
                float s = da_sm[d];
                for (int h = 0; h < H; ++h) s = fmaf(du_sm[h], cb[h * dh + d], s);
                da_sm[d] = s;
            }
        }
        named_bar_sync(1, kDmhaConsumerThreads);

        uint64_t dc2[NV][VP];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < VP; ++e)
                dc2[v][e] = vok[v] ? pack_f32x2(dc_sm[head * dh + vidx[v] * VE + 2 * e], dc_sm[head * dh + vidx[v] * VE + 2 * e + 1]) : zero2;
        const float lse2 = active ? lse2_sm[head] : 0.f;
        const float dcc = active ? dcc_sm[head] : 0.f;

        // ---------------------------------------------------------------- stream x, write dx
        unsigned char* dxb = p.dx + static_cast<size_t>(b) * T * frame_bytes;
        for (int f0 = 0; f0 < Lb; f0 += p.fps) {
            mbar_wait(&full[st], ph);
            const int nf = min(p.fps, Lb - f0);
            const unsigned char* sbase = ring + st * stage_bytes;
            for (int fb = 0; fb < nf; fb += S) {                 // warp-uniform trip count
                const int f = fb + slot;
                const bool valid = active && f < nf;
                const unsigned char* row = sbase + static_cast<uint32_t>(valid ? f : 0) * frame_bytes;
                uint64_t xs[NV][VP];
                uint64_t s2 = zero2, d2 = zero2;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if (RAGGED && !vok[v]) {
#pragma unroll
                        for (int e = 0; e < VP; ++e) xs[v][e] = zero2;
                    } else {
                        bwd_load_row_pairs<VE, BF16>(row + voff[v], xs[v]);
                    }
#pragma unroll
                    for (int e = 0; e < VP; ++e) {
                        s2 = fma_f32x2(xs[v][e], q2[v][e], s2);
                        d2 = fma_f32x2(xs[v][e], dc2[v][e], d2);
                    }
                }
                float s_lo, s_hi, d_lo, d_hi;
                unpack_f32x2(s2, s_lo, s_hi);
                unpack_f32x2(d2, d_lo, d_hi);
                const float sc = group_sum<G>(s_lo + s_hi);
                const float dp = group_sum<G>(d_lo + d_hi);
                const float pr = valid ? fast_exp2(fmaf(sc, p.scale_log2, -lse2)) : 0.f;
                const float ds = pr * (dp - dcc);
                const uint64_t pr2 = pack_f32x2(pr, pr);
                const uint64_t ds2 = pack_f32x2(ds, ds);
                const float dsq = ds * p.inv_sqrt_h;
                const uint64_t dsq2 = pack_f32x2(dsq, dsq);
                unsigned char* orow = dxb + static_cast<size_t>(f0 + f) * frame_bytes;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    uint64_t o2[VP];
#pragma unroll
                    for (int e = 0; e < VP; ++e) {
                        o2[e] = fma_f32x2(pr2, dc2[v][e], bwd_mul_f32x2(dsq2, q2[v][e]));
                        dq2[v][e] = fma_f32x2(ds2, xs[v][e], dq2[v][e]);
                    }
                    if (valid && (!RAGGED || vok[v])) bwd_store_row_pairs<VE, BF16>(orow + voff[v], o2);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
            if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
        // frames beyond the utterance's length receive no gradient
        {
            uint4* z = reinterpret_cast<uint4*>(dxb + static_cast<size_t>(Lb) * frame_bytes);
            const size_t n16 = static_cast<size_t>(T - Lb) * frame_bytes / 16;
            for (size_t i = tid; i < n16; i += kDmhaConsumerThreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        named_bar_sync(1, kDmhaConsumerThreads);    // dc/dcc/lse2/du are rewritten by the next utterance
    }

    // -------------------------------------------------------------------- per-CTA partials of dquery / datt
    // fixed-order accumulation over the S frame slots (deterministic, unlike shared-memory atomics)
    for (int s = 0; s < S; ++s) {
        if (active && slot == s) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (vok[v]) {
#pragma unroll
                    for (int e = 0; e < VP; ++e) {
                        float lo, hi;
                        unpack_f32x2(dq2[v][e], lo, hi);
                        dq_sm[head * dh + vidx[v] * VE + 2 * e] += lo * p.inv_sqrt_h;
                        dq_sm[head * dh + vidx[v] * VE + 2 * e + 1] += hi * p.inv_sqrt_h;
                    }
                }
        }
        named_bar_sync(1, kDmhaConsumerThreads);
    }
    for (int i = tid; i < D; i += kDmhaConsumerThreads) p.ws_dq[static_cast<size_t>(blockIdx.x) * D + i] = dq_sm[i];
    for (int i = tid; i < dh; i += kDmhaConsumerThreads) p.ws_da[static_cast<size_t>(blockIdx.x) * dh + i] = da_sm[i];
}

// ---------------------------------------------------------------------------------- host side
struct DmhaBwdPlan2 { int ok, G, NV, S, fps, stages, ragged; };

static DmhaBwdPlan2 dmha_bwd_make_plan2(int x_dtype, int T, int D, int H) {
    DmhaBwdPlan2 pl{};
    const bool bf16 = x_dtype == 1;
    const int VE = bf16 ? 8 : 4, nvmax = bf16 ? 2 : 5;         // <= 16-20 elements per lane: q, dc, dq and x live in registers
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
    pl.ragged = (pl.G * pl.NV != nvec);
    return pl;
}

template <typename Kern>
static int launch_bwd2_kernel(Kern kern, DmhaBwdParams& p, size_t smem, int max_grid, int* grid_out, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("dmha_bwd: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return 1; }
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDmhaThreads, smem);
    if (occ < 1) { set_error("dmha_bwd: kernel does not fit on an SM (smem %zu B)", smem); return 1; }
    int grid = sms * occ;
    if (grid > p.B) grid = p.B;
    if (grid > max_grid) grid = max_grid;
    *grid_out = grid;
    kern<<<grid, kDmhaThreads, smem, stream>>>(p);
    return check_launch("dmha_bwd");
}

template <bool BF16>
static int dispatch_bwd2(const DmhaBwdPlan2& pl, DmhaBwdParams& p, size_t smem, int max_grid, int* grid, cudaStream_t s) {
#define DASV_CASEB(g, nv) \
    if (pl.G == g && pl.NV == nv) { \
        if (pl.ragged) return launch_bwd2_kernel(dmha_bwd2_kernel<BF16, g, nv, true>, p, smem, max_grid, grid, s); \
        return launch_bwd2_kernel(dmha_bwd2_kernel<BF16, g, nv, false>, p, smem, max_grid, grid, s); \
    }
#define DASV_ROWB(g) DASV_CASEB(g, 1) DASV_CASEB(g, 2) \
    if constexpr (!BF16) { DASV_CASEB(g, 3) DASV_CASEB(g, 4) DASV_CASEB(g, 5) }
    DASV_ROWB(2) DASV_ROWB(4) DASV_ROWB(8) DASV_ROWB(16) DASV_ROWB(32)
#undef DASV_ROWB
#undef DASV_CASEB
    set_error("dmha_bwd: no v2 kernel for G=%d NV=%d", pl.G, pl.NV);
    return 1;
}

int dmha_bwd2_launch(DmhaBwdParams p, int x_dtype, int max_grid, int* grid_out, cudaStream_t stream) {
    DmhaBwdPlan2 p2 = dmha_bwd_make_plan2(x_dtype, p.T, p.D, p.H);
    if (!p2.ok) return -1;
    const bool bf16 = x_dtype == 1;
    const uint32_t stage_bytes = static_cast<uint32_t>(p2.fps) * p.D * (bf16 ? 2 : 4);
    size_t smem = dmha2_bwd_smem(p.D, p.H, p.dh, p2.stages, stage_bytes).total;
    while (smem > 113 * 1024 && p2.stages > 3) smem = dmha2_bwd_smem(p.D, p.H, p.dh, --p2.stages, stage_bytes).total;
    while (smem > 227 * 1024 && p2.stages > 2) smem = dmha2_bwd_smem(p.D, p.H, p.dh, --p2.stages, stage_bytes).total;
    if (smem > 227 * 1024) return -1;
    p.fps = p2.fps; p.stages = p2.stages; p.S = p2.S;
    return bf16 ? dispatch_bwd2<true>(p2, p, smem, max_grid, grid_out, stream)
                : dispatch_bwd2<false>(p2, p, smem, max_grid, grid_out, stream);
}

}  // namespace dasv
