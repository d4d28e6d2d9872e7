// Fused DoubleMHA pooling backward (see include/dasv_b200.h: dasv_dmha_bwd): one pass that reads x
// and writes dx, so train.py's loss.backward() (scripts/train.py:220) can run through the fused
// pooling.  Closed form, with p = exp(s - lse) recomputed from the saved lse:
//   dw = <g,c_h>; du = w (dw - sum w dw); dc_h = w_h g + du_h a (+ g_ctx); datt = sum du_h c_h
//   dp = <dc_h, x>; ds = p (dp - <dc_h,c_h>); dx = p dc_h + ds q_h/sqrt(H); dq_h = sum ds x / sqrt(H)
// Roofline: HBM, 2x the forward's bytes (read x, write dx).
#include "dmha_common.cuh"
#include <math.h>

namespace dasv {

constexpr int kDmhaBwdMaxGrid = 1184;   // a workspace bound (per-CTA partials), not a machine size: grids are min(this, what the device holds)

struct DmhaBwdSmem {
    uint32_t ring, q, a, dc, dq, da, dw, du, dcc, lse2, bars, total;
};

__host__ __device__ inline DmhaBwdSmem dmha_bwd_smem(int D, int H, int dh, int stages, uint32_t stage_bytes) {
    DmhaBwdSmem s;
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

template <int VE, bool BF16>
DASV_DEVICE void store_row_vec(unsigned char* p, const float (&f)[VE]) {
    uint4 v;
    if constexpr (BF16) {
        v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]);
        v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    } else {
        v.x = __float_as_uint(f[0]); v.y = __float_as_uint(f[1]);
        v.z = __float_as_uint(f[2]); v.w = __float_as_uint(f[3]);
    }
    *reinterpret_cast<uint4*>(p) = v;
}

template <bool BF16, int G, int NV, int HPG>
__global__ void __launch_bounds__(kDmhaThreads) dmha_bwd_kernel(const DmhaBwdParams p) {
    constexpr int VE = BF16 ? 8 : 4;
    constexpr int FB = 2;
    constexpr int NG = kDmhaConsumerThreads / G;
    constexpr uint32_t ES = BF16 ? 2u : 4u;
    extern __shared__ __align__(128) unsigned char smem[];

    const int D = p.D, H = p.H, dh = p.dh, S = p.S, T = p.T;
    const uint32_t frame_bytes = static_cast<uint32_t>(D) * ES;
    const uint32_t stage_bytes = p.fps * frame_bytes;
    const DmhaBwdSmem L = dmha_bwd_smem(D, H, dh, p.stages, stage_bytes);
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
    for (int i = tid; i < D; i += kDmhaThreads) {
        const int h = i / dh, d = i - h * dh;
        q_sm[i] = p.query[d * H + h];
        dq_sm[i] = 0.f;
    }
    for (int i = tid; i < dh; i += kDmhaThreads) {
        a_sm[i] = has_head ? p.att[i] : 0.f;
        da_sm[i] = 0.f;
    }
    __syncthreads();

    if (warp == kDmhaConsumerWarps) {
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

    const int gid = tid / G, lig = tid % G;
    const int my_split = (HPG == 1) ? gid / H : 0;
    const bool group_active = (HPG == 1) ? (my_split < S) : true;
    int head_of[HPG];
#pragma unroll
    for (int k = 0; k < HPG; ++k) head_of[k] = (HPG == 1) ? (gid % H) : (gid + k * NG);
    bool vec_ok[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) vec_ok[v] = (lig + v * G) * VE < dh;

    float dqacc[HPG][NV][VE];
#pragma unroll
    for (int k = 0; k < HPG; ++k)
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < VE; ++e) dqacc[k][v][e] = 0.f;

    uint32_t it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        int Lb = p.lengths ? p.lengths[b] : T;
        Lb = max(0, min(Lb, T));
        const float* cb = p.ctx + static_cast<size_t>(b) * D;           // ctx[b] as [H][dh]

        // ---------------------------------------------------------------- head-stage backward
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

        // ---------------------------------------------------------------- stream x, write dx
        unsigned char* dxb = p.dx + static_cast<size_t>(b) * T * frame_bytes;
        for (int f0 = 0; f0 < Lb; f0 += p.fps, ++it) {
            const int st = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1u;
            mbar_wait(&full[st], ph);
            const int nf = min(p.fps, Lb - f0);
            const unsigned char* sbase = ring + st * stage_bytes;
            const int fstart = (S == 1) ? 0 : ((my_split - (f0 % S) + S) % S);
            const int nbatches = (nf + FB * S - 1) / (FB * S);
#pragma unroll
            for (int k = 0; k < HPG; ++k) {
                const int h = head_of[k];
                const bool head_ok = group_active && (h < H);
                float qreg[NV][VE], dcreg[NV][VE];
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int e = 0; e < VE; ++e) {
                        const bool okv = head_ok && vec_ok[v];
                        qreg[v][e] = okv ? q_sm[h * dh + (lig + v * G) * VE + e] : 0.f;
                        dcreg[v][e] = okv ? dc_sm[h * dh + (lig + v * G) * VE + e] : 0.f;
                    }
                const float lse2 = head_ok ? lse2_sm[h] : 0.f;
                const float dcc = head_ok ? dcc_sm[h] : 0.f;
                for (int bi = 0; bi < nbatches; ++bi) {
                    float xs[FB][NV][VE];
                    float sc[FB], dp[FB];
                    bool ok[FB];
#pragma unroll
                    for (int j = 0; j < FB; ++j) {
                        const int f = fstart + (bi * FB + j) * S;
                        ok[j] = head_ok && (f < nf);
                        float ps = 0.f, pd = 0.f;
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            if (ok[j] && vec_ok[v]) {
                                load_row_vec<VE, BF16>(sbase + static_cast<uint32_t>(f) * frame_bytes +
                                                       (static_cast<uint32_t>(h) * dh + (lig + v * G) * VE) * ES, xs[j][v]);
                            } else {
#pragma unroll
                                for (int e = 0; e < VE; ++e) xs[j][v][e] = 0.f;
                            }
#pragma unroll
                            for (int e = 0; e < VE; ++e) {
                                ps = fmaf(xs[j][v][e], qreg[v][e], ps);
                                pd = fmaf(xs[j][v][e], dcreg[v][e], pd);
                            }
                        }
                        sc[j] = ps; dp[j] = pd;
                    }
#pragma unroll
                    for (int j = 0; j < FB; ++j) { sc[j] = group_sum<G>(sc[j]); dp[j] = group_sum<G>(dp[j]); }
#pragma unroll
                    for (int j = 0; j < FB; ++j) {
                        const float pr = ok[j] ? fast_exp2(fmaf(sc[j], p.scale_log2, -lse2)) : 0.f;
                        const float ds = pr * (dp[j] - dcc);
                        const float dsq = ds * p.inv_sqrt_h;
                        const int f = fstart + (bi * FB + j) * S;
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            float o[VE];
#pragma unroll
                            for (int e = 0; e < VE; ++e) {
                                o[e] = fmaf(pr, dcreg[v][e], dsq * qreg[v][e]);
                                dqacc[k][v][e] = fmaf(ds, xs[j][v][e], dqacc[k][v][e]);
                            }
                            if (ok[j] && vec_ok[v])
                                store_row_vec<VE, BF16>(dxb + static_cast<size_t>(f0 + f) * frame_bytes +
                                                        (static_cast<uint32_t>(h) * dh + (lig + v * G) * VE) * ES, o);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
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
#pragma unroll
    for (int k = 0; k < HPG; ++k) {
        const int h = head_of[k];
        if (group_active && h < H) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (vec_ok[v]) {
#pragma unroll
                    for (int e = 0; e < VE; ++e)
                        atomicAdd(&dq_sm[h * dh + (lig + v * G) * VE + e], dqacc[k][v][e] * p.inv_sqrt_h);
                }
        }
    }
    named_bar_sync(1, kDmhaConsumerThreads);
    for (int i = tid; i < D; i += kDmhaConsumerThreads) p.ws_dq[static_cast<size_t>(blockIdx.x) * D + i] = dq_sm[i];
    for (int i = tid; i < dh; i += kDmhaConsumerThreads) p.ws_da[static_cast<size_t>(blockIdx.x) * dh + i] = da_sm[i];
}

// Fixed-order reduction of the per-CTA partials: dquery[d,h] (reference layout [dh,H]) and datt[d].
// Block = 32 outputs x 8 slices of the partial list: coalesced column sums, then the 8 slice sums are added in a
// fixed order, so the result is deterministic (and the kernel takes a few microseconds instead of tens).
constexpr int kBwdRedSlices = 8;
__global__ void __launch_bounds__(32 * kBwdRedSlices) dmha_bwd_reduce_kernel(const float* ws_dq, const float* ws_da, int nparts,
                                                                            float* dquery, float* datt, int D, int H, int dh) {
    __shared__ float red[kBwdRedSlices][33];
    const int tx = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (i < D) {
        for (int c = sl; c < nparts; c += kBwdRedSlices) s += ws_dq[static_cast<size_t>(c) * D + i];
    } else if (i < D + dh) {
        for (int c = sl; c < nparts; c += kBwdRedSlices) s += ws_da[static_cast<size_t>(c) * dh + (i - D)];
    }
    red[sl][tx] = s;
    __syncthreads();
    if (sl == 0) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < kBwdRedSlices; ++k) tot += red[k][tx];
        if (i < D) {
            const int h = i / dh, d = i - h * dh;
            dquery[d * H + h] = tot;
        } else if (i < D + dh && datt != nullptr) {
            datt[i - D] = tot;
        }
    }
}

template <bool BF16, int G, int NV, int HPG>
static int launch_bwd(DmhaBwdParams& p, size_t smem, float* dquery, float* datt, cudaStream_t stream) {
    auto kern = dmha_bwd_kernel<BF16, G, NV, HPG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("dmha_bwd: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return 1; }
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDmhaThreads, smem);
    if (occ < 1) { set_error("dmha_bwd: kernel does not fit on an SM (smem %zu B)", smem); return 1; }
    int grid = sms * occ;
    if (grid > p.B) grid = p.B;
    if (grid > kDmhaBwdMaxGrid) grid = kDmhaBwdMaxGrid;
    kern<<<grid, kDmhaThreads, smem, stream>>>(p);
    if (check_launch("dmha_bwd")) return 1;
    const int n = p.D + p.dh;
    dmha_bwd_reduce_kernel<<<(n + 31) / 32, 32 * kBwdRedSlices, 0, stream>>>(p.ws_dq, p.ws_da, grid, dquery, datt, p.D, p.H, p.dh);
    return check_launch("dmha_bwd_reduce");
}

template <bool BF16>
static int dispatch_bwd(const DmhaPlan& pl, DmhaBwdParams& p, size_t smem, float* dq, float* da, cudaStream_t s) {
#define DASV_CASE(g, nv, hpg) \
    if (pl.G == g && pl.NV == nv && pl.HPG == hpg) return launch_bwd<BF16, g, nv, hpg>(p, smem, dq, da, s);
    DASV_CASE(8, 1, 1) DASV_CASE(8, 1, 4)
    DASV_CASE(16, 1, 1) DASV_CASE(16, 1, 4)
    DASV_CASE(32, 1, 1) DASV_CASE(32, 1, 4)
    DASV_CASE(32, 2, 1) DASV_CASE(32, 2, 4)
    DASV_CASE(32, 4, 1) DASV_CASE(32, 4, 4)
#undef DASV_CASE
    set_error("dmha_bwd: no kernel for G=%d NV=%d HPG=%d", pl.G, pl.NV, pl.HPG);
    return 1;
}

}  // namespace dasv

using namespace dasv;

extern "C" size_t dasv_dmha_bwd_workspace_bytes(int B, int T, int D, int H) {
    (void)T;
    if (B <= 0 || D <= 0 || H <= 0) return 0;
    const size_t parts = static_cast<size_t>(B < kDmhaBwdMaxGrid ? B : kDmhaBwdMaxGrid);
    return parts * (static_cast<size_t>(D) + D / H) * sizeof(float);
}

extern "C" int dasv_dmha_bwd(const void* x, int x_dtype, const int32_t* lengths,
                             const float* query, const float* att,
                             const float* g_out, const float* g_ctx,
                             const float* ctx, const float* lse, const float* headw,
                             void* dx, float* dquery, float* datt, void* workspace,
                             int B, int T, int D, int H, void* stream) {
    if (B <= 0) return 0;
    if (!x || !query || !ctx || !lse || !dx || !dquery || !workspace) { set_error("dmha_bwd: null argument"); return 1; }
    if (x_dtype != 0 && x_dtype != 1) { set_error("dmha_bwd: bad dtype %d", x_dtype); return 1; }
    if (att != nullptr && (!g_out || !headw || !datt)) { set_error("dmha_bwd: att given but g_out/headw/datt missing"); return 1; }
    if (att == nullptr && !g_ctx) { set_error("dmha_bwd: MultiHeadAttention-only mode needs g_ctx"); return 1; }
    DmhaBwdParams p{};
    p.x = static_cast<const unsigned char*>(x);
    p.lengths = lengths; p.query = query; p.att = att; p.g_out = g_out; p.g_ctx = g_ctx;
    p.ctx = ctx; p.lse = lse; p.headw = headw; p.dx = static_cast<unsigned char*>(dx);
    if (H <= 0 || D <= 0 || D % H != 0) { set_error("dmha_bwd: D=%d must be a positive multiple of H=%d", D, H); return 1; }
    p.B = B; p.T = T; p.D = D; p.H = H; p.dh = D / H;
    p.inv_sqrt_h = 1.0f / sqrtf(static_cast<float>(H));
    p.scale_log2 = kLog2e * p.inv_sqrt_h;
    const size_t parts = static_cast<size_t>(B < kDmhaBwdMaxGrid ? B : kDmhaBwdMaxGrid);
    p.ws_dq = static_cast<float*>(workspace);
    p.ws_da = p.ws_dq + parts * D;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {   // v2 mapping (dmha_bwd2.cu): 0 = launched, 1 = error, -1 = shape outside the mapping
        int grid = 0;
        const int r = dmha_bwd2_launch(p, x_dtype, kDmhaBwdMaxGrid, &grid, s);
        if (r == 0) {
            const int n = D + p.dh;
            dmha_bwd_reduce_kernel<<<(n + 31) / 32, 32 * kBwdRedSlices, 0, s>>>(p.ws_dq, p.ws_da, grid, dquery, datt, D, H, p.dh);
            return check_launch("dmha_bwd_reduce");
        }
        if (r > 0) return r;
    }
    DmhaPlan pl = dmha_make_plan(x_dtype, T, D, H, true);
    if (pl.err) { set_error("dmha_bwd: unsupported shape D=%d H=%d dtype=%d (plan error %d)", D, H, x_dtype, pl.err); return 1; }
    const uint32_t stage_bytes = static_cast<uint32_t>(pl.fps) * D * (pl.bf16 ? 2 : 4);
    size_t smem = dmha_bwd_smem(D, H, p.dh, pl.stages, stage_bytes).total;
    while (smem > 227 * 1024 && pl.stages > 2) smem = dmha_bwd_smem(D, H, p.dh, --pl.stages, stage_bytes).total;   // very wide features: shallower ring
    p.fps = pl.fps; p.stages = pl.stages; p.S = pl.S;
    if (smem > 227 * 1024) { set_error("dmha_bwd: D=%d needs %zu B of shared memory (> 227 KB)", D, smem); return 1; }
    return pl.bf16 ? dispatch_bwd<true>(pl, p, smem, dquery, datt, s) : dispatch_bwd<false>(pl, p, smem, dquery, datt, s);
}
