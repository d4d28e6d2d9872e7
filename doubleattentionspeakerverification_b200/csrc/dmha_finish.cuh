// End-of-utterance stage shared by the DoubleMHA forward kernels (dmha_fwd.cu, dmha_fwd2.cu): merge the S frame
// slots of every head, normalise the context vectors, run the attention over heads (poolings.py:45-51, :61-71) and
// write out / ctx / lse / headw / the normalised alignment.  Called by all consumer threads after each of them has
// stored its partial state: pm[h*S+s] (reference max, log2 units), pl[h*S+s] (sum of weights), pacc[(h*S+s)*dh+d]
// (weighted sum).  Ends with a consumer barrier, after which the buffers may be reused.
#pragma once
#include "dmha_common.cuh"

namespace dasv {

template <int NCW = kDmhaConsumerWarps>
DASV_DEVICE void dmha_finish_utterance(const DmhaFwdParams& p, int b, int Lb, int S, float* pacc, float* pm, float* pl,
                                       float* u_sm, float* w_sm, const float* a_sm, int tid, int warp, int lane) {
    const int H = p.H, dh = p.dh, T = p.T;
    constexpr int kWarps = NCW, kThreads = NCW * 32;
    named_bar_sync(1, kThreads);
    for (int h = warp; h < H; h += kWarps) {
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
    named_bar_sync(1, kThreads);
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
        named_bar_sync(1, kThreads);
        if (p.out != nullptr) {
            for (int d = tid; d < dh; d += kThreads) {
                float o = 0.f;
                for (int h = 0; h < H; ++h) o = fmaf(w_sm[h], pacc[(h * S) * dh + d], o);   // poolings.py:68-69
                p.out[static_cast<size_t>(b) * dh + d] = o;
            }
        }
    }
    if (p.align != nullptr) {
        // alignment = softmax over time (poolings.py:77): exp2(raw - lse); frames >= L are 0
        float* ab = p.align + static_cast<size_t>(b) * T * H;
        for (int i = tid; i < T * H; i += kThreads) {
            const int t = i / H, h = i - t * H;
            ab[i] = (t < Lb) ? fast_exp2(ab[i] - pm[h * S]) : 0.f;
        }
    }
    named_bar_sync(1, kThreads);   // pacc/pm/u/w are reused by the next utterance
}

// Same stage with two barriers instead of four (dmha_fwd2.cu): the caller alternates between two copies of the scratch
// buffers (utterance parity), so no trailing barrier is needed before they are reused, and every warp computes the
// softmax over heads itself instead of waiting for warp 0.  u_sm / w_sm / pm / pl / pacc are the copies of this parity.
template <int NCW>
DASV_DEVICE void dmha_finish_utterance2(const DmhaFwdParams& p, int b, int Lb, int S, float* pacc, float* pm, float* pl,
                                        float* u_sm, float* w_sm, const float* a_sm, int tid, int warp, int lane) {
    const int H = p.H, dh = p.dh, T = p.T;
    constexpr int kThreads = NCW * 32;
    named_bar_sync(1, kThreads);
    for (int h = warp; h < H; h += NCW) {
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
            const bool kept = p.keep == nullptr || p.keep[static_cast<size_t>(b) * H + h] != 0;   // poolings.py:42
            u_sm[h] = kept ? dot : -INFINITY;               // poolings.py:47 (no scale)
            const float lse2 = M + log2f(Lsum);             // log2 units; -inf for an empty utterance
            pm[h * S] = lse2;
            if (p.lse != nullptr) p.lse[static_cast<size_t>(b) * H + h] = lse2 * kLn2;
        }
    }
    named_bar_sync(1, kThreads);
    if (p.att != nullptr) {
        // softmax over heads (poolings.py:50), computed by every warp (identical values land in w_sm)
        float mx = -INFINITY;
        for (int h = lane; h < H; h += 32) mx = fmaxf(mx, u_sm[h]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int h = lane; h < H; h += 32) sum += expf(u_sm[h] - mx);      // all heads dropped -> NaN, as in the reference
        sum = warp_sum(sum);
        for (int h = lane; h < H; h += 32) {
            const float w = expf(u_sm[h] - mx) / sum;
            w_sm[h] = w;
            if (warp == 0 && p.headw != nullptr) p.headw[static_cast<size_t>(b) * H + h] = w;
        }
        __syncwarp();
        if (p.out != nullptr) {
            for (int d = tid; d < dh; d += kThreads) {
                float o = 0.f;
                for (int h = 0; h < H; ++h) o = fmaf(w_sm[h], pacc[(h * S) * dh + d], o);   // poolings.py:68-69
                p.out[static_cast<size_t>(b) * dh + d] = o;
            }
        }
    }
    if (p.align != nullptr) {
        // alignment = softmax over time (poolings.py:77): exp2(raw - lse); frames >= L are 0
        float* ab = p.align + static_cast<size_t>(b) * T * H;
        for (int i = tid; i < T * H; i += kThreads) {
            const int t = i / H, h = i - t * H;
            ab[i] = (t < Lb) ? fast_exp2(ab[i] - pm[h * S]) : 0.f;
        }
    }
}

}  // namespace dasv
