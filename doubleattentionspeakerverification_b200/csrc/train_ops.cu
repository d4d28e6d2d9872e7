// Memory-bound pieces of the front-end's backward pass: what autograd does around cuDNN's conv
// backward in the reference (scripts/CNNs.py:72-86: relu / max_pool2d(2, 2, ceil_mode=True) backward, bias gradients, and
// the K = 9 weight gradient of conv11).  Activations and gradients are bf16 NHWC, parameter gradients fp32.
// Every kernel here streams its tensors once; sums are formed per CTA and added in fixed order (deterministic).
#include "common.cuh"

namespace dasv {

// ---------------------------------------------------------------------------------- ReLU backward, in place
// g = y > 0 ? g : 0.  Rows past an utterance's length hold y = 0 (forward masking rule), so they get no gradient.
__global__ void relu_bwd_kernel(uint4* g, const uint4* y, size_t n16) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        uint4 gv = g[i];
        const uint4 yv = y[i];
        const uint32_t ys[4] = {yv.x, yv.y, yv.z, yv.w};
        uint32_t gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // bf16 > 0  <=>  sign bit clear and not (+)zero
            const uint32_t lo_pos = ((ys[k] & 0x8000u) == 0u && (ys[k] & 0x7FFFu) != 0u) ? 0x0000FFFFu : 0u;
            const uint32_t hi_pos = ((ys[k] & 0x80000000u) == 0u && (ys[k] & 0x7FFF0000u) != 0u) ? 0xFFFF0000u : 0u;
            gs[k] &= (lo_pos | hi_pos);
        }
        g[i] = make_uint4(gs[0], gs[1], gs[2], gs[3]);
    }
}

// ---------------------------------------------------------------------------------- max-pool + ReLU backward
// y [B,T,F,C] = relu(conv) before the pool; gp = gradient at the pooled output, either bf16 NHWC [B,T2,F2,C] or
// (REF) fp32 [B,T2,C*F2] with feature c*F2+f (the front-end's output layout, CNNs.py:88-89).  Writes every element of
// g [B,T,F,C]: the window's first maximum (scan order (0,0),(0,1),(1,0),(1,1), as torch) receives gp if it is > 0.
template <bool REF>
__global__ void unpool_relu_bwd_kernel(const void* gp_, const __nv_bfloat16* y, __nv_bfloat16* g, int B, int T, int F, int C) {
    const int T2 = (T + 1) / 2, F2 = (F + 1) / 2, C8 = C / 8;
    const size_t total = static_cast<size_t>(B) * T2 * F2 * C8;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(i % C8);
        size_t r = i / C8;
        const int fp = static_cast<int>(r % F2); r /= F2;
        const int tp = static_cast<int>(r % T2);
        const int b = static_cast<int>(r / T2);
        float gpv[8];
        if (REF) {
            const float* gp = static_cast<const float*>(gp_) + (static_cast<size_t>(b) * T2 + tp) * C * F2;
#pragma unroll
            for (int k = 0; k < 8; ++k) gpv[k] = gp[static_cast<size_t>(c8 * 8 + k) * F2 + fp];
        } else {
            const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(gp_) + ((static_cast<size_t>(b) * T2 + tp) * F2 + fp) * C + c8 * 8);
            gpv[0] = bf16_lo(v.x); gpv[1] = bf16_hi(v.x); gpv[2] = bf16_lo(v.y); gpv[3] = bf16_hi(v.y);
            gpv[4] = bf16_lo(v.z); gpv[5] = bf16_hi(v.z); gpv[6] = bf16_lo(v.w); gpv[7] = bf16_hi(v.w);
        }
        float yv[4][8];
        bool ok[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int t = 2 * tp + (w >> 1), f = 2 * fp + (w & 1);
            ok[w] = t < T && f < F;
            if (ok[w]) {
                const uint4 v = *reinterpret_cast<const uint4*>(y + ((static_cast<size_t>(b) * T + t) * F + f) * C + c8 * 8);
                yv[w][0] = bf16_lo(v.x); yv[w][1] = bf16_hi(v.x); yv[w][2] = bf16_lo(v.y); yv[w][3] = bf16_hi(v.y);
                yv[w][4] = bf16_lo(v.z); yv[w][5] = bf16_hi(v.z); yv[w][6] = bf16_lo(v.w); yv[w][7] = bf16_hi(v.w);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) yv[w][k] = -INFINITY;
            }
        }
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int a = 0;
            float best = yv[0][k];
#pragma unroll
            for (int w = 1; w < 4; ++w)
                if (yv[w][k] > best) { best = yv[w][k]; a = w; }
            arg[k] = best > 0.f ? a : -1;                       // ReLU backward: a zero maximum passes nothing
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (!ok[w]) continue;
            const int t = 2 * tp + (w >> 1), f = 2 * fp + (w & 1);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = pack_bf16(arg[2 * k] == w ? gpv[2 * k] : 0.f, arg[2 * k + 1] == w ? gpv[2 * k + 1] : 0.f);
            *reinterpret_cast<uint4*>(g + ((static_cast<size_t>(b) * T + t) * F + f) * C + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---------------------------------------------------------------------------------- bias gradient: column sums of g [P, C]
// grid (slabs, C / 128); thread = (channel pair, one of 4 row phases); partial [slab][C] -> fixed-order reduce
constexpr int kColThreads = 256;
__global__ void __launch_bounds__(kColThreads) colsum_partial_kernel(const __nv_bfloat16* g, float* part, size_t P, int C, size_t rows_per_slab) {
    __shared__ float red[4][128];
    const int cp = threadIdx.x & 63, ph = threadIdx.x >> 6;
    const int c0 = blockIdx.y * 128 + cp * 2;
    const size_t r0 = static_cast<size_t>(blockIdx.x) * rows_per_slab;
    const size_t r1 = r0 + rows_per_slab < P ? r0 + rows_per_slab : P;
    float s0 = 0.f, s1 = 0.f;
    if (c0 < C)
        for (size_t r = r0 + ph; r < r1; r += 4) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(g + r * C + c0);
            s0 += bf16_lo(v); s1 += bf16_hi(v);
        }
    red[ph][cp * 2] = s0; red[ph][cp * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 128 && blockIdx.y * 128 + threadIdx.x < C)
        part[static_cast<size_t>(blockIdx.x) * C + blockIdx.y * 128 + threadIdx.x] =
            (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}
__global__ void slab_reduce_kernel(const float* part, float* out, int slabs, int n, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < slabs; ++k) s += part[static_cast<size_t>(k) * n + i];
    out[i] = accumulate ? out[i] + s : s;
}

// ---------------------------------------------------------------------------------- conv11 backward (Cin = 1)
// dw[co][ky][kx] = sum_p g[p][co] * x[p + tap],  db[co] = sum_p g[p][co];  x [B,T,F] f32, g [B,T,F,C] bf16, 128 channels per
// grid.y slice.  One CTA per slab of frames of ONE utterance: the slab's input rows (+-1 halo, rows >= L and the border
// columns as zeros) are staged in shared memory; thread = (8 channels, one of 16 pixel phases): one 16-byte load of g per
// pixel feeds 72 FMAs.  Partial [slab][10][C], fixed-order reduce.
constexpr int kC11BwdRows = 8;
__global__ void __launch_bounds__(kColThreads) conv11_bwd_partial_kernel(const float* x, const __nv_bfloat16* g, const int32_t* lengths,
                                                                         float* part, int B, int T, int F, int C, int chunks) {
    extern __shared__ float c11_sm[];
    float* xs = c11_sm;                                         // [kC11BwdRows + 2][F + 2]
    float* red = c11_sm + (kC11BwdRows + 2) * (F + 2);          // [8 warps][10][128]
    const int b = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * kC11BwdRows;
    const int rows = min(kC11BwdRows, T - t0);
    const int L = lengths ? min(max(lengths[b], 0), T) : T;
    const int W2 = F + 2;
    for (int i = threadIdx.x; i < (kC11BwdRows + 2) * W2; i += kColThreads) {
        const int r = i / W2, fc = i - r * W2;
        const int tt = t0 + r - 1, ff = fc - 1;
        xs[i] = (tt >= 0 && tt < L && ff >= 0 && ff < F) ? x[(static_cast<size_t>(b) * T + tt) * F + ff] : 0.f;   // rows >= L count as zero
    }
    __syncthreads();
    const int c8 = threadIdx.x & 15, ph = threadIdx.x >> 4;
    const int c0 = blockIdx.y * 128 + c8 * 8;
    // packed fp32x2 accumulators over channel pairs (fma.rn.f32x2: two FMAs per instruction)
    uint64_t acc2[10][4];
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[k][e] = pack_f32x2(0.f, 0.f);
    const uint64_t one2 = pack_f32x2(1.f, 1.f);
    if (c0 < C)
        for (int p = ph; p < rows * F; p += 16) {
            const int tl = p / F, f = p - tl * F;
            const uint4 v = *reinterpret_cast<const uint4*>(g + ((static_cast<size_t>(b) * T + t0 + tl) * F + f) * C + c0);
            if ((v.x | v.y | v.z | v.w) == 0u) continue;        // ReLU zeros are common
            uint64_t gv2[4];
            gv2[0] = pack_f32x2(bf16_lo(v.x), bf16_hi(v.x)); gv2[1] = pack_f32x2(bf16_lo(v.y), bf16_hi(v.y));
            gv2[2] = pack_f32x2(bf16_lo(v.z), bf16_hi(v.z)); gv2[3] = pack_f32x2(bf16_lo(v.w), bf16_hi(v.w));
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float xv = xs[(tl + dy) * W2 + f + dx];
                    const uint64_t xv2 = pack_f32x2(xv, xv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc2[dy * 3 + dx][e] = fma_f32x2(gv2[e], xv2, acc2[dy * 3 + dx][e]);
                }
#pragma unroll
            for (int e = 0; e < 4; ++e) acc2[9][e] = fma_f32x2(gv2[e], one2, acc2[9][e]);
        }
    float acc[10][8];
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) unpack_f32x2(acc2[k][e], acc[k][2 * e], acc[k][2 * e + 1]);
    // the two phases of a warp (lanes l and l + 16), then the 8 warps through shared memory
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float s = acc[k][e] + __shfl_xor_sync(0xffffffffu, acc[k][e], 16);
            if (lane < 16) red[(warp * 10 + k) * 128 + c8 * 8 + e] = s;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 10 * 128; i += kColThreads) {
        const int k = i / 128, c = i % 128;
        if (blockIdx.y * 128 + c < C) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[(w * 10 + k) * 128 + c];
            part[(static_cast<size_t>(blockIdx.x) * 10 + k) * C + blockIdx.y * 128 + c] = s;
        }
    }
}
// partial [slab][10][C] -> dw [C][9], db [C]; 32 outputs x 8 slab phases per block, fixed summation order
__global__ void __launch_bounds__(256) conv11_bwd_reduce_kernel(const float* part, float* dw, float* db, int slabs, int C, int accumulate) {
    __shared__ float red[8][32];
    const int o = threadIdx.x & 31, ph = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + o;                          // (k, c)
    float s = 0.f;
    if (i < 10 * C)
        for (int sl = ph; sl < slabs; sl += 8) s += part[static_cast<size_t>(sl) * 10 * C + i];
    red[ph][o] = s;
    __syncthreads();
    if (ph == 0 && i < 10 * C) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][o];
        const int k = i / C, c = i % C;
        float* d = k < 9 ? dw + c * 9 + k : db + c;
        *d = accumulate ? *d + t : t;
    }
}

static int grid_for(size_t n, int threads) {
    size_t b = (n + threads - 1) / threads;
    const size_t cap = static_cast<size_t>(sm_count()) * 16;
    return static_cast<int>(b < cap ? (b ? b : 1) : cap);
}

}  // namespace dasv

using namespace dasv;

extern "C" int dasv_relu_bwd_bf16(void* g, const void* y, size_t n, void* stream) {
    if (n == 0) return 0;
    if (!g || !y) { set_error("relu_bwd: null pointer"); return 1; }
    if (n % 8 != 0) { set_error("relu_bwd: element count %zu must be a multiple of 8", n); return 1; }
    relu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint4*>(g), static_cast<const uint4*>(y), n / 8);
    return check_launch("relu_bwd");
}

extern "C" int dasv_unpool_relu_bwd_bf16(const void* gp, int gp_ref_layout_f32, const void* y, void* g, int B, int T, int F, int C, void* stream) {
    if (B <= 0 || T <= 0 || F <= 0) return 0;
    if (!gp || !y || !g) { set_error("unpool_relu_bwd: null pointer"); return 1; }
    if (C % 8 != 0) { set_error("unpool_relu_bwd: C %d must be a multiple of 8", C); return 1; }
    const size_t total = static_cast<size_t>(B) * ((T + 1) / 2) * ((F + 1) / 2) * (C / 8);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (gp_ref_layout_f32)
        unpool_relu_bwd_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(gp, static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(g), B, T, F, C);
    else
        unpool_relu_bwd_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(gp, static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(g), B, T, F, C);
    return check_launch("unpool_relu_bwd");
}

extern "C" size_t dasv_bias_grad_workspace_bytes(int C) { return static_cast<size_t>(sm_count() * 4) * C * sizeof(float); }
extern "C" size_t dasv_conv11_bwd_workspace_bytes(int B, int T, int C) {
    return static_cast<size_t>(B > 0 ? B : 0) * ((T + kC11BwdRows - 1) / kC11BwdRows) * 10 * C * sizeof(float);
}

extern "C" int dasv_bias_grad_bf16(const void* g, float* db, void* workspace, int accumulate, size_t P, int C, void* stream) {
    if (!g || !db || !workspace) { set_error("bias_grad: null pointer"); return 1; }
    if (C % 2 != 0) { set_error("bias_grad: C must be even"); return 1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t slab_cap = static_cast<size_t>(sm_count()) * 4;
    int slabs = static_cast<int>(P < slab_cap ? (P ? P : 1) : slab_cap);
    const size_t rows = (P + slabs - 1) / slabs;
    slabs = static_cast<int>((P + rows - 1) / (rows ? rows : 1));
    if (slabs < 1) slabs = 1;
    colsum_partial_kernel<<<dim3(slabs, (C + 127) / 128), kColThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(g), static_cast<float*>(workspace), P, C, rows);
    if (check_launch("bias_grad")) return 1;
    slab_reduce_kernel<<<(C + 255) / 256, 256, 0, s>>>(static_cast<const float*>(workspace), db, slabs, C, accumulate);
    return check_launch("bias_grad_reduce");
}

extern "C" int dasv_conv11_bwd(const float* x, const void* g, const int32_t* lengths, float* dw, float* db, void* workspace, int accumulate,
                               int B, int T, int F, int C, void* stream) {
    if (!x || !g || !dw || !db || !workspace) { set_error("conv11_bwd: null pointer"); return 1; }
    if (C % 2 != 0) { set_error("conv11_bwd: C must be even"); return 1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (B <= 0 || T <= 0) return 0;
    if (C % 8 != 0) { set_error("conv11_bwd: C must be a multiple of 8"); return 1; }
    const int chunks = (T + kC11BwdRows - 1) / kC11BwdRows;
    const long long slabs_ll = static_cast<long long>(B) * chunks;
    const size_t smem = (static_cast<size_t>(kC11BwdRows + 2) * (F + 2) + 8 * 10 * 128) * sizeof(float);
    if (smem > 200 * 1024 || slabs_ll > 0x7fffffffLL) { set_error("conv11_bwd: shape too large (F=%d)", F); return 1; }
    const int slabs = static_cast<int>(slabs_ll);
    cudaFuncSetAttribute(conv11_bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    conv11_bwd_partial_kernel<<<dim3(slabs, (C + 127) / 128), kColThreads, smem, s>>>(x, static_cast<const __nv_bfloat16*>(g), lengths, static_cast<float*>(workspace), B, T, F, C, chunks);
    if (check_launch("conv11_bwd")) return 1;
    conv11_bwd_reduce_kernel<<<(10 * C + 31) / 32, 256, 0, s>>>(static_cast<const float*>(workspace), dw, db, slabs, C, accumulate);
    return check_launch("conv11_bwd_reduce");
}
