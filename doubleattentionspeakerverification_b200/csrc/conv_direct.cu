// CUDA-core pieces of the VGG front-end (scripts/CNNs.py:56-91):
//   * conv11 (Cin = 1) direct convolution + bias + ReLU -- not a tensor-core shape (K = 9);
//   * fp32 implicit-GEMM conv3x3 + bias + ReLU: the fp32-parity path (1e-4 relative);
//   * 2x2 ceil-mode max-pool on NHWC, optionally writing the front-end's [B,T',C*F'] layout;
//   * weight re-packing for both conv paths.
// Activations are NHWC [B,T,F,C]; rows t >= lengths[b] are written as zero.
#include "common.cuh"
#include <stdlib.h>

namespace dasv {

// ------------------------------------------------------------------------------ conv11 direct
// One CTA per (b, chunk of kC11Rows frames).  A thread owns 8 output channels: their 9x8 weights and bias
// live in registers for the whole chunk, so the inner loop is 9 broadcast LDS + 72 FMA per pixel and the
// kernel is bound by its NHWC output write (the layer's only real traffic).  Threads are laid out
// (channel group fastest), so the 16-byte stores of neighbouring threads form one contiguous run.
constexpr int kC11Rows = 8;            // frames per CTA when the launch is small; 32 when there are CTAs to spare (measured: 0.665 -> 0.615 ms per 256 x 4 s)

template <int OUT>      // 0 = f32, 1 = bf16, 2 = f16 output (the C ABI's dtype codes); 3 = split bf16 [hi(Cout) | lo(Cout)] per pixel
__global__ void __launch_bounds__(256) conv11_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const int32_t* __restrict__ lengths,
                                                           void* __restrict__ y, int B, int T, int F, int Cout, int R, int lazy) {
    extern __shared__ float x_sm[];       // [R + 2][F + 2], zero halo (R = frames per CTA)
    griddep_launch();                     // the next kernel of the stream may start its prologue
    griddep_wait();                       // x and the buffer behind y belong to earlier work of the stream
    const int chunks = (T + R - 1) / R;
    const int b = blockIdx.x / chunks, t0 = (blockIdx.x - b * chunks) * R;
    const int L = lengths ? min(max(lengths[b], 0), T) : T;
    if (lazy && t0 > L) return;           // lazy masking: of the rows >= L the next layer reads row L only (its bottom halo)
    const int rows = min(R, T - t0);
    const int CG = Cout / 8;
    const int PL = 256 / CG;                                  // pixel lanes (CG <= 256 checked by the host)
    const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
    const bool active = pl < PL;
    const int W2 = F + 2;

    for (int i = threadIdx.x; i < (R + 2) * W2; i += blockDim.x) {
        const int r = i / W2, fc = i - r * W2;
        const int tt = t0 + r - 1, ff = fc - 1;
        float v = 0.f;
        if (tt >= 0 && tt < L && ff >= 0 && ff < F) v = x[(static_cast<size_t>(b) * T + tt) * F + ff];
        x_sm[i] = v;                      // input rows >= L count as zero (masking rule)
    }
    uint64_t wr2[9][4], br2[4];             // channel pairs (e, e+1) packed for fma.rn.f32x2
    if (active) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            br2[e >> 1] = pack_f32x2(bias[cg * 8 + e], bias[cg * 8 + e + 1]);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)   // reference layout [Cout,1,3,3]
                wr2[tap][e >> 1] = pack_f32x2(w[(cg * 8 + e) * 9 + tap], w[(cg * 8 + e + 1) * 9 + tap]);
        }
    }
    __syncthreads();
    if (!active) return;
    const size_t row_elems = static_cast<size_t>(F) * Cout;
    // Two horizontally adjacent pixels per iteration (they share 12 of their 18 input taps); the FMAs are packed
    // fma.rn.f32x2 (sm_100) over channel pairs: acc{c,c+1} += x{p,p} * w{c,c+1} -- half the FMA instruction issue.
    const int halfF = F >> 1;                                  // F is even (checked by the host)
    for (int p = pl; p < rows * halfF; p += PL) {
        const int tl = p / halfF, f = (p - tl * halfF) * 2;
        const int t = t0 + tl;
        float xv[3][4];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int j = 0; j < 4; ++j) xv[dy][j] = x_sm[(tl + dy) * W2 + f + j];
        uint64_t accA[4], accB[4];               // pixel f and pixel f+1
#pragma unroll
        for (int e = 0; e < 4; ++e) { accA[e] = br2[e]; accB[e] = br2[e]; }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const uint64_t xa = pack_f32x2(xv[dy][dx], xv[dy][dx]);
                const uint64_t xb = pack_f32x2(xv[dy][dx + 1], xv[dy][dx + 1]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    accA[e] = fma_f32x2(xa, wr2[dy * 3 + dx][e], accA[e]);
                    accB[e] = fma_f32x2(xb, wr2[dy * 3 + dx][e], accB[e]);
                }
            }
        const bool valid = t < L;
        float a0[8], a1[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            unpack_f32x2(accA[e], a0[2 * e], a0[2 * e + 1]);
            unpack_f32x2(accB[e], a1[2 * e], a1[2 * e + 1]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            a0[e] = valid ? fmaxf(a0[e], 0.f) : 0.f;
            a1[e] = valid ? fmaxf(a1[e], 0.f) : 0.f;
        }
        const size_t o = (static_cast<size_t>(b) * T + t) * row_elems + static_cast<size_t>(f) * Cout + cg * 8;
        if (OUT == 3) {
            // fp32x3 mode: v = hi + lo with hi = bf16(v), lo = bf16(v - hi); pixel pitch 2 * Cout
            const size_t o2 = (static_cast<size_t>(b) * T + t) * (2 * row_elems) + static_cast<size_t>(f) * (2 * Cout) + cg * 8;
            uint16_t* yb = static_cast<uint16_t*>(y);
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const float* a = px == 0 ? a0 : a1;
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    hi[e] = pack_bf16(a[2 * e], a[2 * e + 1]);
                    lo[e] = pack_bf16(a[2 * e] - bf16_lo(hi[e]), a[2 * e + 1] - bf16_hi(hi[e]));
                }
                *reinterpret_cast<uint4*>(yb + o2 + px * 2 * Cout) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(yb + o2 + px * 2 * Cout + Cout) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        } else if (OUT != 0) {
            uint4 v;
            v.x = pack16<OUT>(a0[0], a0[1]); v.y = pack16<OUT>(a0[2], a0[3]);
            v.z = pack16<OUT>(a0[4], a0[5]); v.w = pack16<OUT>(a0[6], a0[7]);
            *reinterpret_cast<uint4*>(static_cast<uint16_t*>(y) + o) = v;
            v.x = pack16<OUT>(a1[0], a1[1]); v.y = pack16<OUT>(a1[2], a1[3]);
            v.z = pack16<OUT>(a1[4], a1[5]); v.w = pack16<OUT>(a1[6], a1[7]);
            *reinterpret_cast<uint4*>(static_cast<uint16_t*>(y) + o + Cout) = v;
        } else {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(y) + o);
            op[0] = make_float4(a0[0], a0[1], a0[2], a0[3]);
            op[1] = make_float4(a0[4], a0[5], a0[6], a0[7]);
            float4* oq = reinterpret_cast<float4*>(static_cast<float*>(y) + o + Cout);
            oq[0] = make_float4(a1[0], a1[1], a1[2], a1[3]);
            oq[1] = make_float4(a1[4], a1[5], a1[6], a1[7]);
        }
    }
}

// ------------------------------------------------------------------------------ weight packing
__global__ void pack_w_f32_kernel(const float* __restrict__ w, float* __restrict__ p, int Cout, int Cin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // output index [tap][ci][co]
    if (i >= 9 * Cin * Cout) return;
    const int co = i % Cout, ci = (i / Cout) % Cin, tap = i / (Cout * Cin);
    p[i] = w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap];
}
// [Cout_pad][9][Cin] bf16 (FMT 1) or fp16 (FMT 2), rows co >= Cout are zero (Cout_pad = Cout rounded up to 128).
template <int FMT>
__global__ void pack_w_bf16_kernel(const float* __restrict__ w, uint16_t* __restrict__ p, int Cout, int Cin, int Cout_pad) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // [co][tap][ci]
    if (i >= static_cast<size_t>(Cout_pad) * 9 * Cin) return;
    const int ci = i % Cin, tap = (i / Cin) % 9, co = i / (static_cast<size_t>(Cin) * 9);
    const float v = co < Cout ? w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap] : 0.f;
    p[i] = cvt16_bits<FMT>(v);
}

// ------------------------------------------------------------------------------ fp32 implicit GEMM
// C[pixel, co] = sum_{tap, ci} X[pixel shifted by tap, ci] * Wp[tap][ci][co]; 64x64 tile, 16-deep
// K slices, 256 threads x (4x4) outputs.  Plain fp32 FMAs: this is the 1e-4 parity path.
constexpr int kF32TM = 64, kF32TN = 64, kF32TK = 16;

__global__ void __launch_bounds__(256) conv3x3_f32_kernel(const float* __restrict__ x, const float* __restrict__ wp,
                                                         const float* __restrict__ bias, const int32_t* __restrict__ lengths,
                                                         float* __restrict__ y, int B, int T, int F, int Cin, int Cout) {
    __shared__ float As[kF32TK][kF32TM + 4];
    __shared__ float Bs[kF32TK][kF32TN];
    const long long npix = static_cast<long long>(B) * T * F;
    const long long m0 = static_cast<long long>(blockIdx.x) * kF32TM;
    const int n0 = blockIdx.y * kF32TN;
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;

    // the pixel this thread loads for the A tile
    const int lp = tid / 4, lc = (tid % 4) * 4;
    const long long pix = m0 + lp;
    int pb = 0, pt = 0, pf = 0, pL = 0;
    const bool pix_ok = pix < npix;
    if (pix_ok) {
        pb = static_cast<int>(pix / (static_cast<long long>(T) * F));
        const int r = static_cast<int>(pix - static_cast<long long>(pb) * T * F);
        pt = r / F; pf = r - pt * F;
        pL = lengths ? min(max(lengths[pb], 0), T) : T;
    }
    const int bk = tid / 16, bn = (tid % 16) * 4;     // B tile element this thread loads

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int tt = pt + dy, ff = pf + dx;
        const bool src_ok = pix_ok && tt >= 0 && tt < pL && ff >= 0 && ff < F;
        const float* src = x + ((static_cast<size_t>(pb) * T + tt) * F + ff) * Cin;
        for (int c0 = 0; c0 < Cin; c0 += kF32TK) {
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src_ok && c0 + lc < Cin) av = *reinterpret_cast<const float4*>(src + c0 + lc);
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + bk < Cin && n0 + bn < Cout)
                bv = *reinterpret_cast<const float4*>(wp + (static_cast<size_t>(tap) * Cin + c0 + bk) * Cout + n0 + bn);
            __syncthreads();
            As[lc + 0][lp] = av.x; As[lc + 1][lp] = av.y; As[lc + 2][lp] = av.z; As[lc + 3][lp] = av.w;
            *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kF32TK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
            }
        }
    }
    const int n = n0 + tx * 4;
    if (n >= Cout) return;
    const float4 bv = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long p = m0 + ty * 4 + i;
        if (p >= npix) continue;
        const int b = static_cast<int>(p / (static_cast<long long>(T) * F));
        const int t = static_cast<int>((p - static_cast<long long>(b) * T * F) / F);
        const int L = lengths ? min(max(lengths[b], 0), T) : T;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < L) {
            o.x = fmaxf(acc[i][0] + bv.x, 0.f); o.y = fmaxf(acc[i][1] + bv.y, 0.f);
            o.z = fmaxf(acc[i][2] + bv.z, 0.f); o.w = fmaxf(acc[i][3] + bv.w, 0.f);
        }
        *reinterpret_cast<float4*>(y + static_cast<size_t>(p) * Cout + n) = o;
    }
}

// ------------------------------------------------------------------------------ max-pool 2x2 ceil
template <typename TI>
DASV_DEVICE float ld_act(const TI* p);
template <>
DASV_DEVICE float ld_act<float>(const float* p) { return *p; }
template <>
DASV_DEVICE float ld_act<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
DASV_DEVICE void st_act(float* p, float v) { *p = v; }
DASV_DEVICE void st_act(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TI, typename TO, bool REF>
__global__ void maxpool2x2_kernel(const TI* __restrict__ x, TO* __restrict__ y, int B, int T, int F, int C) {
    const int T2 = (T + 1) / 2, F2 = (F + 1) / 2;
    const size_t n = static_cast<size_t>(B) * T2 * F2 * C;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        int c, f2, t2, b;
        size_t r = i;
        if (REF) { f2 = r % F2; r /= F2; c = r % C; r /= C; }   // feature index = c*F2 + f2 (CNNs.py:88-89)
        else     { c = r % C; r /= C; f2 = r % F2; r /= F2; }
        t2 = r % T2; b = r / T2;
        float m = -INFINITY;
#pragma unroll
        for (int dt = 0; dt < 2; ++dt)
#pragma unroll
            for (int df = 0; df < 2; ++df) {
                const int t = 2 * t2 + dt, f = 2 * f2 + df;
                if (t < T && f < F) m = fmaxf(m, ld_act<TI>(x + ((static_cast<size_t>(b) * T + t) * F + f) * C + c));
            }
        st_act(y + i, m);
    }
}

// bf16 NHWC -> bf16 NHWC, 8 channels (16 bytes) per thread, packed bf16x2 maxima: the training forward's pool
__global__ void maxpool2x2_bf16x8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int T, int F, int C8) {
    const int T2 = (T + 1) / 2, F2 = (F + 1) / 2;
    const size_t n = static_cast<size_t>(B) * T2 * F2 * C8;
    const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C8);
        size_t r = i / C8;
        const int f2 = static_cast<int>(r % F2); r /= F2;
        const int t2 = static_cast<int>(r % T2);
        const int b = static_cast<int>(r / T2);
        __nv_bfloat162 m[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
        for (int dt = 0; dt < 2; ++dt)
#pragma unroll
            for (int df = 0; df < 2; ++df) {
                const int t = 2 * t2 + dt, f = 2 * f2 + df;
                if (t < T && f < F) {
                    const uint4 v = x[((static_cast<size_t>(b) * T + t) * F + f) * C8 + c];
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                    for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], h[k]);
                }
            }
        y[i] = *reinterpret_cast<const uint4*>(m);
    }
}

}  // namespace dasv

using namespace dasv;

static int conv11_launch(const float* x, const float* w, const float* bias, const int32_t* lengths,
                         void* y, int y_dtype, int B, int T, int F, int Cout, void* stream, int lazy) {
    if (!x || !w || !bias || !y) { set_error("conv11_direct: null argument"); return 1; }
    if (Cout % 8 != 0 || Cout <= 0) { set_error("conv11_direct: Cout=%d must be a positive multiple of 8", Cout); return 1; }
    if (y_dtype < 0 || y_dtype > 3) { set_error("conv11_direct: bad dtype %d", y_dtype); return 1; }
    if (Cout > 2048) { set_error("conv11_direct: Cout=%d > 2048", Cout); return 1; }
    if (F % 2 != 0) { set_error("conv11_direct: F=%d must be even", F); return 1; }
    if (B <= 0 || T <= 0) return 0;
    // frames per CTA: 32 amortises the per-CTA prologue (weights to registers, input rows to shared memory) when the launch
    // still has several CTAs per SM; small launches keep 8 so that they spread over the SMs
    int R = kC11Rows;
    while (R < 32 && static_cast<long long>(B) * ((T + 2 * R - 1) / (2 * R)) >= 8LL * sm_count()) R *= 2;
    // ... and fewer when the launch would leave SMs idle: about 200 CTAs is the optimum for 1-8 utterances of 4 s (measured,
    // scripts/ubench/c11_rows.py: one utterance 11.3 us with 8 frames per CTA, 7.8 us with 2, 9.5 us with 1)
    while (R > 1 && static_cast<long long>(B) * ((T + R - 1) / R) < (4LL * sm_count()) / 3) R /= 2;
    if (const char* e = getenv("DASV_C11_ROWS")) { const int v = atoi(e); if (v >= 1 && v <= 32) R = v; }   // tuning override
    size_t smem = static_cast<size_t>(R + 2) * (F + 2) * sizeof(float);
    while (smem > 48 * 1024 && R > 1) { R /= 2; smem = static_cast<size_t>(R + 2) * (F + 2) * sizeof(float); }
    if (smem > 48 * 1024) { set_error("conv11_direct: F=%d too wide", F); return 1; }
    const int chunks = (T + R - 1) / R;
    const unsigned grid = static_cast<unsigned>(B) * chunks;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (y_dtype == 1) e = launch_pdl(conv11_direct_kernel<1>, dim3(grid), dim3(256), smem, s, x, w, bias, lengths, y, B, T, F, Cout, R, lazy);
    else if (y_dtype == 2) e = launch_pdl(conv11_direct_kernel<2>, dim3(grid), dim3(256), smem, s, x, w, bias, lengths, y, B, T, F, Cout, R, lazy);
    else if (y_dtype == 3) e = launch_pdl(conv11_direct_kernel<3>, dim3(grid), dim3(256), smem, s, x, w, bias, lengths, y, B, T, F, Cout, R, lazy);
    else e = launch_pdl(conv11_direct_kernel<0>, dim3(grid), dim3(256), smem, s, x, w, bias, lengths, y, B, T, F, Cout, R, lazy);
    if (e != cudaSuccess) { set_error("conv11_direct: launch failed: %s", cudaGetErrorString(e)); return 1; }
    return check_launch("conv11_direct");
}

extern "C" int dasv_conv11_direct(const float* x, const float* w, const float* bias, const int32_t* lengths,
                                  void* y, int y_dtype, int B, int T, int F, int Cout, void* stream) {
    return conv11_launch(x, w, bias, lengths, y, y_dtype, B, T, F, Cout, stream, 0);
}

// The same layer for pipelines whose consumers read, of the rows at or beyond an utterance's length L, only row L (the 3x3
// kernels that follow: DASV_CONV_LAZY_MASK): rows > L of y may be left unwritten.  Rows < L and row L are as above.
extern "C" int dasv_conv11_direct_lazy(const float* x, const float* w, const float* bias, const int32_t* lengths,
                                       void* y, int y_dtype, int B, int T, int F, int Cout, void* stream) {
    return conv11_launch(x, w, bias, lengths, y, y_dtype, B, T, F, Cout, stream, lengths != nullptr);
}

extern "C" int dasv_pack_conv_weight_f32(const float* w, float* packed, int Cout, int Cin, void* stream) {
    if (!w || !packed) { set_error("pack_conv_weight_f32: null argument"); return 1; }
    const int n = 9 * Cin * Cout;
    pack_w_f32_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, packed, Cout, Cin);
    return check_launch("pack_conv_weight_f32");
}

extern "C" size_t dasv_packed_conv_weight_bf16_elems(int Cout, int Cin) {
    const size_t cp = (static_cast<size_t>(Cout) + 127) / 128 * 128;
    return cp * 9 * static_cast<size_t>(Cin);
}

extern "C" int dasv_pack_conv_weight_16(const float* w, void* packed, int Cout, int Cin, int dtype, void* stream) {
    if (!w || !packed) { set_error("pack_conv_weight_16: null argument"); return 1; }
    if (dtype != 1 && dtype != 2) { set_error("pack_conv_weight_16: dtype %d must be 1 (bf16) or 2 (f16)", dtype); return 1; }
    const int cp = (Cout + 127) / 128 * 128;
    const size_t n = static_cast<size_t>(cp) * 9 * Cin;
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (dtype == 2) pack_w_bf16_kernel<2><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<uint16_t*>(packed), Cout, Cin, cp);
    else pack_w_bf16_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<uint16_t*>(packed), Cout, Cin, cp);
    return check_launch("pack_conv_weight_16");
}

// fp32x3 mode: [Cout_pad][9][3*Cin] bf16 = per tap [hi(w) | hi(w) | lo(w)], to be contracted with activations laid out
// [hi(x) | lo(x)] and read as [hi(x) | lo(x) | hi(x)]:  w*x ~= hi(w)hi(x) + hi(w)lo(x) + lo(w)hi(x)  (error ~2^-16 |w x|).
__global__ void pack_w_x3_kernel(const float* __restrict__ w, uint16_t* __restrict__ p, int Cout, int Cin, int Cout_pad) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // [co][tap][3*Cin]
    if (i >= static_cast<size_t>(Cout_pad) * 27 * Cin) return;
    const int j = i % (3 * Cin), tap = (i / (3 * Cin)) % 9, co = i / (static_cast<size_t>(Cin) * 27);
    const int ci = j % Cin, seg = j / Cin;
    const float v = co < Cout ? w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap] : 0.f;
    const uint16_t hi = cvt_bf16_bits(v);
    p[i] = seg < 2 ? hi : cvt_bf16_bits(v - __uint_as_float(static_cast<uint32_t>(hi) << 16));
}

extern "C" int dasv_pack_conv_weight_x3(const float* w, void* packed, int Cout, int Cin, void* stream) {
    if (!w || !packed) { set_error("pack_conv_weight_x3: null argument"); return 1; }
    const int cp = (Cout + 127) / 128 * 128;
    const size_t n = static_cast<size_t>(cp) * 27 * Cin;
    pack_w_x3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<uint16_t*>(packed), Cout, Cin, cp);
    return check_launch("pack_conv_weight_x3");
}

extern "C" int dasv_pack_conv_weight_bf16(const float* w, void* packed, int Cout, int Cin, void* stream) {
    return dasv_pack_conv_weight_16(w, packed, Cout, Cin, 1, stream);
}

extern "C" int dasv_conv3x3_f32(const float* x, const float* wp, const float* bias, const int32_t* lengths,
                                float* y, int B, int T, int F, int Cin, int Cout, void* stream) {
    if (!x || !wp || !bias || !y) { set_error("conv3x3_f32: null argument"); return 1; }
    if (Cin % 4 != 0 || Cout % 4 != 0) { set_error("conv3x3_f32: Cin=%d and Cout=%d must be multiples of 4", Cin, Cout); return 1; }
    if (B <= 0 || T <= 0) return 0;
    const long long npix = static_cast<long long>(B) * T * F;
    dim3 grid(static_cast<unsigned>((npix + kF32TM - 1) / kF32TM), (Cout + kF32TN - 1) / kF32TN);
    conv3x3_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, wp, bias, lengths, y, B, T, F, Cin, Cout);
    return check_launch("conv3x3_f32");
}

extern "C" int dasv_maxpool2x2(const void* x, int x_dtype, void* y, int y_dtype, int ref_layout,
                               int B, int T, int F, int C, void* stream) {
    if (!x || !y) { set_error("maxpool2x2: null argument"); return 1; }
    if (B <= 0 || T <= 0) return 0;
    const size_t n = static_cast<size_t>(B) * ((T + 1) / 2) * ((F + 1) / 2) * C;
    const unsigned grid = static_cast<unsigned>(n / 256 + 1 < static_cast<size_t>(sm_count()) * 16 ? n / 256 + 1 : static_cast<size_t>(sm_count()) * 16);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define DASV_POOL(TI, TO, REF) \
    maxpool2x2_kernel<TI, TO, REF><<<grid, 256, 0, s>>>(static_cast<const TI*>(x), static_cast<TO*>(y), B, T, F, C)
    if (x_dtype == 0 && y_dtype == 0) { if (ref_layout) DASV_POOL(float, float, true); else DASV_POOL(float, float, false); }
    else if (x_dtype == 1 && y_dtype == 1 && !ref_layout && C % 8 == 0)
        maxpool2x2_bf16x8_kernel<<<grid, 256, 0, s>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), B, T, F, C / 8);
    else if (x_dtype == 1 && y_dtype == 1) { if (ref_layout) DASV_POOL(__nv_bfloat16, __nv_bfloat16, true); else DASV_POOL(__nv_bfloat16, __nv_bfloat16, false); }
    else if (x_dtype == 1 && y_dtype == 0) { if (ref_layout) DASV_POOL(__nv_bfloat16, float, true); else DASV_POOL(__nv_bfloat16, float, false); }
    else { set_error("maxpool2x2: unsupported dtype pair %d -> %d", x_dtype, y_dtype); return 1; }
#undef DASV_POOL
    return check_launch("maxpool2x2");
}
