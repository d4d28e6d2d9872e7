// NOT part of libdasv_b200.so since round 2 (measured slower than the CUDA-core conv11: both are bound by the NHWC write).
// Kept as a worked example of tcgen05 operands built in shared memory by the kernel itself (K = 9 -> 32 by bf16 hi/lo splits).
// conv11 (Conv2d(1, Cout, 3, padding 1) + bias + ReLU, scripts/CNNs.py:72) on the tensor cores.
//
// The CUDA-core kernel (conv_direct.cu) is FMA-pipe bound: 9.4 GFMA per 256-utterance batch take ~0.7 ms against a
// 0.33 ms write roofline (ncu: 'math' throttle is the top stall).  K = 9 is not a tensor-core shape, but K = 32 is:
// with bf16 hi/lo splits  w*x ~= wh*xh + wl*xh + wh*xl  (relative error ~2^-17, i.e. fp32-like) the layer becomes
//   D[pixel, co] = sum_k A[pixel,k] * B[co,k],  A row = [xh(9) | xh(9) | xl(9) | 0(5)],  B row = [wh(9) | wl(9) | wh(9) | 0(5)]
// one tcgen05.mma pair (K = 2 x 16, M = 128 pixels, N = Cout <= 256) per tile.  Pixels sit on the TMEM lanes, so an
// epilogue thread owns one pixel and its Cout channels are TMEM columns: bias + ReLU + bf16 pack in registers, then the
// pixel's NHWC row goes out as 16-byte stores -- no shared-memory staging, no barriers in the epilogue.
// There is no TMA here: the A operand is an im2col of the 1-channel input, built by 128 threads straight into the
// 128-byte-swizzled K-major layout the UMMA descriptor expects (fence.proxy.async before the MMA reads it).  Tiles are
// 128 CONSECUTIVE pixels of one utterance's (t, f) plane.
// MEASURED (B200, B=256,T=400,Cout=128): 762 us vs 673 us for the CUDA-core kernel; a plain fill of the same 2.1 GB
// takes 532 us (pure-write bandwidth is 3.95 TB/s, not the 6.5 TB/s copy figure), so both sit near the write roofline
// and the front-end keeps the CUDA-core kernel.  This entry point stays as a tested alternative.
// Roles (384 threads): warps 0-3 build operands (thread 0 also issues the MMAs), warps 4-11 = two epilogue groups
// taking alternate tiles.
#include "common.cuh"

namespace dasv {

constexpr int kC11Threads = 384;
constexpr int kC11TileM = 128;                  // pixels per tile = UMMA M = TMEM lanes
constexpr uint32_t kC11RowBytes = 128;          // operand rows use the 64-element (128 B) swizzled layout; K = 32 fills half
constexpr int kC11ABufs = 4;
constexpr int kC11PF = 5;                       // prefetch registers per builder thread (patch <= 640 floats)
constexpr int kC11PatchFloats = kC11PF * 128;

struct C11Params {
    const float* x;
    const float* w;
    const float* bias;
    const int32_t* lengths;
    __nv_bfloat16* y;
    int B, T, F, Cout, Npad, n_stages, tiles_per_utt;
};

DASV_DEVICE void c11_tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// Write the 32 K-slots [s0(9) | s1(9) | s2(9) | 0(5)] of one operand row (4 x 16 B chunks) at its swizzled position.
DASV_DEVICE void c11_store_row(unsigned char* tile, int r, const __nv_bfloat16 (&s0)[9], const __nv_bfloat16 (&s1)[9],
                               const __nv_bfloat16 (&s2)[9]) {
    __nv_bfloat16 k[32];
#pragma unroll
    for (int i = 0; i < 9; ++i) { k[i] = s0[i]; k[9 + i] = s1[i]; k[18 + i] = s2[i]; }
#pragma unroll
    for (int i = 27; i < 32; ++i) k[i] = __float2bfloat16_rn(0.f);
    const uint32_t* kw = reinterpret_cast<const uint32_t*>(k);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 v = make_uint4(kw[4 * c], kw[4 * c + 1], kw[4 * c + 2], kw[4 * c + 3]);
        *reinterpret_cast<uint4*>(tile + r * kC11RowBytes + ((c ^ (r & 7)) << 4)) = v;      // SWIZZLE_128B: chunk ^= row % 8
    }
}

__global__ void __launch_bounds__(kC11Threads, 1) conv11_tc_kernel(const C11Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* a_tiles = smem;                                               // kC11ABufs x [128 pixels][128 B]
    unsigned char* b_tile = a_tiles + kC11ABufs * kC11TileM * kC11RowBytes;      // [256 channels][128 B]
    float* bias_sm = reinterpret_cast<float*>(b_tile + 256 * kC11RowBytes);      // [256]
    float* patch = bias_sm + 256;                                                // input rows t_lo-1 .. t_hi+1, F floats each
    uint64_t* a_free = reinterpret_cast<uint64_t*>(patch + kC11PatchFloats);     // [kC11ABufs]
    uint64_t* acc_full = a_free + kC11ABufs;                                     // [4]
    uint64_t* acc_empty = acc_full + 4;                                          // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, F = p.F, Cout = p.Cout;
    const int n_tiles = p.B * p.tiles_per_utt;
    const int plane = T * F;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kC11ABufs; ++i) mbar_init(&a_free[i], 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // B operand: every output channel's [wh | wl | wh | 0] row, built once per CTA; bias to shared memory
    if (threadIdx.x < 128) {
        for (int row = threadIdx.x; row < p.Npad; row += 128) {
            __nv_bfloat16 wh[9], wl[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const float wv = row < Cout ? p.w[row * 9 + i] : 0.f;           // reference layout [Cout,1,3,3]
                wh[i] = __float2bfloat16_rn(wv);
                wl[i] = __float2bfloat16_rn(wv - __bfloat162float(wh[i]));
            }
            c11_store_row(b_tile, row, wh, wl, wh);
        }
        for (int i = threadIdx.x; i < 256; i += 128) bias_sm[i] = i < Cout ? p.bias[i] : 0.f;
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ------------------------------------------------------------ operand builders (+ MMA issue by thread 0)
        const uint32_t idesc = umma_idesc_bf16(kC11TileM, static_cast<uint32_t>(p.Npad));
        const uint64_t b_desc = umma_desc_k128(smem_u32(b_tile));
        // The input rows of the NEXT tile are fetched into registers while the current tile is built, so the global
        // load latency is off the per-tile critical path.  Rows t_lo-1 .. t_hi+1 are one contiguous span of x.
        float pf[kC11PF];
        auto fetch_patch = [&](int tile) {
            const int b = tile / p.tiles_per_utt, p0 = (tile - b * p.tiles_per_utt) * kC11TileM;
            const int L = p.lengths ? min(max(p.lengths[b], 0), T) : T;
            const int g0 = (p0 / F - 1) * F;                    // first element of row t_lo-1 within the utterance's plane
            const float* xb = p.x + static_cast<size_t>(b) * plane;
#pragma unroll
            for (int k = 0; k < kC11PF; ++k) {
                const int g = g0 + static_cast<int>(threadIdx.x) + k * 128;
                pf[k] = (g >= 0 && g < L * F) ? xb[g] : 0.f;    // rows < 0 and >= L count as zero (padding / masking rule)
            }
        };
        if (static_cast<int>(blockIdx.x) < n_tiles) fetch_patch(blockIdx.x);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int b = tile / p.tiles_per_utt, p0 = (tile - b * p.tiles_per_utt) * kC11TileM;
            const int t_lo = p0 / F;
            const uint32_t buf = it % kC11ABufs, bph = (it / kC11ABufs) & 1u;
            (void)b;
            // (`patch` is free: every builder passed the previous tile's last barrier after its final read)
#pragma unroll
            for (int k = 0; k < kC11PF; ++k) patch[threadIdx.x + k * 128] = pf[k];
            if (tile + static_cast<int>(gridDim.x) < n_tiles) fetch_patch(tile + gridDim.x);   // in flight during the build below
            mbar_wait(&a_free[buf], bph ^ 1u);                  // the MMAs that read this A buffer have retired
            named_bar_sync(3, 128);
            unsigned char* at = a_tiles + buf * (kC11TileM * kC11RowBytes);
            {
                const int pix = p0 + threadIdx.x;               // this thread's pixel = operand row
                __nv_bfloat16 xh[9], xl[9];
                const int t = pix / F, f = pix - t * F;
                const float* c = patch + (t - t_lo) * F + f;    // patch row 0 is frame t_lo-1: c points at (t-1, f)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int ff = f + dx - 1;
                        const float v = (pix < plane && ff >= 0 && ff < F) ? c[dy * F + dx - 1] : 0.f;
                        xh[dy * 3 + dx] = __float2bfloat16_rn(v);
                        xl[dy * 3 + dx] = __float2bfloat16_rn(v - __bfloat162float(xh[dy * 3 + dx]));
                    }
                c11_store_row(at, threadIdx.x, xh, xh, xl);
            }
            fence_proxy_async();                                // generic-proxy SMEM writes -> visible to the tensor core
            named_bar_sync(3, 128);
            if (threadIdx.x == 0) {
                const uint32_t as = it % p.n_stages, aph = (it / p.n_stages) & 1u;
                mbar_wait(&acc_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint64_t a_desc = umma_desc_k128(smem_u32(at));
                const uint32_t d_tmem = tmem_base + as * static_cast<uint32_t>(p.Npad);
                umma_bf16(d_tmem, a_desc, b_desc, idesc, 0u);
                umma_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);             // K slots 16..31
                umma_commit(&acc_full[as]);
                umma_commit(&a_free[buf]);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: thread = pixel, TMEM columns = channels
        const int q = warp & 3, grp = (warp - 4) >> 2;          // TMEM lane quarter, epilogue group (alternate tiles)
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int row = q * 32 + lane;                          // pixel within the tile
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            if ((it & 1u) != static_cast<uint32_t>(grp)) continue;
            const int b = tile / p.tiles_per_utt, p0 = (tile - b * p.tiles_per_utt) * kC11TileM;
            const int L = p.lengths ? min(max(p.lengths[b], 0), T) : T;
            const uint32_t as = it % p.n_stages, aph = (it / p.n_stages) & 1u;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t tcol = tmem_base + lane_addr + as * static_cast<uint32_t>(p.Npad);
            const int pix = p0 + row;
            const bool in_plane = pix < plane;
            const bool live = in_plane && (pix / F) < L;        // frames >= L are written as zeros
            __nv_bfloat16* yp = p.y + (static_cast<size_t>(b) * plane + pix) * Cout;
            for (int c0 = 0; c0 < Cout; c0 += 32) {
                uint32_t r[32];
                c11_tmem_ld_x32(tcol + c0, r);
                tc_wait_ld();
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 bb = *reinterpret_cast<const float2*>(bias_sm + c0 + 2 * j);
                    const float v0 = live ? fmaxf(__uint_as_float(r[2 * j]) + bb.x, 0.f) : 0.f;
                    const float v1 = live ? fmaxf(__uint_as_float(r[2 * j + 1]) + bb.y, 0.f) : 0.f;
                    o[j] = pack_bf16(v0, v1);
                }
                if (in_plane) {
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        if (c0 + 8 * v < Cout)                  // Cout is a multiple of 8
                            *reinterpret_cast<uint4*>(yp + c0 + 8 * v) = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace dasv

using namespace dasv;

extern "C" int dasv_conv11_tc_bf16(const float* x, const float* w, const float* bias, const int32_t* lengths,
                                   void* y, int B, int T, int F, int Cout, void* stream) {
    if (!x || !w || !bias || !y) { set_error("conv11_tc: null argument"); return 1; }
    if (Cout <= 0 || Cout % 8 != 0 || Cout > 256) { set_error("conv11_tc: Cout=%d must be a multiple of 8 in 8..256", Cout); return 1; }
    if (F <= 0 || (kC11TileM / F + 4) * F > kC11PatchFloats) { set_error("conv11_tc: F=%d not supported (patch of %d floats)", F, kC11PatchFloats); return 1; }
    if (B <= 0 || T <= 0) return 0;
    if (static_cast<long long>(T) * F > 0x3fffffffLL) { set_error("conv11_tc: utterance too long"); return 1; }
    C11Params p{};
    p.x = x; p.w = w; p.bias = bias; p.lengths = lengths; p.y = static_cast<__nv_bfloat16*>(y);
    p.B = B; p.T = T; p.F = F; p.Cout = Cout;
    p.Npad = (Cout + 15) / 16 * 16;
    p.n_stages = 512 / p.Npad > 4 ? 4 : 512 / p.Npad;
    p.tiles_per_utt = (T * F + kC11TileM - 1) / kC11TileM;
    const size_t smem = kC11ABufs * kC11TileM * kC11RowBytes + 256 * kC11RowBytes + 256 * 4 + kC11PatchFloats * 4 +
                        (kC11ABufs + 8) * 8 + 16 + 1024;
    cudaError_t e = cudaFuncSetAttribute(conv11_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) { set_error("conv11_tc: smem attribute: %s", cudaGetErrorString(e)); return 1; }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n_tiles = static_cast<long long>(B) * p.tiles_per_utt;
    if (n_tiles > 0x7fffffffLL) { set_error("conv11_tc: too many tiles"); return 1; }
    const int grid = static_cast<int>(n_tiles < sms ? n_tiles : sms);
    conv11_tc_kernel<<<grid, kC11Threads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("conv11_tc");
}
