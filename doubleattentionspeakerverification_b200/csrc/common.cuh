// Shared device helpers for the sm_100a kernels: mbarrier, bulk/TMA copies, tcgen05, small math.
// Everything here is inline PTX written for this project (no CUTLASS/CuTe dependency).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dasv {

#define DASV_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------ error plumbing (host)
void set_error(const char* fmt, ...);
int  check_launch(const char* what);
bool pdl_enabled();   // programmatic dependent launch on this library's launches (DASV_PDL=0 disables)
int  sm_count();   // multiprocessors of the current device (cached per device; 148 on B200), for grid sizing

// Launch `kern` as a programmatic dependent of the stream's previous kernel (only for kernels that call
// griddep_wait() before touching anything an earlier kernel of the stream produced or still reads).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------ address helpers
DASV_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ------------------------------------------------------------------ mbarrier
DASV_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DASV_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DASV_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

DASV_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DASV_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DASV_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (an error the host sees) instead of a hung GPU.
// ~2^28 polls of a hardware-suspending try_wait is many seconds; legitimate waits are microseconds.
DASV_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}

// ------------------------------------------------------------------ bulk (1-D) async copy global -> shared, completes on an mbarrier
DASV_DEVICE void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ TMA tensor copies (tiled mode)
DASV_DEVICE void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
DASV_DEVICE void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
DASV_DEVICE void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
DASV_DEVICE void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp, ncols power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
DASV_DEVICE void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
DASV_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
DASV_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DASV_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
DASV_DEVICE void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16/fp16 inputs, fp32 accumulate; one thread issues for the CTA.
DASV_DEVICE void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
DASV_DEVICE void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (thread i of the warp reads lane base+i).
DASV_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
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

// ------------------------------------------------------------------ CTA pairs (cta_group::2): two SMs, one 256-row MMA tile
// Both CTAs of a 2-CTA cluster keep identical SMEM layouts; the even-ranked CTA (the leader) issues the MMAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the pair bit of a shared-window address: "the leader's copy"
DASV_DEVICE uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
DASV_DEVICE void cluster_sync_all() {            // every thread of every CTA in the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
DASV_DEVICE void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {   // arrive on the same barrier in another CTA of the cluster
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// TMA loads issued by either CTA of a pair into its OWN shared memory, completing on the LEADER's mbarrier.
DASV_DEVICE void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
DASV_DEVICE void tma_load_4d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
DASV_DEVICE void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {   // one warp in EACH CTA of the pair, same warp id
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
DASV_DEVICE void tmem_relinquish_2sm() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
DASV_DEVICE void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by one thread of the leader.
DASV_DEVICE void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on the same mbarrier in every CTA of `cta_mask` when all previously issued pair MMAs have completed.
DASV_DEVICE void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of 64 bf16 (128 B):
// 8-row groups are 1024 B apart (SBO); LBO unused for a one-atom-wide K extent; version 1 (sm_100);
// layout type 2 = SWIZZLE_128B.  Advancing K by 16 bf16 inside the atom = +32 B on the start address.
DASV_DEVICE uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);          // [0,14)  start address >> 4
    d |= static_cast<uint64_t>(1) << 16;                             // [16,30) LBO (ignored, canonical value 1)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                     // [32,46) SBO = 1024 B
    d |= static_cast<uint64_t>(1) << 46;                             // [46,48) descriptor version = 1
    d |= static_cast<uint64_t>(2) << 61;                             // [61,64) SWIZZLE_128B
    return d;
}
// NB (measured on B200): an operand view may start at ANY 128-byte row of a swizzled tile with this same
// descriptor -- the hardware applies the 128-byte swizzle to absolute SMEM address bits; encoding the row phase in
// the base-offset field [49,52) instead gives wrong results.
// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4)            // [4,6)   D format: f32
         | (1u << 7)            // [7,10)  A format: bf16
         | (1u << 10)           // [10,13) B format: bf16
         | (0u << 15) | (0u << 16)   // A, B K-major
         | ((N >> 3) << 17)     // [17,23) N >> 3
         | ((M >> 4) << 24);    // [24,29) M >> 4
}

// Same, with the operand formats chosen independently: fp16 (1 sign, 5 exponent, 10 mantissa bits) or bf16 (8, 7).
__host__ __device__ constexpr uint32_t umma_idesc_f16kind(uint32_t M, uint32_t N, bool a_f16, bool b_f16) {
    return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the
// stream is still running: everything up to griddep_wait() (barrier init, TMEM allocation, descriptor prefetch)
// overlaps the predecessor's tail; griddep_wait() returns once the predecessor has completed and its writes are
// visible.  griddep_launch() lets the NEXT kernel of the stream begin its own prologue early.  Both are no-ops for a
// kernel launched without the attribute.
DASV_DEVICE void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DASV_DEVICE void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ misc math
DASV_DEVICE float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
DASV_DEVICE float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
DASV_DEVICE uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// fp16 storage saturates at the largest finite value instead of overflowing to infinity
DASV_DEVICE uint32_t pack_f16(float lo, float hi) {
    __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&v);
}
DASV_DEVICE uint16_t cvt_f16_bits(float v) { return __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f))); }
DASV_DEVICE uint16_t cvt_bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
// 16-bit activation formats of the tensor-core path: 1 = bf16, 2 = fp16 (the C ABI's dtype codes)
template <int FMT> DASV_DEVICE uint16_t cvt16_bits(float v) { return FMT == 2 ? cvt_f16_bits(v) : cvt_bf16_bits(v); }
template <int FMT> DASV_DEVICE uint32_t pack16(float lo, float hi) { return FMT == 2 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
// packed fp32x2 arithmetic (sm_100+): one instruction, two independent fp32 FMAs on a 64-bit register pair
DASV_DEVICE uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
DASV_DEVICE void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
DASV_DEVICE uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
DASV_DEVICE float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
DASV_DEVICE void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace dasv
