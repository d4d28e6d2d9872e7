// micro-benchmark: legacy mma.sync (HMMA.16816 bf16) and ldmatrix latency / throughput on sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int ILP>
__global__ void k_mma(float* out, int iters, long long* cyc) {
    float c[ILP][4];
    uint32_t a[4] = {threadIdx.x, 2, 3, 4};
    for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) c[i][e] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma(c[i], a, 5u + i, 6u);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) s += c[i][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, bool TRANS>
__global__ void k_ldsm(float* out, int iters, long long* cyc) {
    __shared__ __align__(128) unsigned char sm[64 * 1024 > 48 * 1024 ? 40 * 1024 : 0];
    for (int i = threadIdx.x; i < 10 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    __syncthreads();
    uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) + (threadIdx.x & 15) * 2064 + (threadIdx.x >> 4 & 1) * 16;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            uint32_t r[4];
            if (TRANS) asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base + i * 32 + (acc & 0)));
            else asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base + i * 32 + (acc & 0)));
            acc += r[0] ^ r[1] ^ r[2] ^ r[3];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    int warps[] = {1, 4, 8, 16, 32};
    for (int w : warps) {
        k_mma<1><<<148, w * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        printf("mma  ILP1 warps/SM %2d: %.1f cyc per mma per warp, %.2f cyc per mma per SM\n", w, (double)*cyc / iters, (double)*cyc / iters / w);
        k_mma<8><<<148, w * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        printf("mma  ILP8 warps/SM %2d: %.1f cyc per mma per warp, %.2f cyc per mma per SM\n", w, (double)*cyc / iters / 8, (double)*cyc / iters / 8 / w);
        k_ldsm<4, false><<<148, w * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        printf("ldsm ILP4 warps/SM %2d: %.1f cyc per ldsm.x4 per warp, %.2f per SM\n", w, (double)*cyc / iters / 4, (double)*cyc / iters / 4 / w);
        k_ldsm<4, true><<<148, w * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        printf("ldsmT ILP4 warps/SM %2d: %.1f cyc per ldsm.x4.trans per warp, %.2f per SM\n", w, (double)*cyc / iters / 4, (double)*cyc / iters / 4 / w);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
