// mma_rate.cu -- cycles per tcgen05.mma (cta_group::2, M = 256) by kind, operand source and issue pattern, measured on one
// CTA pair with zeroed operands: N back-to-back MMAs into one accumulator, one commit, clock64 around issue..completion.
// Answers what the pair kernel's design rests on: what an fp8 (kind::f8f6f4, K = 32) MMA costs next to an fp16 one
// (K = 16), from shared memory and from tensor memory, and what changing kind between consecutive MMAs costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I twisterl_b200/csrc scripts/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "twr_tc_ptx.cuh"

namespace {
__device__ __forceinline__ void commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int KIND>   // 0 = f16, 1 = f8f6f4
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
constexpr uint32_t idesc(uint32_t n, uint32_t af, uint32_t bf) { return (1u << 4) | (af << 7) | (bf << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24); }

constexpr int REPS = 64;

template <int mode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1) k_rate(long long* out, int outer, int ld_warps, int st_too, int random_data) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t sbase = smem_u32(smem), bar = smem_u32(&bar_mem);
    for (int i = threadIdx.x; i < 8 * 16384 / 16; i += blockDim.x) {
        // zeros, or fp16 values in [1, 2) with random mantissas (as fp8 bytes: finite values of mixed magnitude)
        uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        auto next = [&]() { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return random_data ? ((x & 0x03FF03FFu) | 0x3C003C00u) : 0u; };
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(next(), next(), next(), next());
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (cluster_ctarank() == 0 && threadIdx.x == 32) {
        const uint64_t a0 = make_desc(sbase), b0 = make_desc(sbase + 4 * 16384);
        const uint32_t d = tmem, at = tmem + 256;
        constexpr uint32_t F16_256 = idesc(256, 0, 0), F8_256 = idesc(256, 1, 0), F8_256_55 = idesc(256, 1, 1), F8_256_44 = idesc(256, 0, 0);
        constexpr uint32_t F16_128 = idesc(128, 0, 0), F8_128 = idesc(128, 1, 0);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int o = 0; o < outer; ++o) {
            // descriptors walk over 4 tiles x 4 k-steps like the real kernel (a different 32-byte k-step every MMA)
#pragma unroll
            for (int i = 0; i < REPS; ++i) {
                const uint64_t a = a0 + (uint64_t)((i >> 2) & 3) * 1024u + 2u * (i & 3), b = b0 + (uint64_t)((i >> 2) & 3) * 1024u + 2u * (i & 3);
                const uint32_t ta = at + 8u * (i & 15);
                if constexpr (mode == 0) { mma_ss<0>(d, a, b, F16_256); }
                if constexpr (mode == 1) { mma_ss<1>(d, a, b, F8_256); }
                if constexpr (mode == 2) { mma_ts<0>(d, ta, b, F16_256); }
                if constexpr (mode == 3) { mma_ts<1>(d, ta, b, F8_256_55); }
                if constexpr (mode == 4) { if (i & 1) mma_ss<1>(d, a, b, F8_256); else mma_ss<0>(d, a, b, F16_256); }
                if constexpr (mode == 5) { if ((i >> 2) & 1) mma_ss<1>(d, a, b, F8_256); else mma_ss<0>(d, a, b, F16_256); }
                if constexpr (mode == 6) { if ((i % 12) >= 8) mma_ts<1>(d, ta, b, F8_256_55); else mma_ts<0>(d, ta, b, F16_256); }
                if constexpr (mode == 7) { mma_ss<0>(d, a, b, F16_128); }
                if constexpr (mode == 8) { mma_ss<1>(d, a, b, F8_128); }
                if constexpr (mode == 9) { mma_ss<1>(d, a, b, F8_256_44); }
                if constexpr (mode == 10) { mma_ss<0>(d, a0, b, F16_256); }
                if constexpr (mode == 11) { if ((i >> 4) & 1) mma_ss<1>(d, a, b, F8_256); else mma_ss<0>(d, a, b, F16_256); }
                if constexpr (mode == 12) { if (i & 1) mma_ts<1>(d, ta, b, F8_256_55); else mma_ts<0>(d, ta, b, F16_256); }
                if constexpr (mode == 13) { mma_ts<1>(d, ta, b, F8_256); }
            }
            }
            const long long t1 = clock64();
            commit2(bar);
            mbar_wait(bar, rep & 1);
            const long long t2 = clock64();
            if (rep == 2) { __threadfence(); *((volatile long long*)(out + 7)) = 1; }
            if (blockIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
    } else if (cluster_ctarank() == 1 && threadIdx.x == 32) {
        for (int rep = 0; rep < 3; ++rep) mbar_wait(bar, rep & 1);
    } else if (warp >= 2 && warp < 2 + ld_warps) {
        // epilogue-like traffic on the tensor memory while the MMAs run: 32x32b.x32 loads (and stores) of columns the MMAs
        // do not touch ([256, 512)), by `ld_warps` warps per CTA (warp % 4 = its lane quarter), until the issuer is done
        const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
        volatile long long* done = out + 7;
        uint32_t v[32];
        long long n_ld = 0;
        const long long t0 = clock64();
        while (*done == 0) {
#pragma unroll 1
            for (int i = 0; i < 16; ++i) {
                tc_ld32(tmem + lane_addr + 256u + 32u * (uint32_t)(i & 7), v);
                tc_wait_ld();
                if (st_too) { tc_st32(tmem + lane_addr + 256u + 32u * (uint32_t)(i & 7), v); tc_wait_st(); }
                ++n_ld;
            }
        }
        if (blockIdx.x == 0 && warp == 2 && (threadIdx.x & 31) == 0) { out[8] = n_ld; out[9] = clock64() - t0; }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}
}  // namespace

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 2, outer = argc > 2 ? atoi(argv[2]) : 1;
    const int ld_warps = argc > 3 ? atoi(argv[3]) : 0, st_too = argc > 4 ? atoi(argv[4]) : 0, random_data = argc > 6 ? atoi(argv[6]) : 0;
    printf("operands: %s\n", random_data ? "random fp16 in [1, 2)" : "zeros");
    printf("tensor-memory traffic beside the MMAs: %d warps per CTA looping tcgen05.ld 32x32b.x32%s\n", ld_warps, st_too ? " + tcgen05.st" : "");
    printf("grid %d CTAs, %d x %d MMAs per timed pass\n", grid, outer, REPS);
    const char* names[] = {"f16 SS N256", "f8 SS N256 (e5m2 x e4m3)", "f16 TS N256", "f8 TS N256 (e5m2 x e5m2)", "SS alternate f16/f8 every MMA",
                           "SS alternate in groups of 4", "TS 8 x f16 then 4 x f8, repeated", "f16 SS N128", "f8 SS N128", "f8 SS N256 (e4m3 x e4m3)",
                           "f16 SS N256, same A tile", "SS alternate in groups of 16", "TS alternate f16/f8 every MMA", "f8 TS N256 (e5m2 x e4m3)"};
    long long* d_out;
    cudaMalloc(&d_out, 128);
    const int smem = 8 * 16384;
    void (*kerns[14])(long long*, int, int, int, int) = {k_rate<0>, k_rate<1>, k_rate<2>, k_rate<3>, k_rate<4>, k_rate<5>, k_rate<6>, k_rate<7>, k_rate<8>, k_rate<9>,
                                     k_rate<10>, k_rate<11>, k_rate<12>, k_rate<13>};
    const int n_modes = argc > 5 ? atoi(argv[5]) : 14;
    for (int mode = 0; mode < n_modes; ++mode) {
        cudaMemset(d_out, 0, 128);
        cudaFuncSetAttribute(kerns[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kerns[mode]<<<grid, 320, smem>>>(d_out, outer, ld_warps, st_too, random_data);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long h[10];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-36s issue %6.1f  complete %6.1f cycles per MMA (%d MMAs; first pass %0.1f)\n", names[mode], (double)h[4] / (REPS * outer), (double)h[5] / (REPS * outer), REPS * outer,
               (double)h[1] / (REPS * outer));
        if (ld_warps) printf("%-36s   tcgen05.ld: %.1f cycles per x32 load per warp\n", "", (double)h[9] / (double)(h[8] ? h[8] : 1));
    }
    return 0;
}
