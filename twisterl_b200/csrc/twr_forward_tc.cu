// twr_forward_tc.cu -- K2, tensor-core variant for sm_100a: batched Policy::_raw_predict
// (rust/src/nn/policy.rs:79-100) as two chained tcgen05 GEMMs per tile of 128 envs.
//
//   GEMM1 (embedding, layers.rs:56-86):  D1[128 x E]  = onehot(obs)[128 x K1] * Emb[K1 x E]
//   epilogue1:                           h1 = relu(D1 + emb_bias) -> fp16 hi + lo
//   GEMM2 (common Linear, layers.rs:31-37): D2[128 x H] = h1[128 x E] * W1[E x H]
//   epilogue2 (heads, policy.rs:89-97):  h2 = relu(D2 + b1); logits = Wa h2 + ba; value = wv.h2 + bv
//
// Precision: every fp32 operand is split into two fp16 terms (x = hi + lo, |lo| <= 2^-11 |x|) and the
// products hi*hi, hi*lo, lo*hi are accumulated in fp32 in TMEM -- 2 MMAs for GEMM1 (the one-hot A is
// exact), 3 for GEMM2.  On the shipped ppo_puzzle15_v1 weights this is within 1e-5 of fp32 (bar 1e-3).
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer: streams 16 KB operand tiles (pre-swizzled in HBM/L2 by k_tc_pack, so a
//               tile is one contiguous cp.async.bulk) into a 5-slot shared-memory ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=128, K=16, kind::f16);
//               accumulators live in TMEM: D2 in columns [0,H), D1 double-buffered in [256,512)
//   warps 2..5  epilogue: thread == env row.  Build the one-hot A1 tile (K-major, 128B swizzle) from
//               the 16-byte env state, tcgen05.ld D1 -> bias/ReLU/split -> st.shared A2 (swizzled),
//               tcgen05.ld D2 -> heads on CUDA cores -> logits/values
//   All hand-offs are mbarriers; tcgen05.commit releases ring slots / publishes accumulators.
//   E is processed in chunks of 128 columns: G1(c) fills D1[c&1], epilogue1(c) turns it into the
//   A2 k-chunk, G2(c) accumulates it into D2, while G1(c+1) already runs -- and epilogue2 of tile i
//   overlaps G1 of tile i+1.
//
// Algorithmic flop per env-step: 2*E*H + n_obs*E + 2*H*5 (272 896 for puzzle15); executed tensor
// flop: 2*K1*E*2 + 2*E*H*3.  Bound: tensor pipe / L2->SMEM operand streaming (1 MB per tile).
#include "twr_kernels.cuh"

#include <atomic>
#include <cuda_fp16.h>

extern std::atomic<long long> g_twr_launches;

namespace {

constexpr int TM = 128;               // envs per tile (UMMA M)
constexpr int NTHREADS = 192;
constexpr int TILE_BYTES = 16384;     // [128 rows x 64 k] fp16, K-major, 128B swizzle
constexpr int NSLOTS = 5;
constexpr int MAX_KB1 = 4;            // obs_size <= 256
constexpr int MAX_OBS = 32;

// ---- shared memory map (dynamic, 1024-byte aligned base) ----
constexpr int SM_A1 = 0;                                  // MAX_KB1 tiles
constexpr int SM_A2H = SM_A1 + MAX_KB1 * TILE_BYTES;      // 2 tiles (128 k)
constexpr int SM_A2L = SM_A2H + 2 * TILE_BYTES;           // 2 tiles
constexpr int SM_RING = SM_A2L + 2 * TILE_BYTES;          // NSLOTS tiles
constexpr int SM_MISC = SM_RING + NSLOTS * TILE_BYTES;    // 212992
constexpr int SM_HEADW = SM_MISC;                         // [256][8] float = 8192
constexpr int SM_B1 = SM_HEADW + 8192;                    // [256] float
constexpr int SM_EMBB = SM_B1 + 1024;                     // [1024] float (E <= 1024)
constexpr int SM_ROWS = SM_EMBB + 4096;                   // [128][MAX_OBS] u8 = 4096
constexpr int SM_BARS = SM_ROWS + TM * MAX_OBS;           // mbarriers
constexpr int SM_TOTAL = SM_BARS + 256;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");

enum Bar { B_FULL0 = 0, B_EMPTY0 = NSLOTS, B_A1_FULL = 2 * NSLOTS, B_A1_EMPTY, B_D1_FULL0, B_D1_FULL1, B_D1_EMPTY0,
           B_D1_EMPTY1, B_A2_FULL, B_A2_EMPTY, B_D2_FULL, B_D2_EMPTY, B_COUNT };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");

// ---- PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin with a watchdog: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {   // arrives on `bar` when all prior MMAs of this thread finish
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; M=128, N=128, K=16, fp16 in, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base+i), columns [col, col+32)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (=1024 B between
// 8-row groups) | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=b=F16 (0), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t IDESC_128x128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// byte offset of element (row, k) inside one [128 x 64] fp16 tile, K-major with 128B swizzle
__host__ __device__ __forceinline__ uint32_t tile_off(uint32_t row, uint32_t k) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 3) ^ row) & 7u) << 4) + (k & 7u) * 2u;
}

struct TcParams {
    int NC;      // E / 128 chunks
    int NKB1;    // ceil(obs_size / 64) k-blocks of GEMM1
    int NH;      // H / 128 column halves of GEMM2
    int E, H;
};

__host__ __device__ inline size_t g1_tiles(const TcParams& t) { return (size_t)t.NC * t.NKB1 * 2; }
__host__ __device__ inline size_t g2_tiles(const TcParams& t) { return (size_t)t.NC * 2 * 2 * t.NH; }

// ---------------------------------------------------------------------- pack ---
// Operand image in HBM == the shared-memory image, in the exact order the producer streams it:
//   G1 region: for c, kb: [hi tile][lo tile]       tile rows n = feature c*128+r, k = obs row kb*64+kk
//   G2 region: for j, kb, part(hi,lo), half:       tile rows n = output half*128+r, k = feature j*128+kb*64+kk
__global__ void __launch_bounds__(256) k_tc_pack(PolicyDev p, TcParams t, __half* __restrict__ pack) {
    const size_t n1 = g1_tiles(t), n2 = g2_tiles(t);
    const size_t total = (n1 + n2) * (TILE_BYTES / 2);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t tile = idx / (TILE_BYTES / 2);
        const uint32_t within = (uint32_t)(idx % (TILE_BYTES / 2));
        const uint32_t r = within >> 6, kk = within & 63u;
        float x = 0.0f;
        int lo_part;
        if (tile < n1) {
            const int c = (int)(tile / (2 * t.NKB1)), rem = (int)(tile % (2 * t.NKB1));
            const int kb = rem >> 1;
            lo_part = rem & 1;
            const int f = c * 128 + (int)r, k = kb * 64 + (int)kk;
            if (k < p.obs_size) x = p.emb[(size_t)k * p.E + f];
        } else {
            size_t q = tile - n1;
            const int per_kb = 2 * t.NH;
            const int j = (int)(q / (2 * per_kb)), rem = (int)(q % (2 * per_kb));
            const int kb = rem / per_kb, rem2 = rem % per_kb;
            lo_part = rem2 / t.NH;
            const int half = rem2 % t.NH;
            const int o = half * 128 + (int)r, f = j * 128 + kb * 64 + (int)kk;
            x = p.w1[(size_t)f * p.H + o];
        }
        const __half hi = __float2half_rn(x);
        const __half v = lo_part ? __float2half_rn(x - __half2float(hi)) : hi;
        *reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(pack) + tile * TILE_BYTES + tile_off(r, kk)) = v;
    }
}

// -------------------------------------------------------------------- kernel ---
template <int NH>   // H = 128 * NH
__global__ void __launch_bounds__(NTHREADS, 1) k_forward_tc(PolicyDev p, ForwardArgs a, TcParams t) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = a.n_live_ptr ? (int64_t)*a.n_live_ptr : a.n;
    const int n_tiles = (int)((n + TM - 1) / TM);
    if ((int)blockIdx.x >= n_tiles) return;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + SM_BARS;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BARS + 8 * B_COUNT);
    float* headw = reinterpret_cast<float*>(smem + SM_HEADW);
    float* b1s = reinterpret_cast<float*>(smem + SM_B1);
    float* embb = reinterpret_cast<float*>(smem + SM_EMBB);
    uint8_t* rows_s = smem + SM_ROWS;
    const int NC = t.NC, NKB1 = t.NKB1, H = t.H;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOTS; ++i) { mbar_init(bar(B_FULL0 + i), 1); mbar_init(bar(B_EMPTY0 + i), 1); }
        mbar_init(bar(B_A1_FULL), 128); mbar_init(bar(B_A1_EMPTY), 1);
        mbar_init(bar(B_D1_FULL0), 1); mbar_init(bar(B_D1_FULL1), 1);
        mbar_init(bar(B_D1_EMPTY0), 128); mbar_init(bar(B_D1_EMPTY1), 128);
        mbar_init(bar(B_A2_FULL), 128); mbar_init(bar(B_A2_EMPTY), 1);
        mbar_init(bar(B_D2_FULL), 1); mbar_init(bar(B_D2_EMPTY), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the one-hot tile, stage head weights and biases
    for (int i = threadIdx.x; i < MAX_KB1 * TILE_BYTES / 16; i += NTHREADS)
        reinterpret_cast<uint4*>(smem + SM_A1)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < H; i += NTHREADS) {
        float* w = headw + i * 8;
        for (int o = 0; o < 4; ++o) w[o] = o < p.A ? p.wa[(size_t)i * p.A + o] : 0.0f;
        w[4] = p.wv[i]; w[5] = 0.f; w[6] = 0.f; w[7] = 0.f;
        b1s[i] = p.b1[i];
    }
    for (int i = threadIdx.x; i < t.E; i += NTHREADS) embb[i] = p.emb_b[i];
    for (int i = threadIdx.x; i < TM * MAX_OBS; i += NTHREADS) rows_s[i] = 0xFF;
    if (warp == 1) {   // TMEM: all 512 columns (1 CTA per SM by construction: > 113 KB shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t D2_COL = 0, D1_COL = 256;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            const unsigned char* g1 = reinterpret_cast<const unsigned char*>(p.tc_pack);
            const unsigned char* g2 = g1 + g1_tiles(t) * TILE_BYTES;
            uint32_t use = 0;
            auto push = [&](const unsigned char* src) {
                const uint32_t slot = use % NSLOTS, round = use / NSLOTS;
                mbar_wait(bar(B_EMPTY0 + slot), (round & 1u) ^ 1u);
                mbar_expect_tx(bar(B_FULL0 + slot), TILE_BYTES);
                bulk_g2s(sbase + SM_RING + slot * TILE_BYTES, src, TILE_BYTES, bar(B_FULL0 + slot));
                ++use;
            };
            auto push_g1 = [&](int c) { for (int i = 0; i < 2 * NKB1; ++i) push(g1 + ((size_t)c * 2 * NKB1 + i) * TILE_BYTES); };
            auto push_g2 = [&](int j) { for (int i = 0; i < 4 * NH; ++i) push(g2 + ((size_t)j * 4 * NH + i) * TILE_BYTES); };
            for (int it = 0; it < my_tiles; ++it) {
                push_g1(0);
                for (int c = 1; c < NC; ++c) { push_g1(c); push_g2(c - 1); }
                push_g2(NC - 1);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer =================================
        if (lane == 0) {
            uint32_t use = 0, d1use = 0, a2use = 0;
            auto wait_slot = [&]() -> uint32_t {
                const uint32_t slot = use % NSLOTS, round = use / NSLOTS;
                mbar_wait(bar(B_FULL0 + slot), round & 1u);
                tc_fence_after();
                ++use;
                return slot;
            };
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(bar(B_A1_FULL), it & 1);
                tc_fence_after();
                auto g1 = [&](int c) {
                    const uint32_t buf = d1use & 1u;
                    mbar_wait(bar(B_D1_EMPTY0 + buf), ((d1use >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d = tmem + D1_COL + buf * 128u;
                    for (int kb = 0; kb < NKB1; ++kb) {
                        for (int part = 0; part < 2; ++part) {
                            const uint32_t slot = wait_slot();
                            const uint64_t ad = make_desc(sbase + SM_A1 + kb * TILE_BYTES);
                            const uint64_t bd = make_desc(sbase + SM_RING + slot * TILE_BYTES);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc_mma(d, ad + 2u * ks, bd + 2u * ks, IDESC_128x128, (kb | part | ks) != 0);
                            tc_commit(bar(B_EMPTY0 + slot));
                        }
                    }
                    tc_commit(bar(B_D1_FULL0 + buf));
                    if (c == NC - 1) tc_commit(bar(B_A1_EMPTY));
                    ++d1use;
                };
                auto g2 = [&](int j) {
                    mbar_wait(bar(B_A2_FULL), a2use & 1u);
                    tc_fence_after();
                    if (j == 0) { mbar_wait(bar(B_D2_EMPTY), (it & 1) ^ 1); tc_fence_after(); }
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t ah = make_desc(sbase + SM_A2H + kb * TILE_BYTES);
                        const uint64_t al = make_desc(sbase + SM_A2L + kb * TILE_BYTES);
                        for (int part = 0; part < 2; ++part) {          // 0: W hi tiles, 1: W lo tiles
                            for (int half = 0; half < NH; ++half) {
                                const uint32_t slot = wait_slot();
                                const uint64_t bd = make_desc(sbase + SM_RING + slot * TILE_BYTES);
                                const uint32_t d = tmem + D2_COL + half * 128u;
                                const bool first = (j == 0 && kb == 0 && part == 0);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    tc_mma(d, ah + 2u * ks, bd + 2u * ks, IDESC_128x128, !(first && ks == 0));
                                if (part == 0) {
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks) tc_mma(d, al + 2u * ks, bd + 2u * ks, IDESC_128x128, 1u);
                                }
                                tc_commit(bar(B_EMPTY0 + slot));
                            }
                        }
                    }
                    tc_commit(bar(B_A2_EMPTY));
                    ++a2use;
                    if (j == NC - 1) tc_commit(bar(B_D2_FULL));
                };
                g1(0);
                for (int c = 1; c < NC; ++c) { g1(c); g2(c - 1); }
                g2(NC - 1);
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue warps =============================
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;          // env row inside the tile
        const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
        const int n_obs = p.n_obs;
        uint32_t d1use = 0, a2use = 0;
        int perm_cur = -1;

        // one-hot A1 tile for tile `it`: clear the previous ones of this row, set the new ones
        auto build_a1 = [&](int it) -> int {
            mbar_wait(bar(B_A1_EMPTY), (it & 1) ^ 1);
            const int64_t pos = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TM + row;
            int perm = -1;
            EnvState s; s.lo = 0; s.hi = 0; s.blank = 0; s.depth = 0;
            int64_t e = 0;
            if (pos < n) {
                e = a.live ? a.live[pos] : pos;
                if (!a.obs_rows) {
                    const uint4 c = a.cells[e];
                    s.lo = (uint64_t)c.x | ((uint64_t)c.y << 32);
                    s.hi = (uint64_t)c.z | ((uint64_t)c.w << 32);
                }
                if (a.perm_idx) {
                    perm = a.perm_idx[pos];
                } else if (p.n_perms > 0 && a.t >= 0) {        // get_perm_id, nn/policy.rs:67-77
                    uint32_t w[4];
                    philox4x32_10(a.env_id_base + (uint32_t)e, (uint32_t)a.t, TWR_RNG_PERM, a.cid, (uint32_t)a.seed,
                                  (uint32_t)(a.seed >> 32), w);
                    perm = (int)mulhi_u32(w[0], (uint32_t)p.n_perms);
                }
            }
            for (int i = 0; i < n_obs; ++i) {
                const uint32_t old = rows_s[row * MAX_OBS + i];
                if (old != 0xFF || it > 0)
                    *reinterpret_cast<__half*>(smem + SM_A1 + (old >> 6) * TILE_BYTES + tile_off(row, old & 63u)) = __ushort_as_half(0);
            }
            for (int i = 0; i < n_obs; ++i) {
                int r = 0;
                if (pos < n) {
                    r = a.obs_rows ? a.obs_rows[pos * n_obs + i] : i * a.env.N + (int)env_board(a.env, s, i);
                    if (perm >= 0) r = p.obs_perms[(size_t)perm * p.obs_size + r];   // twist-in, policy.rs:81-83
                }
                rows_s[row * MAX_OBS + i] = (uint8_t)r;
                *reinterpret_cast<__half*>(smem + SM_A1 + (r >> 6) * TILE_BYTES + tile_off(row, (uint32_t)r & 63u)) = __ushort_as_half(0x3C00);
            }
            fence_async_smem();
            mbar_arrive(bar(B_A1_FULL));
            return perm;
        };

        int perm_next = build_a1(0);
        for (int it = 0; it < my_tiles; ++it) {
            perm_cur = perm_next;
            // ---- epilogue 1: D1 chunk -> relu(x + bias) -> fp16 hi/lo -> A2 (swizzled K-major)
            for (int c = 0; c < NC; ++c) {
                const uint32_t buf = d1use & 1u;
                mbar_wait(bar(B_D1_FULL0 + buf), (d1use >> 1) & 1u);
                tc_fence_after();
                for (int q = 0; q < 4; ++q) {                       // 4 x 32 columns = 2 k-blocks of 64
                    uint32_t v[32];
                    tc_ld32(tmem + lane_addr + D1_COL + buf * 128u + q * 32u, v);
                    tc_wait_ld();
                    if (q == 0) mbar_wait(bar(B_A2_EMPTY), (a2use & 1u) ^ 1u);   // GEMM2 of the previous chunk done with A2
                    const float* bias = embb + c * 128 + q * 32;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {                   // 8 columns -> one 16-byte chunk per hi / lo
                        uint32_t hp[4], lp[4];
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            float x0 = __uint_as_float(v[g * 8 + e2 * 2]) + bias[g * 8 + e2 * 2];
                            float x1 = __uint_as_float(v[g * 8 + e2 * 2 + 1]) + bias[g * 8 + e2 * 2 + 1];
                            x0 = x0 > 0.f ? x0 : 0.f; x1 = x1 > 0.f ? x1 : 0.f;
                            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                            const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
                            hp[e2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                            lp[e2] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                        }
                        const uint32_t kcol = q * 32 + g * 8;       // k inside the 128-wide chunk
                        const uint32_t off = (kcol >> 6) * TILE_BYTES + tile_off(row, kcol & 63u);
                        *reinterpret_cast<uint4*>(smem + SM_A2H + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        *reinterpret_cast<uint4*>(smem + SM_A2L + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                    }
                }
                tc_fence_before();
                mbar_arrive(bar(B_D1_EMPTY0 + buf));
                fence_async_smem();
                mbar_arrive(bar(B_A2_FULL));
                ++d1use; ++a2use;
            }
            // ---- next tile's one-hot operand, so its GEMM1 overlaps this tile's heads
            if (it + 1 < my_tiles) perm_next = build_a1(it + 1);

            // ---- epilogue 2: heads on CUDA cores from the fp32 accumulator row
            mbar_wait(bar(B_D2_FULL), it & 1);
            tc_fence_after();
            float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            for (int q = 0; q < H / 32; ++q) {
                uint32_t v[32];
                tc_ld32(tmem + lane_addr + D2_COL + q * 32u, v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = q * 32 + j;
                    float h = __uint_as_float(v[j]) + b1s[col];
                    h = h > 0.f ? h : 0.f;
                    const float4 w0 = *reinterpret_cast<const float4*>(headw + col * 8);
                    const float w4 = headw[col * 8 + 4];
                    acc[0] = fmaf(h, w0.x, acc[0]); acc[1] = fmaf(h, w0.y, acc[1]);
                    acc[2] = fmaf(h, w0.z, acc[2]); acc[3] = fmaf(h, w0.w, acc[3]);
                    acc[4] = fmaf(h, w4, acc[4]);
                }
            }
            tc_fence_before();
            mbar_arrive(bar(B_D2_EMPTY));
            const int64_t pos = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TM + row;
            if (pos < n) {
                float l[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) l[o] = o < p.A ? acc[o] + p.ba[o] : 0.0f;
                float out[4] = {l[0], l[1], l[2], l[3]};
                if (perm_cur >= 0) {                                  // twist-out, nn/policy.rs:95-97
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (o < p.A) {
                            const int src = p.act_perms[perm_cur * p.A + o];
                            out[o] = src == 0 ? l[0] : src == 1 ? l[1] : src == 2 ? l[2] : l[3];
                        }
                    }
                }
                a.logits[pos] = make_float4(out[0], out[1], out[2], out[3]);
                a.values[pos] = acc[4] + p.bv[0];
            }
        }
    }

    // ---- teardown: everyone done with TMEM, then the allocating warp frees it
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

TcParams make_params(const PolicyDev& p) {
    TcParams t;
    t.E = p.E; t.H = p.H;
    t.NC = p.E / 128;
    t.NKB1 = (p.obs_size + 63) / 64;
    t.NH = p.H / 128;
    return t;
}

int g_num_sms = 0;

}  // namespace

int forward_tc_supported(const PolicyDev& p, const EnvParams&, const char** why) {
    static const char* m1 = "tensor-core forward needs obs_size <= 256 (one-hot operand must fit 64 KB of shared memory)";
    static const char* m2 = "tensor-core forward needs embedding size a multiple of 128, <= 1024";
    static const char* m3 = "tensor-core forward needs common width 128 or 256";
    static const char* m4 = "tensor-core forward needs <= 32 observation indices per env";
    if (p.obs_size > 64 * MAX_KB1) { *why = m1; return 0; }
    if (p.E % 128 || p.E > 1024 || p.E < 128) { *why = m2; return 0; }
    if (p.H != 128 && p.H != 256) { *why = m3; return 0; }
    if (p.n_obs > MAX_OBS) { *why = m4; return 0; }
    return 1;
}

size_t forward_tc_pack_bytes(const PolicyDev& p) {
    const TcParams t = make_params(p);
    return (g1_tiles(t) + g2_tiles(t)) * TILE_BYTES;
}

void launch_forward_tc_pack(cudaStream_t st, const PolicyDev& p, void* pack) {
    const TcParams t = make_params(p);
    k_tc_pack<<<1024, 256, 0, st>>>(p, t, reinterpret_cast<__half*>(pack));
    g_twr_launches.fetch_add(1, std::memory_order_relaxed);
}

void launch_forward_tc(cudaStream_t st, const PolicyDev& p, const ForwardArgs& a) {
    if (a.n <= 0) return;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const TcParams t = make_params(p);
    const int n_tiles = (int)((a.n + TM - 1) / TM);
    const int grid = n_tiles < g_num_sms ? n_tiles : g_num_sms;
    if (t.NH == 2) {
        cudaFuncSetAttribute(k_forward_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        k_forward_tc<2><<<grid, NTHREADS, SM_TOTAL, st>>>(p, a, t);
    } else {
        cudaFuncSetAttribute(k_forward_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        k_forward_tc<1><<<grid, NTHREADS, SM_TOTAL, st>>>(p, a, t);
    }
    g_twr_launches.fetch_add(1, std::memory_order_relaxed);
}
