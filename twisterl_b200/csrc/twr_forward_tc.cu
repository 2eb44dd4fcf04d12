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
//               tile is one contiguous cp.async.bulk) into an 8-slot shared-memory ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=128, K=16, kind::f16);
//               accumulators live in TMEM: D2 in columns [0,H), D1 double-buffered in [256,512)
//   warps 2..5  epilogue: thread == env row.  Build the one-hot A1 tile (K-major, 128B swizzle) from
//               the 16-byte env state, tcgen05.ld D1 -> bias/ReLU/fp16 split -> tcgen05.st back into the
//               same TMEM columns (GEMM2 reads its A operand from TMEM), tcgen05.ld D2 -> heads on CUDA
//               cores -> logits/values
//   All hand-offs are mbarriers; tcgen05.commit releases ring slots / publishes accumulators.
//   E is processed in chunks of 128 columns: G1(c) fills D1[c&1], epilogue1(c) turns it into the
//   A2 k-chunk, G2(c) accumulates it into D2, while G1(c+1) already runs -- and epilogue2 of tile i
//   overlaps G1 of tile i+1.
//
// Algorithmic flop per env-step: 2*E*H + n_obs*E + 2*H*5 (272 896 for puzzle15); executed tensor
// flop: 2*K1*E*2 + 2*E*H*3.  Bound: tensor pipe / L2->SMEM operand streaming (1 MB per tile).
#include "twr_kernels.cuh"
#include "twr_tc_ptx.cuh"

#include <atomic>
#include <cstdlib>
#include <cuda_fp16.h>

extern std::atomic<long long> g_twr_launches;

namespace {

constexpr int TM = 128;               // envs per tile (UMMA M)
constexpr int NTHREADS = 192;
constexpr int TILE_BYTES = 16384;     // [128 rows x 64 k] fp16, K-major, 128B swizzle
constexpr int NSLOTS = 8;
constexpr int MAX_KB1 = 4;            // obs_size <= 256
constexpr int MAX_OBS = 32;

// ---- shared memory map (dynamic, 1024-byte aligned base) ----
constexpr int SM_A1 = 0;                                  // MAX_KB1 tiles
constexpr int SM_RING = SM_A1 + MAX_KB1 * TILE_BYTES;     // NSLOTS tiles
constexpr int SM_MISC = SM_RING + NSLOTS * TILE_BYTES;    // 196608
constexpr int SM_HEADW = SM_MISC;                         // [256][8] float = 8192
constexpr int SM_B1 = SM_HEADW + 8192;                    // [256] float
constexpr int SM_EMBB = SM_B1 + 1024;                     // [1024] float (E <= 1024)
constexpr int SM_ROWS = SM_EMBB + 4096;                   // [128][MAX_OBS] u8 = 4096
constexpr int SM_BARS = SM_ROWS + TM * MAX_OBS;           // mbarriers
constexpr int SM_TOTAL = SM_BARS + 256;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");

enum Bar { B_FULL0 = 0, B_EMPTY0 = NSLOTS, B_A1_FULL = 2 * NSLOTS, B_A1_EMPTY, B_D1_FULL0, B_D1_FULL1, B_A2_FULL0,
           B_A2_FULL1, B_D2_FULL, B_D2_EMPTY, B_COUNT };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");

// PTX wrappers, descriptors and tile_off(): twr_tc_ptx.cuh


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
// CSZ = thread-block cluster size: the CSZ CTAs of a cluster walk the same operand stream, each CTA
// fetches every CSZ-th tile and TMA-multicasts it into the ring of all of them (L2 -> SM traffic / CSZ).
template <int NH, int CSZ>   // H = 128 * NH
__global__ void __launch_bounds__(NTHREADS, 1) k_forward_tc(PolicyDev p, ForwardArgs a, TcParams t) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = a.n_live_ptr ? (int64_t)*a.n_live_ptr : a.n;
    const int n_tiles = (int)((n + TM - 1) / TM);
    const int crank = CSZ > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = (int)blockIdx.x / CSZ, n_clusters = (int)gridDim.x / CSZ;
    const int n_groups = (n_tiles + CSZ - 1) / CSZ;            // a group = CSZ consecutive tiles, one per CTA
    if (cluster_id >= n_groups) return;                        // cluster-uniform
    const int my_tiles = (n_groups - cluster_id + n_clusters - 1) / n_clusters;
    constexpr uint16_t CMASK = (uint16_t)((1u << CSZ) - 1u);
    auto tile_of = [&](int it) -> int64_t { return ((int64_t)cluster_id + (int64_t)it * n_clusters) * CSZ + crank; };

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
        for (int i = 0; i < NSLOTS; ++i) { mbar_init(bar(B_FULL0 + i), 1); mbar_init(bar(B_EMPTY0 + i), CSZ); }
        mbar_init(bar(B_A1_FULL), 128); mbar_init(bar(B_A1_EMPTY), 1);
        mbar_init(bar(B_D1_FULL0), 1); mbar_init(bar(B_D1_FULL1), 1);
        mbar_init(bar(B_A2_FULL0), 128); mbar_init(bar(B_A2_FULL1), 128);
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
    if (CSZ > 1) cluster_sync_all();       // peers' barriers are initialised before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t D2_COL = 0, D1_COL = 256;
    const bool timed = a.dbg != nullptr;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            const unsigned char* g1 = reinterpret_cast<const unsigned char*>(p.tc_pack);
            const unsigned char* g2 = g1 + g1_tiles(t) * TILE_BYTES;
            uint32_t use = 0;
            long long w_empty = 0;
            auto push = [&](const unsigned char* src) {
                const uint32_t slot = use % NSLOTS, round = use / NSLOTS;
                mbar_wait_t(bar(B_EMPTY0 + slot), (round & 1u) ^ 1u, w_empty, timed);
                mbar_expect_tx(bar(B_FULL0 + slot), TILE_BYTES);
                if (CSZ == 1) bulk_g2s(sbase + SM_RING + slot * TILE_BYTES, src, TILE_BYTES, bar(B_FULL0 + slot));
                else if ((int)(use % CSZ) == crank)
                    bulk_g2s_mc(sbase + SM_RING + slot * TILE_BYTES, src, TILE_BYTES, bar(B_FULL0 + slot), CMASK);
                ++use;
            };
            auto push_g1 = [&](int c) { for (int i = 0; i < 2 * NKB1; ++i) push(g1 + ((size_t)c * 2 * NKB1 + i) * TILE_BYTES); };
            auto push_g2 = [&](int j) { for (int i = 0; i < 4 * NH; ++i) push(g2 + ((size_t)j * 4 * NH + i) * TILE_BYTES); };
            for (int it = 0; it < my_tiles; ++it) {
                push_g1(0);
                for (int c = 1; c < NC; ++c) { push_g1(c); push_g2(c - 1); }
                push_g2(NC - 1);
            }
            if (a.dbg) a.dbg[blockIdx.x * 16 + 8] = w_empty;
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer =================================
        if (lane == 0) {
            uint32_t use = 0, d1use = 0, a2use = 0;
            long long w_slot = 0, w_a1 = 0, w_a2 = 0, w_d2 = 0;
            const long long t_begin = clock64();
            auto wait_slot = [&]() -> uint32_t {
                const uint32_t slot = use % NSLOTS, round = use / NSLOTS;
                mbar_wait_t(bar(B_FULL0 + slot), round & 1u, w_slot, timed);
                tc_fence_after();
                ++use;
                return slot;
            };
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait_t(bar(B_A1_FULL), it & 1, w_a1, timed);
                tc_fence_after();
                auto g1 = [&](int c) {
                    // D1 buffer (c & 1) is free: its previous contents (chunk c-2, rewritten in place as the
                    // fp16 A operand) were consumed by G2(c-2), issued earlier on this same in-order pipe
                    const uint32_t buf = d1use & 1u;
                    const uint32_t d = tmem + D1_COL + buf * 128u;
                    for (int kb = 0; kb < NKB1; ++kb) {
                        for (int part = 0; part < 2; ++part) {
                            const uint32_t slot = wait_slot();
                            const uint64_t ad = make_desc(sbase + SM_A1 + kb * TILE_BYTES);
                            const uint64_t bd = make_desc(sbase + SM_RING + slot * TILE_BYTES);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc_mma(d, ad + 2u * ks, bd + 2u * ks, IDESC_128x128, (kb | part | ks) != 0);
                            if (CSZ == 1) tc_commit(bar(B_EMPTY0 + slot)); else tc_commit_mc(bar(B_EMPTY0 + slot), CMASK);
                        }
                    }
                    tc_commit(bar(B_D1_FULL0 + buf));
                    if (c == NC - 1) tc_commit(bar(B_A1_EMPTY));
                    ++d1use;
                };
                auto g2 = [&](int j) {
                    const uint32_t buf = a2use & 1u;
                    mbar_wait_t(bar(B_A2_FULL0 + buf), (a2use >> 1) & 1u, w_a2, timed);
                    tc_fence_after();
                    if (j == 0) { mbar_wait_t(bar(B_D2_EMPTY), (it & 1) ^ 1, w_d2, timed); tc_fence_after(); }
                    const uint32_t a_base = tmem + D1_COL + buf * 128u;   // hi at +32q+8r, lo 16 columns further
                    for (int kb = 0; kb < 2; ++kb) {
                        for (int part = 0; part < 2; ++part) {          // 0: W hi tiles, 1: W lo tiles
                            for (int half = 0; half < NH; ++half) {
                                const uint32_t slot = wait_slot();
                                const uint64_t bd = make_desc(sbase + SM_RING + slot * TILE_BYTES);
                                const uint32_t d = tmem + D2_COL + half * 128u;
                                const bool first = (j == 0 && kb == 0 && part == 0);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint32_t sidx = kb * 4 + ks;   // K=16 step inside the 128-wide chunk
                                    const uint32_t ah = a_base + 32u * (sidx >> 1) + 8u * (sidx & 1u);
                                    tc_mma_ts(d, ah, bd + 2u * ks, IDESC_128x128, !(first && ks == 0));
                                    if (part == 0) tc_mma_ts(d, ah + 16u, bd + 2u * ks, IDESC_128x128, 1u);
                                }
                                if (CSZ == 1) tc_commit(bar(B_EMPTY0 + slot)); else tc_commit_mc(bar(B_EMPTY0 + slot), CMASK);
                            }
                        }
                    }
                    ++a2use;
                    if (j == NC - 1) tc_commit(bar(B_D2_FULL));
                };
                g1(0);
                for (int c = 1; c < NC; ++c) { g1(c); g2(c - 1); }
                g2(NC - 1);
            }
            if (a.dbg) {
                long long* d = a.dbg + blockIdx.x * 16;
                d[0] = clock64() - t_begin; d[1] = w_slot; d[2] = w_a1; d[3] = w_a2; d[4] = w_d2; d[5] = my_tiles;
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue warps =============================
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;          // env row inside the tile
        const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
        const int n_obs = p.n_obs;
        uint32_t d1use = 0;
        int perm_cur = -1;
        long long w_d1 = 0, w_d2f = 0, w_a1e = 0;
        const long long t_begin = clock64();

        // one-hot A1 tile for tile `it`: clear the previous ones of this row, set the new ones
        auto build_a1 = [&](int it) -> int {
            mbar_wait_t(bar(B_A1_EMPTY), (it & 1) ^ 1, w_a1e, timed);
            const int64_t pos = tile_of(it) * TM + row;
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
                    philox4x32_10(a.ids.gid((uint32_t)e), (uint32_t)a.t, TWR_RNG_PERM, a.cid, (uint32_t)a.seed,
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
            // ---- epilogue 1: D1 chunk -> relu(x + bias) -> fp16 hi/lo, written back IN PLACE into the same
            //      TMEM columns as the A operand of GEMM2: batch q (32 fp32 columns) becomes 16 columns of
            //      packed hi pairs followed by 16 columns of packed lo pairs
            for (int c = 0; c < NC; ++c) {
                const uint32_t buf = d1use & 1u;
                mbar_wait_t(bar(B_D1_FULL0 + buf), (d1use >> 1) & 1u, w_d1, timed);
                tc_fence_after();
                for (int q = 0; q < 4; ++q) {
                    uint32_t v[32], w[32];
                    const uint32_t taddr = tmem + lane_addr + D1_COL + buf * 128u + q * 32u;
                    tc_ld32(taddr, v);
                    tc_wait_ld();
                    const float* bias = embb + c * 128 + q * 32;
#pragma unroll
                    for (int e2 = 0; e2 < 16; ++e2) {
                        float x0 = __uint_as_float(v[2 * e2]) + bias[2 * e2];
                        float x1 = __uint_as_float(v[2 * e2 + 1]) + bias[2 * e2 + 1];
                        x0 = x0 > 0.f ? x0 : 0.f; x1 = x1 > 0.f ? x1 : 0.f;
                        const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                        const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
                        w[e2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        w[16 + e2] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    }
                    tc_st32(taddr, w);
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(bar(B_A2_FULL0 + buf));
                ++d1use;
            }
            // ---- next tile's one-hot operand, so its GEMM1 overlaps this tile's heads
            if (it + 1 < my_tiles) perm_next = build_a1(it + 1);

            // ---- epilogue 2: heads on CUDA cores from the fp32 accumulator row
            mbar_wait_t(bar(B_D2_FULL), it & 1, w_d2f, timed);
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
            const int64_t pos = tile_of(it) * TM + row;
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
        if (a.dbg && threadIdx.x == 64) {
            long long* d = a.dbg + blockIdx.x * 16;
            d[9] = clock64() - t_begin; d[10] = w_d1; d[11] = w_d2f; d[12] = w_a1e;
        }
    }

    // ---- teardown: everyone done with TMEM, then the allocating warp frees it
    tc_fence_before();
    __syncthreads();
    if (CSZ > 1) cluster_sync_all();       // no peer may still multicast into / arrive on this CTA's shared memory
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

}  // namespace


static int single_supported(const PolicyDev& p, const char** why) {
    static const char* m1 = "tensor-core forward needs obs_size <= 256 (one-hot operand must fit 64 KB of shared memory)";
    static const char* m2 = "tensor-core forward needs embedding size a multiple of 128, <= 1024";
    static const char* m3 = "tensor-core forward needs common width 128 or 256";
    static const char* m4 = "tensor-core forward needs <= 32 observation indices per env";
    if (p.obs_size > 64 * MAX_KB1) { *why = m1; return 0; }
    if (p.E % 128 || p.E > 1024 || p.E < 128) { *why = m2; return 0; }
    if (p.H != 128 && p.H != 256) { *why = m3; return 0; }
    if (p.n_obs > MAX_OBS) { *why = m4; return 0; }
    if (p.tc_compact_n > 0) { *why = m1; return 0; }
    return 1;
}

// either kernel: the CTA-pair kernel (twr_forward_tc2.cu; also takes the compact GridWorld table) or the single-CTA one
int forward_tc_supported(const PolicyDev& p, const EnvParams&, const char** why) {
    if (forward_tc2_supported(p) && p.n_obs <= MAX_OBS) return 1;
    return single_supported(p, why);
}

static size_t single_pack_bytes(const PolicyDev& p) {
    const char* why = "";
    PolicyDev q = p;
    q.n_obs = 1;
    if (!single_supported(q, &why)) return 0;
    const TcParams t = make_params(p);
    return (g1_tiles(t) + g2_tiles(t)) * TILE_BYTES;
}

// [single-CTA operand image][CTA-pair operand image]
size_t forward_tc_pack_bytes(const PolicyDev& p) { return single_pack_bytes(p) + forward_tc2_pack_bytes(p); }

void launch_forward_tc_pack(cudaStream_t st, const PolicyDev& p, void* pack) {
    if (single_pack_bytes(p)) {
        const TcParams t = make_params(p);
        k_tc_pack<<<1024, 256, 0, st>>>(p, t, reinterpret_cast<__half*>(pack));
        g_twr_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (forward_tc2_supported(p)) launch_forward_tc2_pack(st, p, reinterpret_cast<unsigned char*>(pack) + single_pack_bytes(p));
}

template <int NH, int CSZ>
void launch_cfg(cudaStream_t st, const PolicyDev& p, const ForwardArgs& a, const TcParams& t, int grid) {
    auto kern = k_forward_tc<NH, CSZ>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = SM_TOTAL; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CSZ; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, p, a, t);
}

int g_cluster = -1, g_pair = -1;

static void resolve_mode() {
    if (g_pair < 0) {
        const char* m = getenv("TWISTERL_B200_TC_MODE");      // "pair" (default) | "single"
        g_pair = (m && m[0] == 's') ? 0 : 1;
    }
}

int forward_tc_can_fuse(const PolicyDev& p) {
    resolve_mode();
    const char* f = getenv("TWISTERL_B200_FUSE_STEP");         // "0" keeps the separate k_collect_step launch
    if (f && f[0] == '0') return 0;
    return g_pair && forward_tc2_supported(p);
}

static bool use_pair(const PolicyDev& p) {
    resolve_mode();
    return (g_pair || !single_pack_bytes(p)) && forward_tc2_supported(p);
}

bool forward_tc_prepare(const PolicyDev& p) {
    if (!use_pair(p)) return true;
    return forward_tc2_prepare(p, reinterpret_cast<const unsigned char*>(p.tc_pack) + single_pack_bytes(p));
}

void forward_tc_forget(const PolicyDev& p) {
    if (p.tc_pack && forward_tc2_supported(p)) forward_tc2_forget(reinterpret_cast<const unsigned char*>(p.tc_pack) + single_pack_bytes(p));
}

bool launch_forward_tc(cudaStream_t st, const PolicyDev& p, const ForwardArgs& a) {
    if (a.n <= 0) return true;
    // no silent change of kernel: a pair-kernel launch that cannot be made is an error of the call
    if (use_pair(p)) return launch_forward_tc2(st, p, a, reinterpret_cast<const unsigned char*>(p.tc_pack) + single_pack_bytes(p));
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_cluster < 0) {
        const char* e = getenv("TWISTERL_B200_CLUSTER");
        g_cluster = e ? atoi(e) : 1;
        if (g_cluster != 1 && g_cluster != 2 && g_cluster != 4) g_cluster = 1;
    }
    const TcParams t = make_params(p);
    const int csz = g_cluster;
    const int n_tiles = (int)((a.n + TM - 1) / TM);
    const int n_groups = (n_tiles + csz - 1) / csz;
    const int max_clusters = sms / csz;
    const int grid = (n_groups < max_clusters ? n_groups : max_clusters) * csz;
    if (t.NH == 2) {
        if (csz == 1) launch_cfg<2, 1>(st, p, a, t, grid); else if (csz == 2) launch_cfg<2, 2>(st, p, a, t, grid); else launch_cfg<2, 4>(st, p, a, t, grid);
    } else {
        if (csz == 1) launch_cfg<1, 1>(st, p, a, t, grid); else if (csz == 2) launch_cfg<1, 2>(st, p, a, t, grid); else launch_cfg<1, 4>(st, p, a, t, grid);
    }
    g_twr_launches.fetch_add(1, std::memory_order_relaxed);
    return true;
}
