// twr_mcts.cu -- K6: batched Monte-Carlo tree search for the AlphaZero collector
// (rust/src/rl/search.rs:20-189, rust/src/rl/tree.rs, rust/src/collector/az.rs:51-109).
//
// One tree per env in a pre-sized node pool (node (e, i) at index e*P + i, root at i = 0;
// P = 1 + A*(n_sims*max(1,max_expand_depth) + 1) nodes is the reference's worst case).  A node stores what the
// reference's MCTSNode stores -- the cloned+stepped env (16-byte cells + meta) and ONE 16-byte record
// {visit_count, value_sum, prior, first_child | n_children | action_taken} -- plus a parent link.  The children of a
// node are contiguous, so the UCB descent reads one 64-byte line per tree level (statistics AND the links of the
// child it moves to), and the descent records its path so that the backup is a set of independent updates instead
// of a parent-pointer walk.  All envs advance in lockstep, one simulation at a time:
//   k_mcts_select : UCB descent to a leaf (search.rs:134-137, ucb :29-39, next :77-91); terminal leaves back up
//                   their reward at once, the others are compacted (warp ballot) into the leaf batch
//   forward       : the SAME policy forward kernel as the PPO path evaluates the whole leaf batch (K2)
//   k_mcts_expand : Policy::predict epilogue (masked exp / (sum+1e-6)), expand (search.rs:56-75), child draw
//                   proportional to the priors (next_sample :94-100), backup (search.rs:45-53)
// k_az_finish turns root visit counts into the MCTS policy (search.rs:166-188), draws the action
// (az.rs:72), writes the record and steps the env; k_az_remaining computes az.rs:93's reward-to-go.
// These kernels are latency / L2 bound (pointer chasing, one thread per env); the flop live in K2.
#include "twr_kernels.cuh"

#include <atomic>
extern std::atomic<long long> g_twr_launches;
#define TWR_COUNT_LAUNCH() g_twr_launches.fetch_add(1, std::memory_order_relaxed)

namespace {

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

__device__ __forceinline__ EnvState node_state(const MctsPool& m, int64_t g) {
    return env_load(m.cells, m.meta, g);
}

// node record: x = visit_count, y = value_sum (f32 bits), z = prior (f32 bits), w = first_child | n_children << 20 | action << 24
__device__ __forceinline__ uint32_t link_pack(int first, int nch, int action) {
    return ((uint32_t)first & 0xFFFFFu) | ((uint32_t)nch << 20) | (((uint32_t)action & 0xFFu) << 24);
}
__device__ __forceinline__ int link_first(uint32_t w) { return (int)(w & 0xFFFFFu); }
__device__ __forceinline__ int link_nch(uint32_t w) { return (int)((w >> 20) & 0xFu); }
__device__ __forceinline__ int link_action(uint32_t w) { return (int)(w >> 24); }

// one node of MCTSTree::backpropagate (search.rs:45-53): value_sum += v, visit_count += 1
__device__ __forceinline__ void bump(const MctsPool& m, int64_t g, float v) {
    uint2* q = reinterpret_cast<uint2*>(m.node + g);
    uint2 r = *q;
    r.x += 1u;
    r.y = __float_as_uint(__uint_as_float(r.y) + v);
    *q = r;
}
// MCTSTree::backpropagate over the recorded path (root .. node); falls back to the parent walk when the path
// outgrew its buffer
__device__ __forceinline__ void backprop(const MctsPool& m, int e, int64_t base, int node, int len, float v) {
    if (len < 0) {
        for (int b = node; b >= 0; b = m.parent[base + b]) bump(m, base + b, v);
        return;
    }
    // the nodes of a path are distinct: all loads of a block of 8 are issued before the first store
    for (int k0 = 0; k0 < len; k0 += 8) {
        int idx[8];
        uint2 r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) idx[j] = k0 + j < len ? m.path[(int64_t)(k0 + j) * m.B + e] : -1;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (idx[j] >= 0) r[j] = *reinterpret_cast<const uint2*>(m.node + base + idx[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (idx[j] >= 0) {
                r[j].x += 1u;
                r[j].y = __float_as_uint(__uint_as_float(r[j].y) + v);
                *reinterpret_cast<uint2*>(m.node + base + idx[j]) = r[j];
            }
        }
    }
}

// Policy::predict's epilogue (nn/policy.rs:43-47) on raw logits
__device__ __forceinline__ void masked_probs(const float4 raw, uint32_t mask, int A, float pr[4]) {
    const float l[4] = {raw.x, raw.y, raw.z, raw.w};
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { pr[i] = (i < A && ((mask >> i) & 1u)) ? expf(l[i]) : 0.0f; sum += pr[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = pr[i] / (sum + 0.000001f);
}

// MCTSTree::expand (search.rs:56-75): one child per action with prior > 0, in action order.  Returns the first
// child index and fills cp[] with the children's priors (cnt entries).
__device__ __forceinline__ int expand(const MctsPool& m, const EnvParams& env, int e, int64_t base, int node, const float pr[4],
                                      float cp[4], int& cnt) {
    const EnvState s = node_state(m, base + node);
    int n = m.n_nodes[e];
    const int first = n;
    cnt = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        cp[a] = 0.0f;
        if (a < m.A && pr[a] > 0.0f && n < m.P) {
            EnvState c = s;
            env_step(env, c, a);
            env_store(m.cells, m.meta, base + n, c);
            m.parent[base + n] = node;
            m.node[base + n] = make_uint4(0u, __float_as_uint(0.0f), __float_as_uint(pr[a]), link_pack(0, 0, a));
            // cp[] is indexed by child order; cnt <= a, so this never overwrites an entry still to be read
            if (cnt == 0) cp[0] = pr[a]; else if (cnt == 1) cp[1] = pr[a]; else if (cnt == 2) cp[2] = pr[a]; else cp[3] = pr[a];
            ++n; ++cnt;
        }
    }
    uint32_t* link = &m.node[base + node].w;
    *link = link_pack(first, cnt, link_action(*link));
    m.n_nodes[e] = n;
    return first;
}

// rand's WeightedIndex as used by nn/policy.rs:153-167
__device__ __forceinline__ int weighted_index(const float* w, int n, float u) {
    float tw = 0.0f;
    for (int i = 0; i < n; ++i) tw += w[i];
    if (!(tw > 0.0f)) return 0;
    const float chosen = u * tw;
    float cum = 0.0f;
    int last = 0;
    for (int i = 0; i < n; ++i) {
        if (w[i] > 0.0f) {
            cum += w[i]; last = i;
            if (cum > chosen) return i;
        }
    }
    return last;
}

// warp-aggregated append of (global node index, env) to the leaf batch
__device__ __forceinline__ void push_leaf(const MctsArgs& a, bool want, int64_t gnode, int e, int32_t* counter) {
    const unsigned bal = __ballot_sync(0xffffffffu, want);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (want) {
        const int slot = base + __popc(bal & ((1u << lane) - 1u));
        a.fwd_list[slot] = (int32_t)gnode;
        a.fwd_env[slot] = e;
        a.leaf_pos[e] = slot;
    }
}

__global__ void __launch_bounds__(128) k_mcts_begin(MctsArgs a, const int32_t* __restrict__ live, const int32_t* __restrict__ n_live) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = *n_live;
    if (pos == 0) { a.fwd_count[0] = nl; a.fwd_count[1] = 0; }
    if (pos >= nl) return;
    const int e = live[pos];
    const MctsPool& m = a.pool;
    const int64_t base = (int64_t)e * m.P;
    m.cells[base] = a.env_cells[e]; m.meta[base] = a.env_meta[e];
    m.parent[base] = -1;
    m.node[base] = make_uint4(1u, __float_as_uint(0.0f), __float_as_uint(0.0f), link_pack(0, 0, 0xFF));   // search.rs:119-125: root visit_count = 1
    m.n_nodes[e] = 1;
    m.path[e] = 0; m.path_len[e] = 1;                                          // path[0][e] = root
    a.fwd_list[pos] = (int32_t)base;
    a.fwd_env[pos] = e;
}

// mode 0: root expansion (search.rs:112-128); mode 1: simulation expansion round d (search.rs:142-160)
__device__ __forceinline__ void expand_body(const MctsArgs& a, int e, int pos, int mode, int sim, int d) {
    const MctsPool& m = a.pool;
    const int64_t base = (int64_t)e * m.P;
    const int node = mode == 0 ? 0 : a.cur_node[e];
    const EnvState s = node_state(m, base + node);
    float pr[4];
    masked_probs(a.logits[pos], env_masks(a.env, s), m.A, pr);
    float cp[4];
    int nch;
    const int first = expand(m, a.env, e, base, node, pr, cp, nch);
    if (mode == 0) return;
    // next_sample (search.rs:94-100): child drawn proportionally to the priors
    uint32_t w[4];
    const int med = a.max_expand_depth > 1 ? a.max_expand_depth : 1;
    philox4x32_10(a.ids.gid((uint32_t)e), (uint32_t)((a.t * (a.n_sims + 1) + sim) * med + d), TWR_RNG_MCTS, a.cid,
                  (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w);
    const int child = nch > 0 ? first + weighted_index(cp, nch, u32_to_unit_f32(w[0])) : node;
    if (a.trace) a.trace[((int64_t)sim * m.B + e) * 2 + 1] = child;
    const float v = a.values[pos];
    int len = m.path_len[e];
    if (child != node && len >= 0) {                     // the drawn child joins the recorded path
        if (len < TWR_MCTS_PATH) { m.path[(int64_t)len * m.B + e] = child; ++len; } else len = -1;
        m.path_len[e] = len;
    }
    if (d + 1 >= a.max_expand_depth) {
        backprop(m, e, base, child, len, v);
        a.active[e] = 0;
    } else {
        a.cur_node[e] = child;
        a.cur_value[e] = v;
    }
}
__global__ void __launch_bounds__(128) k_mcts_expand(MctsArgs a, int mode, int sim, int d, int which) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= a.fwd_count[which]) return;
    expand_body(a, a.fwd_env[pos], pos, mode, sim, d);
}

// descent + first rollout round (search.rs:132-146) of one env; returns whether its leaf needs a policy evaluation
__device__ __forceinline__ bool select_body(const MctsArgs& a, int e, int64_t& gnode, int sim) {
    bool want = false;
    {
        const MctsPool& m = a.pool;
        const int64_t base = (int64_t)e * m.P;
        int node = 0, len = 1;
        uint4 cur = m.node[base];                                              // root record (its path entry is permanent)
        while (link_nch(cur.w) > 0) {                                          // next(): argmax UCB, strict '>'
            const int first = link_first(cur.w), nch = link_nch(cur.w);
            const float sq = sqrtf((float)cur.x);
            uint4 ch[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k < nch) ch[k] = m.node[base + first + k];   // one 64-byte line
            int best = -1;
            float best_ucb = -INFINITY;
            uint4 best_rec = cur;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < nch) {
                    const uint32_t n = ch[k].x;
                    const float q = n == 0u ? 0.0f : __fdiv_rn(__uint_as_float(ch[k].y), (float)n);
                    const float ucb = __fadd_rn(q, __fmul_rn(__fmul_rn(a.C, __fdiv_rn(sq, __fadd_rn((float)n, 1.0f))), __uint_as_float(ch[k].z)));
                    if (ucb > best_ucb) { best = first + k; best_ucb = ucb; best_rec = ch[k]; }
                }
            }
            if (best < 0) break;
            node = best; cur = best_rec;
            if (len >= 0) {
                if (len < TWR_MCTS_PATH) { m.path[(int64_t)len * m.B + e] = node; ++len; } else len = -1;
            }
        }
        m.path_len[e] = len;
        a.cur_node[e] = node;
        if (a.trace && sim < a.n_sims) { a.trace[((int64_t)sim * m.B + e) * 2] = node; a.trace[((int64_t)sim * m.B + e) * 2 + 1] = node; }
        if (a.max_expand_depth <= 0) {
            backprop(m, e, base, node, len, 0.0f);                             // the rollout loop never runs: value stays 0
            a.active[e] = 0;
        } else {
            const EnvState s = node_state(m, base + node);
            if (env_is_final(a.env, s) || m.n_nodes[e] + m.A > m.P) {
                backprop(m, e, base, node, len, env_reward(a.env, s));
                a.active[e] = 0;
            } else {
                a.active[e] = 1;
                want = true;
                gnode = base + node;
            }
        }
    }
    return want;
}
__global__ void __launch_bounds__(128) k_mcts_select(MctsArgs a, const int32_t* __restrict__ live, const int32_t* __restrict__ n_live,
                                                     int sim, int which) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos == 0) a.fwd_count[which ^ 1] = 0;          // the other counter is idle until the next round
    const bool valid = pos < *n_live;
    bool want = false;
    int e = 0;
    int64_t gnode = 0;
    if (valid) {
        e = live[pos];
        want = select_body(a, e, gnode, sim);
    }
    push_leaf(a, want, gnode, e, a.fwd_count + which);
}
// max_expand_depth == 1: expansion + backup of simulation `sim` and the descent of simulation sim+1 in one launch
// (one thread per live env; the leaf batch of sim was compacted into counter `which`, the next goes into which^1)
__global__ void __launch_bounds__(128) k_mcts_expand_select(MctsArgs a, const int32_t* __restrict__ live,
                                                            const int32_t* __restrict__ n_live, int sim, int which) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos == 0) a.fwd_count[which] = 0;              // consumed by the forward before this launch; the next launch fills it
    const bool valid = pos < *n_live;
    bool want = false;
    int e = 0;
    int64_t gnode = 0;
    if (valid) {
        e = live[pos];
        if (a.active[e]) expand_body(a, e, a.leaf_pos[e], 1, sim, 0);
        want = select_body(a, e, gnode, sim + 1);
    }
    push_leaf(a, want, gnode, e, a.fwd_count + (which ^ 1));
}

// rollout rounds d > 0 (search.rs:142-146 for the child sampled in the previous round)
__global__ void __launch_bounds__(128) k_mcts_pre(MctsArgs a, const int32_t* __restrict__ live, const int32_t* __restrict__ n_live, int which) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos == 0) a.fwd_count[which ^ 1] = 0;
    const bool valid = pos < *n_live;
    bool want = false;
    int e = 0;
    int64_t gnode = 0;
    if (valid) {
        e = live[pos];
        if (a.active[e]) {
            const MctsPool& m = a.pool;
            const int64_t base = (int64_t)e * m.P;
            const int node = a.cur_node[e];
            const EnvState s = node_state(m, base + node);
            if (env_is_final(a.env, s) || m.n_nodes[e] + m.A > m.P) {
                backprop(m, e, base, node, m.path_len[e], env_reward(a.env, s));
                a.active[e] = 0;
            } else {
                want = true;
                gnode = base + node;
            }
        }
    }
    push_leaf(a, want, gnode, e, a.fwd_count + which);
}

// root visit counts -> MCTS policy (search.rs:166-188); used by the debug entry point
__device__ __forceinline__ void root_probs(const MctsPool& m, int64_t base, float pr[4], int32_t vis[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { pr[i] = 0.0f; vis[i] = 0; }
    const uint32_t link = m.node[base].w;
    const int first = link_first(link), nch = link_nch(link);
    for (int k = 0; k < nch; ++k) {
        const uint4 rec = m.node[base + first + k];
        const int act = link_action(rec.w);
        const uint32_t n = rec.x;
        const float f = (float)n;
        if (act == 0) { pr[0] = f; vis[0] = (int32_t)n; } else if (act == 1) { pr[1] = f; vis[1] = (int32_t)n; }
        else if (act == 2) { pr[2] = f; vis[2] = (int32_t)n; } else { pr[3] = f; vis[3] = (int32_t)n; }
    }
    float sum = 0.0f;
    for (int i = 0; i < m.A; ++i) sum += pr[i];
    if (sum > 0.0f) { for (int i = 0; i < 4; ++i) pr[i] = i < m.A ? __fdiv_rn(pr[i], sum) : 0.0f; }
    else { for (int i = 0; i < 4; ++i) pr[i] = i < m.A ? __fdiv_rn(1.0f, (float)m.A) : 0.0f; }
}

__global__ void __launch_bounds__(128) k_mcts_read(MctsArgs a, int64_t n, float* __restrict__ probs, int32_t* __restrict__ visits) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float pr[4]; int32_t vis[4];
    root_probs(a.pool, e * a.pool.P, pr, vis);
    for (int i = 0; i < a.pool.A; ++i) { probs[e * a.pool.A + i] = pr[i]; visits[e * a.pool.A + i] = vis[i]; }
}

// one AZCollector loop iteration (az.rs:66-88) for every live env
__global__ void __launch_bounds__(128) k_az_finish(MctsArgs a, CollectBuffers b, const int32_t* __restrict__ live,
                                                   int32_t* __restrict__ live_next) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = b.n_live[a.t];
    bool survives = false;
    int e = 0;
    if (pos < nl) {
        e = live[pos];
        float pr[4]; int32_t vis[4];
        root_probs(a.pool, (int64_t)e * a.pool.P, pr, vis);
        uint32_t w[4];
        philox4x32_10(a.ids.gid((uint32_t)e), (uint32_t)a.t, TWR_RNG_AZ_ACT, a.cid, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w);
        const int act = weighted_index(pr, a.pool.A, u32_to_unit_f32(w[0]));       // sample(&mcts_probs), az.rs:72
        EnvState s = env_load(b.cells, b.meta, e);
        const float rew = env_reward(a.env, s);
        const int64_t r = (int64_t)a.t * b.B + e;
        b.rec_state[r] = env_pack_cells(s);
        b.rec_logits[r] = make_float4(pr[0], pr[1], pr[2], pr[3]);               // CollectedData.logits holds the MCTS probs
        b.rec_reward[r] = rew;
        b.rec_action[r] = (uint8_t)act;
        b.rec_perm[r] = (int8_t)-1;
        b.rec_value[r] = 0.0f; b.rec_adv[r] = 0.0f;
        if (env_is_final(a.env, s)) {
            b.ep_len[e] = a.t + 1;
            if (env_success(a.env, s)) atomicAdd(&b.stats[0], 1ull);
            atomicAdd(reinterpret_cast<double*>(&b.stats[2]), (double)rew);
        } else {
            env_step(a.env, s, act);
            env_store(b.cells, b.meta, e, s);
            survives = true;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, survives);
    if (bal) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(&b.n_live[a.t + 1], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (survives) live_next[base + __popc(bal & ((1u << lane) - 1u))] = e;
    }
}

// remaining_values[t] = total - sum_{k<t} r_k, f32 in the reference's order (az.rs:74-76,93)
__global__ void __launch_bounds__(256) k_az_remaining(CollectBuffers b) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= b.B) return;
    const int n = b.ep_len[e];
    float total = 0.0f;
    for (int t = 0; t < n; ++t) total = __fadd_rn(total, b.rec_reward[(int64_t)t * b.B + e]);
    float prefix = 0.0f;
    for (int t = 0; t < n; ++t) {
        const int64_t r = (int64_t)t * b.B + e;
        b.rec_ret[r] = __fsub_rn(total, prefix);
        prefix = __fadd_rn(prefix, b.rec_reward[r]);
    }
}

}  // namespace

void launch_mcts_begin(cudaStream_t st, const MctsArgs& a, const int32_t* live, const int32_t* n_live, int64_t max_n) {
    k_mcts_begin<<<grid_for(max_n, 128), 128, 0, st>>>(a, live, n_live);
    TWR_COUNT_LAUNCH();
}
void launch_mcts_expand(cudaStream_t st, const MctsArgs& a, int mode, int sim, int d, int which, int64_t max_n) {
    k_mcts_expand<<<grid_for(max_n, 128), 128, 0, st>>>(a, mode, sim, d, which);
    TWR_COUNT_LAUNCH();
}
void launch_mcts_select(cudaStream_t st, const MctsArgs& a, const int32_t* live, const int32_t* n_live, int sim, int which, int64_t max_n) {
    k_mcts_select<<<grid_for(max_n, 128), 128, 0, st>>>(a, live, n_live, sim, which);
    TWR_COUNT_LAUNCH();
}
void launch_mcts_expand_select(cudaStream_t st, const MctsArgs& a, const int32_t* live, const int32_t* n_live, int sim, int which, int64_t max_n) {
    k_mcts_expand_select<<<grid_for(max_n, 128), 128, 0, st>>>(a, live, n_live, sim, which);
    TWR_COUNT_LAUNCH();
}
void launch_mcts_pre(cudaStream_t st, const MctsArgs& a, const int32_t* live, const int32_t* n_live, int which, int64_t max_n) {
    k_mcts_pre<<<grid_for(max_n, 128), 128, 0, st>>>(a, live, n_live, which);
    TWR_COUNT_LAUNCH();
}
void launch_mcts_read(cudaStream_t st, const MctsArgs& a, int64_t n, float* probs, int32_t* visits) {
    k_mcts_read<<<grid_for(n, 128), 128, 0, st>>>(a, n, probs, visits);
    TWR_COUNT_LAUNCH();
}
void launch_az_finish(cudaStream_t st, const MctsArgs& a, const CollectBuffers& b, const int32_t* live, int32_t* live_next) {
    k_az_finish<<<grid_for(b.B, 128), 128, 0, st>>>(a, b, live, live_next);
    TWR_COUNT_LAUNCH();
}
void launch_az_remaining(cudaStream_t st, const CollectBuffers& b) {
    k_az_remaining<<<grid_for(b.B, 256), 256, 0, st>>>(b);
    TWR_COUNT_LAUNCH();
}
