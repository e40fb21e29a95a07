// Thread-per-tree vanilla MCTS engine: the root-parallel hot path (BASELINE config C5:
// 65 536 independent CartPole trees x n_sim simulations; also C1 FrozenLake).
//
// One thread owns one tree and runs the reference's whole `for s in range(n_sim)` loop
// (islaMcts/agents/mcts_hash.py:30-34) in a single launch:
//   selection   StateNodeHash.build_tree_state   mcts_hash.py:59-83  (= mcts.py:46-67)
//   UCB1        ucb1                             utils/action_selection_functions.py:9-25
//   expansion   ActionNodeHash.build_tree_action mcts_hash.py:87-131 (= mcts.py:72-120)
//   rollout     AbstractStateNode.rollout        agents/abstract_mcts.py:72-91
//   backup      the recursion unwind             mcts_hash.py:79-81,119-131
//
// B200-first design instead of the reference's object-per-node recursion:
//   * SoA-of-blocks device tree in HBM: the A edge records (32 B each) of one parent are one aligned
//     block, so one UCB scan = one 64 B (A=2) / 128 B (A=4) contiguous load; a child's record carries
//     the block index of ITS children, so the pointer chase is one dependent load per tree level.
//   * states are stored per node (no env replay from the root: the reference re-steps the env on
//     every tree level, 186 of 205 env.step calls per simulation in SURVEY 3.3).
//   * descent records (edge slot, reward) in a level-major scratch stack (coalesced across the
//     warp); backup walks it backwards with fire-and-forget fp64/int RED atomics -- no re-read.
//   * visit counts / totals are exact: int32 counts, fp64 totals, fp64 UCB with separately rounded
//     operations (the reference computes in float64), ln(n) from a per-CTA shared-memory table.
//   * env state and rollouts live in registers; Philox4x32-10 streams replace numpy/gym RNGs.
#include <cmath>
#include <cstring>

#include "isla_envs.cuh"
#include "isla_internal.h"
#include "isla_philox.cuh"

namespace isla {

namespace {

constexpr int kTptBlock = 64;

struct TraceCursor {
    const uint8_t* k;
    const double* v;
    long long len, pos;
    int fault;
    __device__ __forceinline__ double take(int kind) {
        if (fault) return 0.0;
        if (pos >= len) { fault = ISLA_ERR_TRACE; return 0.0; }
        if (k[pos] != kind) { fault = ISLA_ERR_TRACE; return 0.0; }
        return v[pos++];
    }
};

__device__ __forceinline__ int4 ldcg_256(const int4* p) {
    int4 v;
    asm volatile("ld.global.cg.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <bool PF256 = false>
__device__ __forceinline__ EdgeRec load_rec(const EdgeRec* p) {
    // L2-only loads: the records are updated through RED atomics at L2, never reuse L1.
    // PF256: ask L2 to fetch the whole 256-byte neighbourhood -- the blocks of the principal
    // variation are allocated consecutively, so the next tree levels usually live there.
    const int4 lo = PF256 ? ldcg_256(reinterpret_cast<const int4*>(p)) : __ldcg(reinterpret_cast<const int4*>(p));
    const int4 hi = __ldcg(reinterpret_cast<const int4*>(p) + 1);
    EdgeRec r;
    r.total_in = __hiloint2double(lo.y, lo.x);
    r.g_first = __hiloint2double(lo.w, lo.z);
    r.na_in = hi.x;
    r.r_in = __int_as_float(hi.y);
    r.block = hi.z;
    r.flags = (uint32_t)hi.w;
    return r;
}
__device__ __forceinline__ void store_rec(EdgeRec* p, const EdgeRec& r) {
    int4 lo, hi;
    lo.x = __double2loint(r.total_in); lo.y = __double2hiint(r.total_in);
    lo.z = __double2loint(r.g_first);  lo.w = __double2hiint(r.g_first);
    hi.x = r.na_in; hi.y = __float_as_int(r.r_in); hi.z = r.block; hi.w = (int)r.flags;
    __stcg(reinterpret_cast<int4*>(p), lo);
    __stcg(reinterpret_cast<int4*>(p) + 1, hi);
}

// One simulation = three warp-converged phases.  Every lane owns a tree, but the lanes of a warp walk
// their phases in lock-step so that the SIMT width is not lost to divergence:
//   phase 1  descent      all lanes advance one tree level per iteration until each has found its
//                          leaf parent (or a terminal edge); lanes that are done idle until the deepest is
//   phase 2  expand+roll   the pending expansions step the env once, then all rollouts advance one env
//                          step per iteration (state in registers)
//   phase 3  backup        levels are walked backwards in lock-step, path entries fetched four at a time
enum : int { MODE_DESCEND = 0, MODE_EXPAND = 1, MODE_BACKUP = 2, MODE_IDLE = 3 };

template <int ENV, typename T, bool SCRIPTED>
__global__ void __launch_bounds__(kTptBlock) tpt_search_kernel(const TptParams p) {
    constexpr int A = EnvTraits<ENV>::A;
    constexpr int S = EnvTraits<ENV>::S;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ double ln_smem[];  // copy of the host-computed ln(n) table when it fits
    if (p.ln_in_smem) {
        for (int i = threadIdx.x; i < p.ln_count; i += blockDim.x) ln_smem[i] = p.ln_table[i];
        __syncthreads();
    }
    const double* const ln_table = p.ln_in_smem ? ln_smem : p.ln_table;

    // scripted (parity) mode drives ONE tree from lane 0 of a single warp; the other lanes are dummies
    const long long tree = SCRIPTED ? (threadIdx.x == 0 ? p.scripted_tree : p.n_trees)
                                    : (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = tree < p.n_trees;
    const long long tix = valid ? tree : 0;

    EdgeRec* const recs = p.recs + tix * p.slots_per_tree;
    T* const states = reinterpret_cast<T*>(p.states) + tix * p.slots_per_tree * S;
    int32_t* const meta = p.meta + tix * META_WORDS;
    int2* const path = p.path + tix;  // level-major: path[level * n_trees]

    int alloc = valid ? meta[META_ALLOC] : 0;
    uint32_t sims_done = valid ? (uint32_t)meta[META_SIMS] : 0u;
    int fault = valid ? meta[META_FAULT] : 0;
    const uint32_t tree_id = (uint32_t)(p.tree_id_base + (unsigned long long)tix);
    Stream ts;
    ts.resume(p.seed, 0u, tree_id, kStreamTree, valid ? (uint32_t)meta[META_RNG] : 0u);
    TraceCursor tc{p.trace_kinds, p.trace_values, p.trace_len, 0, 0};

    unsigned long long c_env = 0, c_roll = 0, c_levels = 0, c_nodes = 0;
    const double gamma = p.gamma, Cexp = p.C;
    const float Cf = (float)p.C;

    auto choice = [&](int n) -> int {  // np.random.choice(candidates) -> position
        if constexpr (SCRIPTED) {
            const int v = (int)tc.take(ISLA_K_CHOICE);
            if (v < 0 || v >= n) { if (!tc.fault) tc.fault = ISLA_ERR_TRACE; return 0; }
            return v;
        } else {
            return n > 1 ? (int)mulhi_u32(ts.next(), (uint32_t)n) : 0;
        }
    };

    int s_done = 0;
    for (int sim = 0; sim < p.n_sim; ++sim) {
        const bool live = valid && !fault;
        int mode = live ? MODE_DESCEND : MODE_IDLE;
        int node = 0, blk = 1, level = 0, exp_a = 0, exp_rank = 0;
        bool fresh = false;  // block just allocated: all children unvisited, nothing to load
        double G = 0.0;

        // ------------------------------------------------------------------ phase 1: descent
        while (__any_sync(kFull, mode == MODE_DESCEND)) {
            if (mode == MODE_DESCEND) {
                ++c_levels;
                EdgeRec c[A];
                int nun = 0;
                if (fresh) {
#pragma unroll
                    for (int a = 0; a < A; ++a) { c[a] = EdgeRec{0.0, 0.0, 0, 0.f, 0, 0u}; }
                    nun = A;
                } else {
#pragma unroll
                    for (int a = 0; a < A; ++a) {
                        c[a] = (p.tuning & 2) ? load_rec<true>(recs + (long long)blk * A + a) : load_rec<false>(recs + (long long)blk * A + a);
                        nun += (c[a].na_in == 0);
                    }
                }
                if (nun > 0) {
                    // uniformly random unvisited action (mcts_hash.py:70-74); expanded in phase 2
                    const int pos = choice(nun);
                    int seen = 0;
#pragma unroll
                    for (int a = 0; a < A; ++a) { if (c[a].na_in == 0) { if (seen == pos) exp_a = a; ++seen; } }
                    exp_rank = A - nun;
                    mode = MODE_EXPAND;
                } else {
                    // UCB1 over the A visited children (action_selection_functions.py:9-25).  The reference
                    // evaluates the scores in float64 and breaks exact ties at random.  An fp32 screening pass
                    // (error < 1e-6 relative, tolerance 1e-5) settles the arg-max whenever it is unambiguous;
                    // only near-ties take the exact float64 path, so the DECISIONS are the float64 ones.
                    int ns = (node != 0);
#pragma unroll
                    for (int a = 0; a < A; ++a) ns += c[a].na_in;
                    const double lg = ns < p.ln_count ? ln_table[ns] : ::log((double)ns);
                    int a_sel = 0;
                    bool settled = false;
                    if (!(p.tuning & 1)) {
                        const float lgf = (float)lg;
                        float sf[A], m1 = -INFINITY, scale = 0.f;
#pragma unroll
                        for (int a = 0; a < A; ++a) {
                            const float naf = (float)c[a].na_in;
                            const float qf = __fdividef((float)c[a].total_in, naf);
                            const float ef = Cf * sqrtf(__fdividef(lgf, naf));
                            sf[a] = qf + ef;
                            m1 = fmaxf(m1, sf[a]);
                            scale = fmaxf(scale, fabsf(qf) + fabsf(ef));
                        }
                        const float thr = m1 - 1e-5f * scale - 1e-30f;
                        int ncand = 0;
#pragma unroll
                        for (int a = 0; a < A; ++a) { if (sf[a] >= thr) { ++ncand; a_sel = a; } }
                        settled = (ncand == 1);
                    }
                    if (!settled) {
                        double sc[A], best = -INFINITY;
#pragma unroll
                        for (int a = 0; a < A; ++a) {
                            const double na = (double)c[a].na_in;
                            const double q = __ddiv_rn(c[a].total_in, na);
                            const double ex = __dmul_rn(Cexp, __dsqrt_rn(__ddiv_rn(lg, na)));
                            sc[a] = __dadd_rn(q, ex);
                            best = fmax(best, sc[a]);
                        }
                        int ties = 0;
#pragma unroll
                        for (int a = 0; a < A; ++a) ties += (sc[a] == best);
                        int pos = (SCRIPTED || ties > 1) ? choice(ties) : 0;
                        // the tied set is ordered like node.actions (insertion order = rank)
#pragma unroll
                        for (int rk = 0; rk < A; ++rk) {
#pragma unroll
                            for (int a = 0; a < A; ++a) {
                                if ((int)((c[a].flags >> 8) & 0xffu) == rk && sc[a] == best) { if (pos == 0) a_sel = a; --pos; }
                            }
                        }
                    } else if constexpr (SCRIPTED) {
                        (void)choice(1);  // the reference draws even when the maximum is unique
                    }
                    const int child = blk * A + a_sel;
                    EdgeRec sel = c[0];
#pragma unroll
                    for (int a = 1; a < A; ++a) if (a == a_sel) sel = c[a];
                    if (sel.flags & kFlagTerminal) {
                        // terminal child re-visited: total += r, na += 1 (mcts_hash.py:96-107)
                        G = (double)sel.r_in;
                        atomicAdd(&recs[child].total_in, G);
                        atomicAdd(&recs[child].na_in, 1);
                        mode = MODE_BACKUP;
                    } else {
                        path[(long long)level * p.n_trees] = make_int2(child, __float_as_int(sel.r_in));
                        ++level;
                        node = child;
                        blk = sel.block;
                        fresh = false;
                        if (blk == 0) {
                            if (alloc >= p.block_cap) { fault = ISLA_ERR_CAPACITY; mode = MODE_IDLE; }
                            else { blk = alloc++; recs[child].block = blk; fresh = true; }
                        }
                    }
                }
            }
        }

        // ------------------------------------------------------------------ phase 2: expansion + rollout
        T st[S];
        T r_exp = T(0);
        bool rolling = false;
        int budget = 0;
        if (mode == MODE_EXPAND) {
            if constexpr (S == 4 && sizeof(T) == 4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(states + (long long)node * S));
                st[0] = v.x; st[1] = v.y; st[2] = v.z; st[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < S; ++i) st[i] = __ldcg(states + (long long)node * S + i);
            }
            const bool term = Env<ENV, T>::step(st, T(exp_a), r_exp);
            ++c_env;
            if (term) {
                G = (double)r_exp;  // terminal child: total += r, na += 1, child.ns += 1 (mcts_hash.py:96-107)
            } else {
                const int child = blk * A + exp_a;
                if constexpr (S == 4 && sizeof(T) == 4) {
                    __stcg(reinterpret_cast<float4*>(states + (long long)child * S), make_float4(st[0], st[1], st[2], st[3]));
                } else {
#pragma unroll
                    for (int i = 0; i < S; ++i) __stcg(states + (long long)child * S + i, st[i]);
                }
                budget = p.budget_mode == ISLA_BUDGET_ZERO ? 0 : min(p.max_depth - (level + 1), p.disc_count);  // mcts_hash.py:116
                rolling = true;
            }
        } else {
#pragma unroll
            for (int i = 0; i < S; ++i) st[i] = T(0);
        }
        // rollout (abstract_mcts.py:72-91): one env step per iteration for every lane still playing
        {
            double R = 0.0;
            Stream rs;
            rs.init(p.seed, sims_done, tree_id, 0u);
            int t = 0;
            bool go = rolling && budget > 0;
            while (__any_sync(kFull, go)) {
                if (go) {
                    T act;
                    if constexpr (SCRIPTED) {
                        const int v = (int)tc.take(ISLA_K_SAMPLE_D);
                        if (v < 0 || v >= A) { if (!tc.fault) tc.fault = ISLA_ERR_TRACE; }
                        act = T(tc.fault ? 0 : v);
                    } else {
                        act = sample_action<ENV, T>(rs);
                    }
                    T rr;
                    const bool done = Env<ENV, T>::step(st, act, rr);
                    R = __dadd_rn(R, __dmul_rn((double)rr, __ldg(p.disc_table + t)));  // r * pow(gamma, t)
                    ++t;
                    go = !done && t < budget;
                    if constexpr (SCRIPTED) { if (tc.fault) go = false; }
                }
            }
            c_roll += t; c_env += t;
            if (rolling) G = __dadd_rn((double)r_exp, __dmul_rn(gamma, R));
        }
        if (mode == MODE_EXPAND) {
            EdgeRec nr;
            nr.na_in = 1; nr.r_in = (float)r_exp; nr.block = 0; nr.flags = (uint32_t)exp_rank << 8;
            nr.total_in = G;
            if (rolling) { nr.g_first = G; } else { nr.g_first = 0.0; nr.flags |= kFlagTerminal; }
            store_rec(recs + blk * A + exp_a, nr);
            ++c_nodes;
            mode = MODE_BACKUP;
        }
        if constexpr (SCRIPTED) { if (tc.fault && !fault) { fault = tc.fault; mode = MODE_IDLE; } }

        // ------------------------------------------------------------------ phase 3: backup
        // G <- r + gamma*G up the recorded path (mcts_hash.py:126-131, 79-81); fire-and-forget REDs
        const int my_levels = (mode == MODE_BACKUP) ? level : 0;
        int max_levels = my_levels;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_levels = max(max_levels, __shfl_xor_sync(kFull, max_levels, o));
        for (int l0 = max_levels - 1; l0 >= 0; l0 -= 4) {
            int2 e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int l = l0 - j;
                if (l >= 0 && l < my_levels) e[j] = __ldcg(path + (long long)l * p.n_trees);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int l = l0 - j;
                if (l >= 0 && l < my_levels) {
                    G = __dadd_rn((double)__int_as_float(e[j].y), __dmul_rn(gamma, G));
                    atomicAdd(&recs[e[j].x].total_in, G);
                    atomicAdd(&recs[e[j].x].na_in, 1);
                }
            }
        }
        if (mode == MODE_BACKUP) { ++sims_done; ++s_done; }
    }

    if (valid) {
        meta[META_ALLOC] = alloc;
        meta[META_SIMS] = (int32_t)sims_done;
        meta[META_RNG] = (int32_t)ts.consumed();
        meta[META_FAULT] = fault;
        if constexpr (SCRIPTED) meta[4] = (int32_t)tc.pos;
        // counters (once per thread at kernel end)
        atomicAdd(p.counters + CNT_ENV_STEPS, c_env);
        atomicAdd(p.counters + CNT_ROLLOUT_STEPS, c_roll);
        atomicAdd(p.counters + CNT_LEVELS, c_levels);
        atomicAdd(p.counters + CNT_SIMS, (unsigned long long)s_done);
        atomicAdd(p.counters + CNT_NODES, c_nodes);
        if (fault) atomicAdd(p.counters + CNT_FAULTS, 1ull);
    }
}

template <typename T>
__global__ void tpt_set_roots_kernel(const double* roots, T* states, int32_t* meta, long long n_trees,
                                     long long slots_per_tree, int S) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    for (int i = 0; i < S; ++i) states[tree * slots_per_tree * S + i] = (T)roots[tree * S + i];
    int32_t* m = meta + tree * META_WORDS;
    m[META_ALLOC] = 2;  // block 0 holds the root slot, block 1 the root's edges
    m[META_SIMS] = 0; m[META_RNG] = 0; m[META_FAULT] = 0; m[4] = 0; m[5] = 0; m[6] = 0; m[7] = 0;
}

__global__ void tpt_root_stats_kernel(const EdgeRec* recs, long long n_trees, long long slots_per_tree, int A,
                                      int64_t* na_out, double* total_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_trees * A) return;
    const long long tree = i / A;
    const int a = (int)(i % A);
    const EdgeRec r = load_rec(recs + tree * slots_per_tree + A + a);  // block 1
    na_out[i] = r.na_in;
    total_out[i] = r.total_in;
}

// end of fit(): q = total/na over the root's children, uniformly random arg-max (mcts_hash.py:42-50)
__global__ void tpt_best_kernel(const EdgeRec* recs, int32_t* meta, long long n_trees, long long slots_per_tree, int A,
                                unsigned long long seed, unsigned long long tree_id_base, int scripted_pos,
                                int32_t* best_out, double* best_action_out) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    double q[16];
    int rank[16];
    double best = -INFINITY;
    int ties = 0;
    for (int a = 0; a < A; ++a) {
        const EdgeRec r = load_rec(recs + tree * slots_per_tree + A + a);
        rank[a] = r.na_in > 0 ? (int)((r.flags >> 8) & 0xffu) : -1;
        q[a] = r.na_in > 0 ? __ddiv_rn(r.total_in, (double)r.na_in) : -INFINITY;
        if (r.na_in > 0) best = fmax(best, q[a]);
    }
    for (int a = 0; a < A; ++a) ties += (rank[a] >= 0 && q[a] == best);
    if (ties == 0) { best_out[tree] = -1; if (best_action_out) best_action_out[tree] = 0.0; return; }
    int pos = 0;
    if (scripted_pos >= 0) pos = scripted_pos < ties ? scripted_pos : ties - 1;
    else if (ties > 1) {
        int32_t* m = meta + tree * META_WORDS;
        Stream ts;
        ts.resume(seed, 0u, (uint32_t)(tree_id_base + (unsigned long long)tree), kStreamTree, (uint32_t)m[META_RNG]);
        pos = (int)mulhi_u32(ts.next(), (uint32_t)ties);
        m[META_RNG] = (int32_t)ts.consumed();
    }
    int sel = -1;
    for (int rk = 0; rk < A; ++rk)
        for (int a = 0; a < A; ++a)
            if (rank[a] == rk && q[a] == best) { if (pos == 0) sel = a; --pos; }
    best_out[tree] = sel >= 0 ? rank[sel] : -1;
    if (best_action_out) best_action_out[tree] = (double)sel;
}

template <int ENV, typename T>
int launch_search(const TptParams& p, bool scripted, cudaStream_t st) {
    const size_t smem = p.ln_in_smem ? (size_t)p.ln_count * sizeof(double) : 0;
    if (scripted) {
        tpt_search_kernel<ENV, T, true><<<1, 32, smem, st>>>(p);
    } else {
        const unsigned grid = (unsigned)((p.n_trees + kTptBlock - 1) / kTptBlock);
        tpt_search_kernel<ENV, T, false><<<grid, kTptBlock, smem, st>>>(p);
    }
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : cuda_fail(err, "tpt_search_kernel launch");
}

}  // namespace

int tpt_fill_layout(const isla_config& cfg, isla_layout& lay) {
    const int A = env_n_actions(cfg.env_id), S = env_state_dim(cfg.env_id);
    if (A == 0 || cfg.agent != ISLA_AGENT_VANILLA) { set_error("thread-per-tree kernel: vanilla agent on a discrete env only"); return ISLA_ERR_INVALID; }
    if (cfg.n_actions != A) { set_error("n_actions=%d does not match env (%d)", cfg.n_actions, A); return ISLA_ERR_INVALID; }
    if (cfg.rollouts_per_leaf > 1) { set_error("thread-per-tree kernel: rollouts_per_leaf must be 1"); return ISLA_ERR_INVALID; }
    const int64_t blocks = (int64_t)cfg.sim_capacity + 2;
    const int rb = cfg.precision == ISLA_PRECISION_F64 ? 8 : 4;
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    lay.kernel = ISLA_KERNEL_THREAD_PER_TREE;
    lay.state_dim = S; lay.real_bytes = rb; lay.n_actions = A;
    lay.slots_per_tree = blocks * A;
    int64_t off = 0;
    lay.rec_offset = off; lay.rec_stride = lay.slots_per_tree * 32; off = up(off + cfg.n_trees * lay.rec_stride);
    lay.state_offset = off; lay.state_stride = lay.slots_per_tree * S * rb; off = up(off + cfg.n_trees * lay.state_stride);
    lay.meta_offset = off; off = up(off + cfg.n_trees * META_WORDS * 4);
    lay.path_offset = off; off = up(off + cfg.n_trees * blocks * 8);
    lay.counters_offset = off; off = up(off + 8 * 8);
    add_table_layout(cfg, lay, off);
    lay.s_cap = lay.a_cap = 0; lay.pool_offset = lay.pool_stride = 0;
    lay.total_bytes = off;
    return 0;
}

int tpt_set_roots(isla_engine* e, const double* roots_dev, cudaStream_t st) {
    const isla_layout& L = e->lay;
    // records + states + meta + counters cleared; the path scratch needs no clearing
    ISLA_CUDA_OK(cudaMemsetAsync(e->ws + L.rec_offset, 0, (size_t)(e->cfg.n_trees * L.rec_stride), st));
    ISLA_CUDA_OK(cudaMemsetAsync(e->ws + L.counters_offset, 0, 64, st));
    const unsigned grid = (unsigned)((e->cfg.n_trees + 255) / 256);
    if (L.real_bytes == 4)
        tpt_set_roots_kernel<float><<<grid, 256, 0, st>>>(roots_dev, (float*)(e->ws + L.state_offset), (int32_t*)(e->ws + L.meta_offset),
                                                          e->cfg.n_trees, L.slots_per_tree, L.state_dim);
    else
        tpt_set_roots_kernel<double><<<grid, 256, 0, st>>>(roots_dev, (double*)(e->ws + L.state_offset), (int32_t*)(e->ws + L.meta_offset),
                                                           e->cfg.n_trees, L.slots_per_tree, L.state_dim);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int tpt_run(isla_engine* e, int n_sim, long long scripted_tree, const uint8_t* kinds, const double* values,
            long long trace_len, cudaStream_t st) {
    const isla_layout& L = e->lay;
    TptParams p;
    p.recs = (EdgeRec*)(e->ws + L.rec_offset);
    p.states = e->ws + L.state_offset;
    p.meta = (int32_t*)(e->ws + L.meta_offset);
    p.path = (int2*)(e->ws + L.path_offset);
    p.counters = (unsigned long long*)(e->ws + L.counters_offset);
    p.trace_kinds = kinds; p.trace_values = values; p.trace_len = trace_len;
    p.n_trees = e->cfg.n_trees; p.slots_per_tree = L.slots_per_tree; p.scripted_tree = scripted_tree;
    p.seed = e->cfg.seed; p.tree_id_base = e->cfg.tree_id_base;
    p.C = e->cfg.C; p.gamma = e->cfg.gamma;
    p.n_sim = n_sim; p.max_depth = e->cfg.max_depth; p.budget_mode = e->cfg.budget_mode;
    p.block_cap = e->cfg.sim_capacity + 2;
    p.ln_table = (const double*)(e->ws + L.ln_offset);
    p.disc_table = (const double*)(e->ws + L.disc_offset);
    p.ln_count = (int)L.ln_count; p.disc_count = (int)L.disc_count;
    p.ln_in_smem = (L.ln_count * 8 <= 32 * 1024) ? 1 : 0;
    p.tuning = e->cfg.tuning;
    const bool scripted = scripted_tree >= 0;
    const bool f64 = e->cfg.precision == ISLA_PRECISION_F64;
    switch (e->cfg.env_id) {
    case ISLA_ENV_CARTPOLE:
        return f64 ? launch_search<ENV_CARTPOLE, double>(p, scripted, st) : launch_search<ENV_CARTPOLE, float>(p, scripted, st);
    case ISLA_ENV_FROZENLAKE:
        return f64 ? launch_search<ENV_FROZENLAKE, double>(p, scripted, st) : launch_search<ENV_FROZENLAKE, float>(p, scripted, st);
    default:
        set_error("thread-per-tree kernel: env %d unsupported", e->cfg.env_id);
        return ISLA_ERR_INVALID;
    }
}

int tpt_root_stats(isla_engine* e, int64_t* na_dev, double* total_dev, cudaStream_t st) {
    const isla_layout& L = e->lay;
    const long long n = e->cfg.n_trees * L.n_actions;
    tpt_root_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const EdgeRec*)(e->ws + L.rec_offset), e->cfg.n_trees,
                                                                       L.slots_per_tree, L.n_actions, na_dev, total_dev);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int tpt_best(isla_engine* e, int scripted_pos, int32_t* best_dev, double* best_action_dev, cudaStream_t st) {
    const isla_layout& L = e->lay;
    tpt_best_kernel<<<(unsigned)((e->cfg.n_trees + 255) / 256), 256, 0, st>>>(
        (const EdgeRec*)(e->ws + L.rec_offset), (int32_t*)(e->ws + L.meta_offset), e->cfg.n_trees, L.slots_per_tree,
        L.n_actions, e->cfg.seed, e->cfg.tree_id_base, scripted_pos, best_dev, best_action_dev);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace isla
