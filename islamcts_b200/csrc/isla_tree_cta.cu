// CTA-per-tree general MCTS engine: vanilla / APW / SPW / DPW, leaf-parallel rollout batches, VOO.
// Serves the single-tree configurations (BASELINE C1-C4) and any agent the thread-per-tree kernel
// does not cover.
//
// Reference semantics followed (paths under islaMcts/):
//   state node   vanilla/spw  agents/mcts.py:46-67, mcts_state_progressive_widening_hash.py:47-67
//                apw          agents/mcts_action_progressive_widening_hash.py:52-84  (cap test :62)
//                dpw          agents/mcts_double_progressive_widening_hash.py:49-75  (cap test :56)
//   action node  vanilla/apw  agents/mcts_hash.py:87-131, mcts_action_...:89-136
//                spw          agents/mcts_state_...:72-130   (cap :97, inverse-visit resampling :116-119)
//                dpw          agents/mcts_double_...:79-135  (cap :87, resampling :119-129)
//   ucb1         utils/action_selection_functions.py:9-25 ;  voo  :139-205
//   rollout      agents/abstract_mcts.py:72-91 ;  end of fit()  agents/mcts_hash.py:42-50
//
// Execution model (B200-first, not the reference's recursion):
//   * one CTA owns one tree; its node pools (SoA, indices not pointers) live in HBM and stay
//     L2-resident for single-tree work;
//   * warp 0 runs the sequential select/expand/backup logic with REPLICATED scalar state (every lane
//     holds the same cursors and RNG stream, lane 0 does the writes) and uses its 32 lanes for the wide
//     scans: UCB arg-max over up to 10^4 progressive-widening children, action de-duplication, VOO
//     Voronoi tests -- shuffle-reduced, children visited in insertion order so tie-breaks match;
//   * ALL threads of the CTA play the leaf's R rollouts (one trajectory per thread and pass, env state
//     in registers, own Philox stream (seed, sim, tree, j)); returns are reduced in a fixed
//     power-of-two tree in shared memory, so results do not depend on the CTA size;
//   * the descent records (state, action, reward) per level; backup walks it backwards.
// Envs on the device are deterministic, hence an action node has at most ONE child state;
// the dict-overwrite quirks of spw/dpw (a13, a14) are reproduced on that single slot.
#include <cmath>
#include <cooperative_groups.h>

#include "isla_envs.cuh"
#include "isla_internal.h"
#include "isla_philox.cuh"

namespace isla {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxRollSmem = 4096;  // doubles of shared memory for the rollout reduction

struct Pool {
    // state nodes
    double* s_total; double* s_term_reward; int* s_ns; int* s_flags; int* s_coff; int* s_ccap; int* s_ccnt;
    void* s_state;
    // action nodes
    double* a_total; double* a_value; double* a_value2; int* a_na; int* a_visits; int* a_child;
    int* s_next;   // next entry of the parent action node's children dict (insertion order), -1 = last; a_child = the first
    // child index arrays, path
    int* ipool; int* p_s; int* p_a; double* p_r; int* p_flag; double* p_g;
    int* meta;
};

struct CtaParams {
    char* pool_base; long long pool_stride;
    long long s_cap, a_cap, i_cap;
    long long off[24];
    unsigned long long* counters;
    const uint8_t* trace_kinds; const double* trace_values; long long trace_len;
    long long n_trees, scripted_tree;
    unsigned long long seed, tree_id_base;
    double C, gamma, alpha1, k1, alpha2, k2, epsilon;
    const double* disc_table; const double* cum_table; const double* ln_table; const double* pw1_table; const double* pw2_table;
    int ln_count, disc_count;
    int n_sim, max_depth, budget_mode, agent, n_actions, R, expansion, voo_n, voo_max_try, real_bytes;
    int pool_in_smem;  // the tree pool (a compact view sized for this launch) fits in shared memory: search there, copy back at the end
    long long off_s[24]; long long stride_s;   // array offsets / bytes of that compact view (capacity cap_s nodes)
    int state_bytes;   // bytes of one env state in the pool
    char* hash_base; long long hash_cap;  // per-tree action hash set (apw/dpw): hash_cap int4 entries per tree, 0 = none
    ActionGrid grid;                       // DiscreteCurveEnv action grid (g0 == 0: not a grid)
    int action_dim;                        // 1, or 2 for the continuous CurveEnv box (second coordinate in a_value2)
    int cluster;                           // CTAs (one thread-block cluster) sharing one tree's leaf-parallel rollouts
    const double* heur;                    // grid_policy table (FrozenLake rollouts), nullptr = random rollouts
    const double* noise; int noise_n;      // slippery FrozenLake: noise[k] of the k-th env.step of a simulation
    int tree_parallel;                     // > 1: that many simulations of ONE tree in flight (virtual loss); else sequential
    double virtual_value;                  // the return a leaf is provisionally credited with while its rollouts are in flight
};

enum : int { O_S_TOTAL = 0, O_S_TERMR, O_S_NS, O_S_FLAGS, O_S_COFF, O_S_CCAP, O_S_CCNT, O_S_STATE, O_A_TOTAL, O_A_VALUE,
             O_A_NA, O_A_VISITS, O_A_CHILD, O_IPOOL, O_P_S, O_P_A, O_P_R, O_P_FLAG, O_META, O_P_G, O_A_VALUE2, O_S_NEXT, O_COUNT };
enum : int { CM_NS = 0, CM_NA = 1, CM_ITOP = 2, CM_SIMS = 3, CM_RNG = 4, CM_FAULT = 5, CM_TRACE_POS = 6, CM_NOISE_POS = 7, CM_WORDS = 8 };

__device__ __forceinline__ Pool make_pool_off(const long long* off, char* b) {
    Pool q;
    q.s_total = (double*)(b + off[O_S_TOTAL]); q.s_term_reward = (double*)(b + off[O_S_TERMR]);
    q.s_ns = (int*)(b + off[O_S_NS]); q.s_flags = (int*)(b + off[O_S_FLAGS]);
    q.s_coff = (int*)(b + off[O_S_COFF]); q.s_ccap = (int*)(b + off[O_S_CCAP]); q.s_ccnt = (int*)(b + off[O_S_CCNT]);
    q.s_state = (void*)(b + off[O_S_STATE]);
    q.a_total = (double*)(b + off[O_A_TOTAL]); q.a_value = (double*)(b + off[O_A_VALUE]);
    q.a_na = (int*)(b + off[O_A_NA]); q.a_visits = (int*)(b + off[O_A_VISITS]); q.a_child = (int*)(b + off[O_A_CHILD]);
    q.ipool = (int*)(b + off[O_IPOOL]);
    q.p_s = (int*)(b + off[O_P_S]); q.p_a = (int*)(b + off[O_P_A]); q.p_r = (double*)(b + off[O_P_R]);
    q.p_flag = (int*)(b + off[O_P_FLAG]);
    q.meta = (int*)(b + off[O_META]);
    q.p_g = (double*)(b + off[O_P_G]);
    q.a_value2 = (double*)(b + off[O_A_VALUE2]);
    q.s_next = (int*)(b + off[O_S_NEXT]);
    return q;
}
__device__ __forceinline__ Pool make_pool_at(const CtaParams& p, char* b) { return make_pool_off(p.off, b); }
// Copies the USED part of one tree's pool between its HBM layout (capacity s_cap) and the compact shared-memory view
// (capacity of this launch): nS state entries, nA action entries, itop child indices and the meta words.  The path
// arrays are scratch of a launch and are not copied.
__device__ __forceinline__ void copy_pool(const CtaParams& p, char* dst, const long long* doff, const char* src, const long long* soff,
                                          int nS, int nA, int itop, int tid, int nt) {
    auto cp = [&](int idx, long long bytes) {
        const int4* s4 = reinterpret_cast<const int4*>(src + soff[idx]);
        int4* d4 = reinterpret_cast<int4*>(dst + doff[idx]);
        const int n16 = (int)((bytes + 15) >> 4);           // arrays start 256-byte aligned and are padded to 256
        for (int i = tid; i < n16; i += nt) d4[i] = s4[i];
    };
    cp(O_S_TOTAL, nS * 8ll); cp(O_S_TERMR, nS * 8ll); cp(O_S_NS, nS * 4ll); cp(O_S_FLAGS, nS * 4ll);
    cp(O_S_COFF, nS * 4ll); cp(O_S_CCAP, nS * 4ll); cp(O_S_CCNT, nS * 4ll); cp(O_S_STATE, (long long)nS * p.state_bytes);
    cp(O_S_NEXT, nS * 4ll);
    cp(O_A_TOTAL, nA * 8ll); cp(O_A_VALUE, nA * 8ll); cp(O_A_NA, nA * 4ll); cp(O_A_VISITS, nA * 4ll); cp(O_A_CHILD, nA * 4ll);
    if (p.action_dim == 2) cp(O_A_VALUE2, nA * 8ll);
    cp(O_IPOOL, itop * 4ll);
    cp(O_META, CM_WORDS * 4ll);
}
__device__ __forceinline__ Pool make_pool(const CtaParams& p, long long tree) {
    return make_pool_at(p, p.pool_base + tree * p.pool_stride);
}

struct Cursor {  // replicated in every lane of warp 0
    const uint8_t* k; const double* v; long long len, pos; int fault;
    __device__ __forceinline__ double take(int kind) {
        if (fault) return 0.0;
        if (pos >= len || k[pos] != kind) { fault = ISLA_ERR_TRACE; return 0.0; }
        return v[pos++];
    }
};

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(kFull, x, o));
    return x;
}
// ---------------------------------------------------------------- CTA-wide evaluation of voo's rejection tries
// voo draws uniform samples until one falls into the Voronoi cell of the best action (up to max_try tries, O(children)
// each).  The tries of one sample are independent given their position in the counter-based tree stream, so in
// free-running mode every thread of the CTA evaluates one try and the FIRST try (in stream order) that ends the
// reference's loop wins -- the result and the stream position afterwards are exactly the serial ones.
struct VooWork {
    int cmd;            // VOO_EVAL: evaluate a batch of tries; VOO_FIND: locate the closest tried sample; VOO_EXIT: descent over
    int s, best, base, first;
    unsigned c0;        // tree-stream words consumed before try 0 of this sample
    int ties;           // VOO_FIND: number of tries at distance tmin
    double abest, abest2;
    unsigned long long tmin_bits;   // smallest distance to the best action over the tries so far (bits of a double >= 0)
};
enum : int { VOO_EVAL = 0, VOO_EXIT = 1, VOO_FIND = 2 };

struct RollShared {
    double leaf[8];
    int budget, need, exit_flag;
    int depth;          // transition-noise index of the leaf's first rollout step (slippery FrozenLake): the env.step calls of
                        // the simulation so far (cloned env), or of the whole fit() so far (spw / dpw: live env)
    int roll_steps;     // steps rollout 0 of the leaf played (advances the live env's noise index)
    int roll_next;      // this CTA's next unclaimed rollout of the current leaf (0 between rollout phases)
    unsigned sim;
    double result;
    VooWork vw;
};

// Tree-parallel mode (opt-in build extension; north_star subsystem 2: "virtual loss for leaf-parallel batches"): the
// record of one in-flight simulation between its descent and the arrival of its rollouts.
constexpr int kTpMax = 64;
struct TpRec {
    double leaf[8];
    double r_leaf, g_prov;
    int budget, need, depth, leaf_a, leaf_s, level, pb;
    unsigned sim;
};

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Box bounds per action dimension widened to double; gym boxes are float32 (returns true), CurveEnv's box is float64
template <int ENV>
__device__ __forceinline__ bool env_box_bounds(double (&lo)[2], double (&hi)[2]) {
    if constexpr (ENV == ENV_CURVE) { lo[0] = -5.0; hi[0] = 5.0; lo[1] = -30.0; hi[1] = 30.0; return false; }
    float lof, hif; env_action_bounds(ENV, lof, hif);
    lo[0] = (double)lof; hi[0] = (double)hif; lo[1] = hi[1] = 0.0;
    return true;
}

// word w of a tree stream (random access: the stream is counter based)
__device__ __forceinline__ uint32_t tree_stream_word(unsigned long long seed, uint32_t tree_id, uint32_t w) {
    uint32_t buf[4];
    philox4x32_10(w >> 2, 0u, tree_id, kStreamTree, (uint32_t)seed, (uint32_t)(seed >> 32), buf);
    const uint32_t r = w & 3u;
    return r == 0 ? buf[0] : r == 1 ? buf[1] : r == 2 ? buf[2] : buf[3];
}
// np.random.uniform(low, high, shape) of try j of the sample whose tries start at stream word c0
template <int ENV>
__device__ __forceinline__ void voo_sample_at(const CtaParams& p, uint32_t tree_id, unsigned c0, int j, double& x, double& x2) {
    double lo[2], hi[2];
    env_box_bounds<ENV>(lo, hi);
    const uint32_t w = c0 + (uint32_t)j * (uint32_t)p.action_dim;
    x = __dadd_rn(lo[0], __dmul_rn(__dsub_rn(hi[0], lo[0]), u32_to_d01(tree_stream_word(p.seed, tree_id, w))));
    x2 = p.action_dim == 2 ? __dadd_rn(lo[1], __dmul_rn(__dsub_rn(hi[1], lo[1]), u32_to_d01(tree_stream_word(p.seed, tree_id, w + 1)))) : 0.0;
}
// try j = w.base + tid: does it end the loop of :160-190 (coincides with the best action, or is at least as close to
// it as to every other action)?  Distances like scipy's cdist: |d| in one dimension, sqrt(d0^2 + d1^2) in two -- the
// square roots are only taken when the squared distances are too close to order them safely.
template <int ENV>
__device__ __forceinline__ void voo_eval_tries(const CtaParams& p, const Pool& q, VooWork& w, uint32_t tree_id, int tid) {
    const int j = w.base + tid;
    const bool live = j < p.voo_max_try;
    double x = 0.0, x2 = 0.0;
    if (live) voo_sample_at<ENV>(p, tree_id, w.c0, j, x, x2);
    {   // closest tried sample so far (needed only if all max_try tries fail): one shared atomic per warp
        const double d0 = __dsub_rn(x, w.abest), d1 = __dsub_rn(x2, w.abest2);
        double db = p.action_dim != 2 ? fabs(d0) : __dsqrt_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)));
        if (!live) db = INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) db = fmin(db, __shfl_xor_sync(kFull, db, o));
        if ((tid & 31) == 0) atomicMin(&w.tmin_bits, (unsigned long long)__double_as_longlong(db));
    }
    if (!live || j > *(volatile int*)&w.first) return;
    const int off = q.s_coff[w.s], n = q.s_ccnt[w.s];
    bool stop = true;
    if (p.action_dim != 2) {
        const double db = fabs(__dsub_rn(x, w.abest));
        if (db != 0.0) {
            for (int i = 0; i < n; ++i) {
                if (i == w.best) continue;
                if (!(db <= fabs(__dsub_rn(x, q.a_value[q.ipool[off + i]])))) { stop = false; break; }
            }
        }
    } else {
        const double d0 = __dsub_rn(x, w.abest), d1 = __dsub_rn(x2, w.abest2);
        const double sb = __dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1));
        if (sb != 0.0) {
            const double db = __dsqrt_rn(sb);
            for (int i = 0; i < n; ++i) {
                if (i == w.best) continue;
                const int a = q.ipool[off + i];
                const double e0 = __dsub_rn(x, q.a_value[a]), e1 = __dsub_rn(x2, q.a_value2[a]);
                const double si = __dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1));
                if (sb <= si) continue;                                   // sqrt is monotone: db <= d_i
                if (sb > si * (1.0 + 1e-15) || !(db <= __dsqrt_rn(si))) { stop = false; break; }
            }
        }
    }
    if (stop) atomicMin(&w.first, j);
}
// all max_try tries failed: which tries lie at distance tmin from the best action (thread tid looks at tries tid, tid+nt, ..)
template <int ENV>
__device__ __forceinline__ void voo_find_closest(const CtaParams& p, VooWork& w, uint32_t tree_id, int tid, int nt) {
    const double tmin = __longlong_as_double((long long)w.tmin_bits);
    for (int j = tid; j < p.voo_max_try; j += nt) {
        double x, x2;
        voo_sample_at<ENV>(p, tree_id, w.c0, j, x, x2);
        const double d0 = __dsub_rn(x, w.abest), d1 = __dsub_rn(x2, w.abest2);
        const double db = p.action_dim != 2 ? fabs(d0) : __dsqrt_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)));
        if (db == tmin) { atomicAdd(&w.ties, 1); atomicMin(&w.first, j); }
    }
}

template <int ENV, typename T, bool SCRIPTED>
struct Search {
    static constexpr int S = EnvTraits<ENV>::S;
    const CtaParams& p;
    Pool q;
    int lane;
    int nS, nA, itop, fault;
    int noise_pos;   // spw / dpw step the LIVE env: index of the next transition-noise draw, running through the whole fit()
    Stream ts;
    Cursor tc;
    uint32_t tree_id;
    char* hash;  // this tree's action hash set (apw/dpw), HBM
    unsigned long long c_env = 0, c_roll = 0, c_levels = 0, c_nodes = 0;

    __device__ Search(const CtaParams& p_, long long tree, char* pool_bytes, bool compact) : p(p_), q(make_pool_off(compact ? p_.off_s : p_.off, pool_bytes)) {
        hash = p.hash_base + tree * p.hash_cap * 16;
        lane = threadIdx.x & 31;
        nS = q.meta[CM_NS]; nA = q.meta[CM_NA]; itop = q.meta[CM_ITOP]; fault = q.meta[CM_FAULT];
        noise_pos = q.meta[CM_NOISE_POS];
        tree_id = (uint32_t)(p.tree_id_base + (unsigned long long)tree);
        ts.resume(p.seed, 0u, tree_id, kStreamTree, (uint32_t)q.meta[CM_RNG]);
        tc = Cursor{p.trace_kinds, p.trace_values, p.trace_len, 0, 0};
    }

    // ---------------------------------------------------------------- randomness (replicated)
    __device__ int choice(int n) {
        if constexpr (SCRIPTED) {
            const int v = (int)tc.take(ISLA_K_CHOICE);
            if (v < 0 || v >= n) { if (!tc.fault) tc.fault = ISLA_ERR_TRACE; return 0; }
            return v;
        } else {
            return n > 1 ? (int)mulhi_u32(ts.next(), (uint32_t)n) : 0;
        }
    }
    // env.action_space.sample() drawn from the tree stream; a1 = second coordinate (continuous CurveEnv), else 0
    __device__ double tree_action(double& a1) {
        a1 = 0.0;
        if constexpr (ENV == ENV_CURVE) {
            if (p.grid.g0 > 0) {
                const int nd = p.grid.g0 * p.grid.g1;
                if constexpr (SCRIPTED) {
                    const int v = (int)tc.take(ISLA_K_SAMPLE_D);
                    if (v < 0 || v >= nd) { if (!tc.fault) tc.fault = ISLA_ERR_TRACE; return 0.0; }
                    return (double)v;
                }
                return (double)mulhi_u32(ts.next(), (uint32_t)nd);
            }
            // Box([-5, -30], [5, 30], dtype=float64): one uniform per dimension (curve_env.py:90-95)
            if constexpr (SCRIPTED) { const double a0 = tc.take(ISLA_K_SAMPLE_C); a1 = tc.take(ISLA_K_SAMPLE_C); return a0; }
            const double a0 = __dadd_rn(-5.0, __dmul_rn(10.0, u32_to_d01(ts.next())));
            a1 = __dadd_rn(-30.0, __dmul_rn(60.0, u32_to_d01(ts.next())));
            return a0;
        } else if constexpr (EnvTraits<ENV>::kContinuous) {
            if constexpr (SCRIPTED) return (double)(float)tc.take(ISLA_K_SAMPLE_C);
            float lo, hi; env_action_bounds(ENV, lo, hi);
            return (double)box_sample_f32(ts.next(), lo, hi);
        } else {
            if constexpr (SCRIPTED) return (double)(int)tc.take(ISLA_K_SAMPLE_D);
            return (double)mulhi_u32(ts.next(), (uint32_t)EnvTraits<ENV>::A);
        }
    }
    __device__ double u01() {
        if constexpr (SCRIPTED) return tc.take(ISLA_K_U01);
        return u32_to_d01(ts.next());
    }
    __device__ double uniform(double lo, double hi) {
        if constexpr (SCRIPTED) return tc.take(ISLA_K_UNIFORM);
        return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u32_to_d01(ts.next())));
    }

    // ---------------------------------------------------------------- node pools
    __device__ T* state_ptr(int s) { return reinterpret_cast<T*>(q.s_state) + (long long)s * S; }

    __device__ int new_state(const T (&st)[S], bool terminal, double term_reward) {
        if (nS >= p.s_cap) { fault = ISLA_ERR_CAPACITY; return 0; }
        const int s = nS++;
        if (lane == 0) {
            q.s_total[s] = 0.0; q.s_term_reward[s] = term_reward; q.s_ns[s] = 0; q.s_flags[s] = terminal ? 1 : 0;
            q.s_coff[s] = 0; q.s_ccap[s] = 0; q.s_ccnt[s] = 0; q.s_next[s] = -1;
            T* d = state_ptr(s);
#pragma unroll
            for (int i = 0; i < S; ++i) {
                T v = st[i];
                // spw/dpw keep node.data = the float32 observation and poke it back into the env
                if (p.agent == ISLA_AGENT_SPW || p.agent == ISLA_AGENT_DPW) v = T(float(v));
                d[i] = v;
            }
        }
        ++c_nodes;
        return s;
    }
    __device__ int new_action(double value, double value2 = 0.0) {
        if (nA >= p.a_cap) { fault = ISLA_ERR_CAPACITY; return 0; }
        const int a = nA++;
        if (lane == 0) {
            q.a_total[a] = 0.0; q.a_value[a] = value; q.a_na[a] = 0; q.a_visits[a] = 0; q.a_child[a] = -1;
            if (p.action_dim == 2) q.a_value2[a] = value2;
        }
        return a;
    }
    // ---- the children dict of an action node (spw / dpw on a stochastic env: several next states per action), a singly
    // linked list in insertion order: a_child = first, s_next = next.  Keys are the observations, compared bit for bit
    // (`observation.tobytes()`, mcts_state_...:79); spw / dpw store the float32 observation as node.data.
    __device__ int find_child(int a, const T (&st)[S], int& prev) const {
        prev = -1;
        for (int c = q.a_child[a]; c >= 0; prev = c, c = q.s_next[c]) {
            const T* d = reinterpret_cast<const T*>(q.s_state) + (long long)c * S;
            bool same = true;
#pragma unroll
            for (int i = 0; i < S; ++i) same &= (d[i] == T(float(st[i])));
            if (same) return c;
        }
        return -1;
    }
    __device__ int count_children(int a) const {
        int n = 0;
        for (int c = q.a_child[a]; c >= 0; c = q.s_next[c]) ++n;
        return n;
    }
    // children[obs_bytes] = node: an existing key keeps its place in the dict, a new key goes to the end
    __device__ void put_child(int a, int old, int prev, int neu) {
        if (lane == 0) {
            if (old >= 0) {
                q.s_next[neu] = q.s_next[old];
                if (prev < 0) q.a_child[a] = neu; else q.s_next[prev] = neu;
            } else if (prev < 0) q.a_child[a] = neu;
            else q.s_next[prev] = neu;
        }
        __syncwarp();
    }

    // append an action node to a state's child index array (insertion order), growing it by doubling
    __device__ void append_action(int s, int a) {
        int off = q.s_coff[s], cap = q.s_ccap[s], cnt = q.s_ccnt[s];
        __syncwarp();
        if (cnt == cap) {
            const int ncap = cap == 0 ? 4 : cap * 2;
            if (itop + ncap > p.i_cap) { fault = ISLA_ERR_CAPACITY; return; }
            const int noff = itop;
            itop += ncap;
            for (int i = lane; i < cnt; i += 32) q.ipool[noff + i] = q.ipool[off + i];
            off = noff; cap = ncap;
            if (lane == 0) { q.s_coff[s] = off; q.s_ccap[s] = cap; }
        }
        if (lane == 0) { q.ipool[off + cnt] = a; q.s_ccnt[s] = cnt + 1; }
        __syncwarp();
    }

    // ---------------------------------------------------------------- UCB1 (float64, exact tie semantics)
    // returns the RANK (position in node.actions) of the chosen child
    __device__ int ucb_pick(int s) {
        const int off = q.s_coff[s], n = q.s_ccnt[s];
        const int ns = q.s_ns[s];
        const double lg = ns < p.ln_count ? p.ln_table[ns] : ::log((double)ns);
        auto score = [&](int i) -> double {
            const int a = q.ipool[off + i];
            const double na = (double)q.a_na[a];
            return __dadd_rn(__ddiv_rn(q.a_total[a], na), __dmul_rn(p.C, __dsqrt_rn(__ddiv_rn(lg, na))));
        };
        if (n <= 32) {
            // narrow node (vanilla, spw, young pw nodes): one score per lane, one max, one ballot
            const double mine = lane < n ? score(lane) : -INFINITY;
            const double top = warp_max(mine);
            const unsigned b = __ballot_sync(kFull, lane < n && mine == top);
            const int nt = __popc(b);
            int pos = (SCRIPTED || nt > 1) ? choice(nt) : 0;
            unsigned bb = b;
            for (int k = 0; k < pos; ++k) bb &= bb - 1;
            return __ffs(bb) - 1;
        }
        {   // wide node: fp32 screening first (MUFU approximations, error ~1e-6 of the score magnitude).  A gap above 1e-5
            // of that magnitude between the two largest scores settles the arg-max without the float64 passes.
            const float lgf = (float)lg, Cf = (float)p.C;
            float top = -INFINITY, second = -INFINITY, scale = 0.f; int arg = -1;
            for (int i = lane; i < n; i += 32) {
                const int a = q.ipool[off + i];
                float rn, ef;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"((float)q.a_na[a]));
                const float qf = (float)q.a_total[a] * rn;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ef) : "f"(lgf * rn));
                ef *= Cf;
                const float sc = qf + ef;
                scale = fmaxf(scale, fabsf(qf) + fabsf(ef));
                if (sc > top) { second = top; top = sc; arg = i; } else second = fmaxf(second, sc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ot = __shfl_xor_sync(kFull, top, o), os = __shfl_xor_sync(kFull, second, o);
                const int oa = __shfl_xor_sync(kFull, arg, o);
                scale = fmaxf(scale, __shfl_xor_sync(kFull, scale, o));
                if (ot > top) { second = fmaxf(top, os); top = ot; arg = oa; }
                else { second = fmaxf(second, ot); }      // ot <= top: an equal top leaves the gap at 0 (undecided)
            }
            if (top - second > 1e-5f * scale + 1e-30f) {
                if constexpr (SCRIPTED) (void)choice(1);   // the reference still draws among its single candidate
                return arg;
            }
        }
        double best = -INFINITY;
        for (int i = lane; i < n; i += 32) best = fmax(best, score(i));
        best = warp_max(best);
        int ties = 0, first = -1;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool tie = i < n && score(i) == best;
            const unsigned b = __ballot_sync(kFull, tie);
            if (first < 0 && b) first = base + __ffs(b) - 1;
            ties += __popc(b);
        }
        int pos = (SCRIPTED || ties > 1) ? choice(ties) : 0;
        if (pos == 0) return first;
        int seen = 0, pick = first;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool tie = i < n && score(i) == best;
            const unsigned b = __ballot_sync(kFull, tie);
            const int c = __popc(b);
            if (pos < seen + c) {
                unsigned bb = b;
                for (int k = 0; k < pos - seen; ++k) bb &= bb - 1;
                pick = base + __ffs(bb) - 1;
                break;
            }
            seen += c;
        }
        return pick;
    }

    // actions.get(action.tobytes()) through the per-tree hash set {action bits, state node} -> rank in the state's child
    // array (open addressing, linear probing; capacity is a power of two >= 2 x action capacity).  Replaces an O(n)
    // scan of the children, which dominated wide progressive-widening roots (C3: 10 000 children).
    __device__ __forceinline__ unsigned long long hash_slot(int s, long long bits) const {
        unsigned long long h = (unsigned long long)bits * 0x9E3779B97F4A7C15ull;
        h ^= (unsigned long long)(unsigned)s * 0xC2B2AE3D27D4EB4Full;
        h ^= h >> 29;
        return h & (unsigned long long)(p.hash_cap - 1);
    }
    // 2-D actions (CurveEnv box) do not fit the 64-bit key field: the entry then holds a 64-bit mix of both
    // coordinates and a match is confirmed against the stored action node.
    __device__ __forceinline__ long long key_bits(double value, double value2) const {
        long long bits = __double_as_longlong(value);
        if (p.action_dim == 2) {
            unsigned long long m = (unsigned long long)__double_as_longlong(value2) * 0xD6E8FEB86659FD93ull;
            bits ^= (long long)(m ^ (m >> 32));
        }
        return bits;
    }
    __device__ __forceinline__ bool confirm(int s, int rank, double value, double value2) const {
        if (p.action_dim != 2) return true;
        const int a = q.ipool[q.s_coff[s] + rank];
        return __double_as_longlong(q.a_value[a]) == __double_as_longlong(value) &&
               __double_as_longlong(q.a_value2[a]) == __double_as_longlong(value2);
    }
    __device__ int find_action(int s, double value, double value2 = 0.0) {
        if (p.hash_cap == 0) return find_action_scan(s, value, value2);
        const long long want = key_bits(value, value2);
        unsigned long long h = hash_slot(s, want);
        for (;;) {
            const int4 e = reinterpret_cast<const int4*>(hash)[h];   // {bits lo, bits hi, state, rank}
            if (e.z < 0) return -1;
            if (e.z == s && e.x == (int)(unsigned)(want & 0xffffffffll) && e.y == (int)(want >> 32) && confirm(s, e.w, value, value2)) return e.w;
            h = (h + 1) & (unsigned long long)(p.hash_cap - 1);
        }
    }
    // record (state, action bits) -> rank; an existing key is overwritten (dpw replaces the node under the same key)
    __device__ void hash_put(int s, double value, int rank, double value2 = 0.0) {
        if (p.hash_cap == 0) return;
        const long long bits = key_bits(value, value2);
        unsigned long long h = hash_slot(s, bits);
        for (;;) {
            const int4 e = reinterpret_cast<const int4*>(hash)[h];
            if (e.z < 0 || (e.z == s && e.x == (int)(unsigned)(bits & 0xffffffffll) && e.y == (int)(bits >> 32) && confirm(s, e.w, value, value2))) break;
            h = (h + 1) & (unsigned long long)(p.hash_cap - 1);
        }
        if (lane == 0) reinterpret_cast<int4*>(hash)[h] = make_int4((int)(unsigned)(bits & 0xffffffffll), (int)(bits >> 32), s, rank);
        __syncwarp();
    }
    __device__ int find_action_scan(int s, double value, double value2 = 0.0) {
        const int off = q.s_coff[s], n = q.s_ccnt[s];
        const long long want = __double_as_longlong(value), want2 = __double_as_longlong(value2);
        int found = -1;
        for (int base = 0; base < n && found < 0; base += 32) {
            const int i = base + lane;
            bool hit = i < n && __double_as_longlong(q.a_value[q.ipool[off + i]]) == want;
            if (hit && p.action_dim == 2) hit = __double_as_longlong(q.a_value2[q.ipool[off + i]]) == want2;
            const unsigned b = __ballot_sync(kFull, hit);
            if (b) found = base + __ffs(b) - 1;
        }
        return found;
    }

    // Box bounds per action dimension widened to double; gym boxes are float32, CurveEnv's box is float64
    __device__ __forceinline__ bool box_bounds(double (&lo)[2], double (&hi)[2]) const { return env_box_bounds<ENV>(lo, hi); }
    __device__ __forceinline__ double act2(int a) const { return p.action_dim == 2 ? q.a_value2[a] : 0.0; }
    // scipy cdist(..., 'euclidean') on float64 rows: sqrt(sum (u_i - v_i)^2); one dimension: |d|
    __device__ __forceinline__ double euclid(double x0, double x1, double y0, double y1) const {
        const double d0 = __dsub_rn(x0, y0);
        if (p.action_dim != 2) return fabs(d0);
        const double d1 = __dsub_rn(x1, y1);
        return __dsqrt_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)));
    }

    // genetic_policy2 (utils/action_selection_functions.py:104-136), one or two action dimensions.  p is drawn
    // first on every call; 0 / 1 / 2 existing actions get the centre / a random bound / the missing bound; later calls
    // take default_policy with probability 1 - epsilon, else the mean of the two LOWEST-q actions (the list is sorted
    // ascending by q and entries [0], [1] are used, :108-112; python's sort is stable, so ties keep insertion order).
    // float32 gym boxes average in float32, CurveEnv's float64 box in float64 (np.mean(..., axis=0)).
    __device__ double genetic2_expand(int s, double& out2) {
        const double pr = u01();
        const int off = q.s_coff[s], n = q.s_ccnt[s];
        double lo[2], hi[2];
        const bool f32 = box_bounds(lo, hi);
        auto mean2 = [&](double x, double y) -> double {
            return f32 ? (double)__fmul_rn(__fadd_rn((float)x, (float)y), 0.5f) : __dmul_rn(__dadd_rn(x, y), 0.5);
        };
        out2 = 0.0;
        if (n == 0) { out2 = mean2(lo[1], hi[1]); return mean2(lo[0], hi[0]); }
        if (n == 1) {
            int idx;
            if constexpr (SCRIPTED) idx = (int)tc.take(ISLA_K_CHOICES);
            else idx = (1.0 > __dmul_rn(u32_to_d01(ts.next()), 2.0)) ? 0 : 1;   // random.choices([high, low])
            out2 = idx == 0 ? hi[1] : lo[1];
            return idx == 0 ? hi[0] : lo[0];
        }
        if (n == 2) {
            const bool has_high = find_action(s, hi[0], hi[1]) >= 0;
            out2 = has_high ? lo[1] : hi[1];
            return has_high ? lo[0] : hi[0];
        }
        if (pr < __dsub_rn(1.0, p.epsilon)) return tree_action(out2);
        auto qv = [&](int i) -> double { const int a = q.ipool[off + i]; return __ddiv_rn(q.a_total[a], (double)q.a_na[a]); };
        int r0 = -1, r1 = -1;
        for (int pass = 0; pass < 2; ++pass) {
            double bq = INFINITY; int br = 0x7fffffff;
            for (int i = lane; i < n; i += 32) {
                if (pass == 1 && i == r0) continue;
                const double v = qv(i);
                if (v < bq) { bq = v; br = i; }          // ascending i per lane: the first minimum is kept
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double oq = __shfl_xor_sync(kFull, bq, o);
                const int orank = __shfl_xor_sync(kFull, br, o);
                if (oq < bq || (oq == bq && orank < br)) { bq = oq; br = orank; }
            }
            if (pass == 0) r0 = br; else r1 = br;
        }
        const int a0 = q.ipool[off + r0], a1 = q.ipool[off + r1];
        out2 = p.action_dim == 2 ? mean2(q.a_value2[a0], q.a_value2[a1]) : 0.0;
        return mean2(q.a_value[a0], q.a_value[a1]);   // np.mean([best, second], axis=0)
    }

    // genetic_policy (utils/action_selection_functions.py:50-102), one or two action dimensions.  0 / 1 / 2 existing
    // actions get the centre / a random bound / the missing bound without drawing p; later calls draw p and, if
    // p < epsilon, sample n_samples points uniformly in the box spanned by the two LOWEST-q actions (:56-59; which one is
    // the box's `low` follows from their distance to action_space.low, :61-69, measured in the box's dtype) and return the
    // sample closest to the first of them (np.argmin: the first minimum).
    __device__ double genetic_expand(int s, double& out2) {
        const int off = q.s_coff[s], n = q.s_ccnt[s];
        double lo[2], hi[2];
        const bool f32 = box_bounds(lo, hi);
        const bool two = p.action_dim == 2;
        out2 = 0.0;
        auto mean2 = [&](double x, double y) -> double {
            return f32 ? (double)__fmul_rn(__fadd_rn((float)x, (float)y), 0.5f) : __dmul_rn(__dadd_rn(x, y), 0.5);
        };
        if (n == 0) { out2 = mean2(hi[1], lo[1]); return mean2(hi[0], lo[0]); }
        if (n == 1) {
            int idx;
            if constexpr (SCRIPTED) idx = (int)tc.take(ISLA_K_CHOICES);
            else idx = (1.0 > __dmul_rn(u32_to_d01(ts.next()), 2.0)) ? 0 : 1;   // random.choices([high, low])
            out2 = idx == 0 ? hi[1] : lo[1];
            return idx == 0 ? hi[0] : lo[0];
        }
        if (n == 2) {
            const bool has_high = find_action(s, hi[0], hi[1]) >= 0;
            out2 = has_high ? lo[1] : hi[1];
            return has_high ? lo[0] : hi[0];
        }
        const double pr = u01();
        if (!(pr < p.epsilon)) return tree_action(out2);
        auto qv = [&](int i) -> double { const int a = q.ipool[off + i]; return __ddiv_rn(q.a_total[a], (double)q.a_na[a]); };
        int r0 = -1, r1 = -1;
        for (int pass = 0; pass < 2; ++pass) {
            double bq = INFINITY; int br = 0x7fffffff;
            for (int i = lane; i < n; i += 32) {
                if (pass == 1 && i == r0) continue;
                const double v = qv(i);
                if (v < bq) { bq = v; br = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double oq = __shfl_xor_sync(kFull, bq, o);
                const int orank = __shfl_xor_sync(kFull, br, o);
                if (oq < bq || (oq == bq && orank < br)) { bq = oq; br = orank; }
            }
            if (pass == 0) r0 = br; else r1 = br;
        }
        const int a0 = q.ipool[off + r0], a1 = q.ipool[off + r1];
        const double b0 = q.a_value[a0], b1 = act2(a0), s0 = q.a_value[a1], s1 = act2(a1);   // "best", "second best"
        double db, ds;
        if (f32) { db = (double)fabsf(__fsub_rn((float)lo[0], (float)b0)); ds = (double)fabsf(__fsub_rn((float)lo[0], (float)s0)); }
        else { db = euclid(lo[0], lo[1], b0, b1); ds = euclid(lo[0], lo[1], s0, s1); }
        const bool best_is_high = db >= ds;
        const double l0 = best_is_high ? s0 : b0, l1 = best_is_high ? s1 : b1, h0 = best_is_high ? b0 : s0, h1 = best_is_high ? b1 : s1;
        double dmin = INFINITY, pick = b0, pick2 = b1;
        for (int k = 0; k < p.voo_n && !tc.fault; ++k) {      // np.random.uniform(low, high, [n_samples, D])
            const double x = uniform(l0, h0);
            const double x2 = two ? uniform(l1, h1) : 0.0;
            const double d = euclid(b0, b1, x, x2);
            if (d < dmin) { dmin = d; pick = x; pick2 = x2; }
        }
        out2 = pick2;
        return pick;
    }

    // voo (utils/action_selection_functions.py:139-205), one or two action dimensions.  `scratch` (the CTA's rollout
    // reduction buffer, idle during the descent) holds the accepted samples: 3 doubles each.
    __device__ double voo_expand(int s, double& out2, double* scratch, RollShared* sh, int nt) {
        const double pr = u01();
        const int off = q.s_coff[s], n = q.s_ccnt[s];
        out2 = 0.0;
        if (pr <= p.epsilon || n == 0) return tree_action(out2);
        const bool two = p.action_dim == 2;
        auto qv = [&](int i) -> double { const int a = q.ipool[off + i]; return __ddiv_rn(q.a_total[a], (double)q.a_na[a]); };
        double qmax = -INFINITY;
        for (int i = lane; i < n; i += 32) qmax = fmax(qmax, qv(i));
        qmax = warp_max(qmax);
        int ties = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            ties += __popc(__ballot_sync(kFull, i < n && qv(i) == qmax));
        }
        int pos = choice(ties), best = 0, seen = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const unsigned b = __ballot_sync(kFull, i < n && qv(i) == qmax);
            const int c = __popc(b);
            if (pos < seen + c) { unsigned bb = b; for (int k = 0; k < pos - seen; ++k) bb &= bb - 1; best = base + __ffs(bb) - 1; break; }
            seen += c;
        }
        const int abest_node = q.ipool[off + best];
        const double abest = q.a_value[abest_node], abest2 = act2(abest_node);
        double lo[2], hi[2];
        box_bounds(lo, hi);
        int n_acc = 0;
        const int n_samples = p.voo_n;
        if constexpr (!SCRIPTED) {
            const int D = p.action_dim;
            for (int k = 0; k < n_samples; ++k) {
                const unsigned c0 = ts.consumed();
                int first = 0x7fffffff;
                for (int base = 0; base < p.voo_max_try && first == 0x7fffffff; base += nt) {
                    if (lane == 0) {
                        VooWork& w = sh->vw;
                        w.cmd = VOO_EVAL; w.s = s; w.best = best; w.base = base; w.first = 0x7fffffff; w.c0 = c0;
                        w.abest = abest; w.abest2 = abest2;
                        if (base == 0) w.tmin_bits = 0x7ff0000000000000ull;   // +inf
                    }
                    __syncwarp();
                    bar_sync(1, nt);                                  // the helper warps wait here during the descent
                    voo_eval_tries<ENV>(p, q, sh->vw, tree_id, lane);
                    bar_sync(2, nt);
                    first = *(volatile int*)&sh->vw.first;
                }
                double x, x2, d;
                if (first != 0x7fffffff) {
                    voo_sample_at<ENV>(p, tree_id, c0, first, x, x2);
                    d = euclid(x, x2, abest, abest2);
                    ts.resume(p.seed, 0u, tree_id, kStreamTree, c0 + (unsigned)(first + 1) * (unsigned)D);
                    if (d == 0.0) continue;                           // :177-178 leaves the loop WITHOUT appending
                } else {
                    // max_try tries failed (:186-190): the closest tried sample, ties broken by a draw
                    ts.resume(p.seed, 0u, tree_id, kStreamTree, c0 + (unsigned)p.voo_max_try * (unsigned)D);
                    auto dist_of = [&](int j) -> double {
                        if (j >= p.voo_max_try) return INFINITY;
                        double y, y2; voo_sample_at<ENV>(p, tree_id, c0, j, y, y2);
                        return euclid(y, y2, abest, abest2);
                    };
                    if (lane == 0) { sh->vw.cmd = VOO_FIND; sh->vw.first = 0x7fffffff; sh->vw.ties = 0; }
                    __syncwarp();
                    bar_sync(1, nt);
                    voo_find_closest<ENV>(p, sh->vw, tree_id, lane, nt);
                    bar_sync(2, nt);
                    const double tmin = __longlong_as_double((long long)sh->vw.tmin_bits);
                    const int tt = *(volatile int*)&sh->vw.ties;
                    int pp = choice(tt), pick = *(volatile int*)&sh->vw.first;
                    if (pp > 0) {   // several tries at exactly the same distance: the pp-th in stream order
                        int seen2 = 0;
                        for (int b0 = 0; b0 < p.voo_max_try; b0 += 32) {
                            const unsigned bm = __ballot_sync(kFull, dist_of(b0 + lane) == tmin);
                            const int c = __popc(bm);
                            if (pp < seen2 + c) { unsigned bb = bm; for (int z = 0; z < pp - seen2; ++z) bb &= bb - 1; pick = b0 + __ffs(bb) - 1; break; }
                            seen2 += c;
                        }
                    }
                    voo_sample_at<ENV>(p, tree_id, c0, pick, x, x2);
                    d = tmin;
                }
                if (lane == 0) { scratch[3 * n_acc] = x; scratch[3 * n_acc + 1] = x2; scratch[3 * n_acc + 2] = d; }
                ++n_acc;
            }
        }
        for (int k = 0; SCRIPTED && k < n_samples && !tc.fault; ++k) {
            int n_try = 0;
            double tmin = INFINITY; int tmin_ties = 0;  // closest tried sample (max_try fallback)
            // the fallback needs the pos-th closest-tried sample: replay it from a saved stream position
            Stream saved = ts; const long long saved_pos = tc.pos;
            for (;;) {
                const double x = uniform(lo[0], hi[0]);
                const double x2 = two ? uniform(lo[1], hi[1]) : 0.0;
                const double db = euclid(x, x2, abest, abest2);
                if (db < tmin) { tmin = db; tmin_ties = 1; } else if (db == tmin) { ++tmin_ties; }
                if (db == 0.0) break;  // :177-178 leaves the loop WITHOUT appending
                bool ok = true;
                for (int base = 0; base < n; base += 32) {
                    const int i = base + lane;
                    bool bad = false;
                    if (i < n && i != best) { const int a = q.ipool[off + i]; bad = !(db <= euclid(x, x2, q.a_value[a], act2(a))); }
                    if (__any_sync(kFull, bad)) { ok = false; break; }
                }
                double acc_x = x, acc_x2 = x2, acc_d = db;
                bool accept = ok;
                if (!ok) {
                    ++n_try;
                    if (n_try >= p.voo_max_try) {
                        int pp = choice(tmin_ties);
                        // replay the tried samples to find the pp-th one at distance tmin
                        Stream after = ts; const long long after_pos = tc.pos;
                        ts = saved; tc.pos = saved_pos;
                        for (int j = 0; j < n_try; ++j) {
                            const double xj = uniform(lo[0], hi[0]);
                            const double xj2 = two ? uniform(lo[1], hi[1]) : 0.0;
                            if (euclid(xj, xj2, abest, abest2) == tmin) { if (pp == 0) { acc_x = xj; acc_x2 = xj2; } --pp; }
                        }
                        ts = after; tc.pos = after_pos;
                        acc_d = tmin;
                        accept = true;
                    }
                }
                if (accept) {
                    if (lane == 0) { scratch[3 * n_acc] = acc_x; scratch[3 * n_acc + 1] = acc_x2; scratch[3 * n_acc + 2] = acc_d; }
                    ++n_acc;
                    break;
                }
            }
        }
        __syncwarp();
        // np.random.choice(np.where(d == d.min())[0]) over the accepted samples
        double out = abest, out_d = INFINITY;
        out2 = abest2;
        int out_ties = 0;
        for (int i = 0; i < n_acc; ++i) { const double d = scratch[3 * i + 2]; if (d < out_d) { out_d = d; out_ties = 1; } else if (d == out_d) ++out_ties; }
        if (n_acc > 0) {
            int pp = choice(out_ties);
            for (int i = 0; i < n_acc; ++i) if (scratch[3 * i + 2] == out_d) { if (pp == 0) { out = scratch[3 * i]; out2 = scratch[3 * i + 1]; } --pp; }
        }
        __syncwarp();
        return out;
    }
};

template <int ENV, typename T, bool SCRIPTED>
__global__ void __launch_bounds__(512) cta_search_kernel(const CtaParams p) {
    constexpr int S = EnvTraits<ENV>::S;
    __shared__ RollShared sh;
    __shared__ double red[kMaxRollSmem];
    __shared__ double partials[16];  // rank 0: the rollout sums of the cluster's CTAs
    // A tree whose leaves are evaluated by thousands of rollouts is served by a thread-block CLUSTER: rank 0 owns the
    // tree (descent, expansion, backup), every rank plays an interleaved power-of-two slice of the rollouts from the leaf
    // state it reads out of rank 0's shared memory (DSMEM), reduces it with the fixed pairwise tree and stores its sum
    // into rank 0's `partials`; rank 0 finishes the same tree over the ranks -- the mean is bit-identical for any
    // cluster and CTA size.  Two cluster barriers per simulation.
    namespace cg = cooperative_groups;
    const int CS = SCRIPTED ? 1 : p.cluster;
    const unsigned crank = CS > 1 ? cg::this_cluster().block_rank() : 0u;
    const bool owner = crank == 0;
    const long long tree = SCRIPTED ? p.scripted_tree : (long long)(blockIdx.x / (unsigned)CS);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const bool leader = owner && tid < 32;  // warp 0 of the owning CTA
    const uint32_t tree_id = (uint32_t)(p.tree_id_base + (unsigned long long)tree);
    const int R = p.R;
    int P = 1;
    while (P < R) P <<= 1;

    // Small trees (C1/C2-sized: <= ~1500 simulations) live in SHARED memory for the whole launch: the pool is copied
    // in, every select/expand/backup access costs ~30 cycles instead of an L2 round trip, and it is copied back at
    // the end.  Larger pools stay in HBM/L2.
    extern __shared__ int4 pool_smem[];
    char* const pool_global = p.pool_base + tree * p.pool_stride;
    if (p.pool_in_smem && owner) {
        const int* gm = reinterpret_cast<const int*>(pool_global + p.off[O_META]);
        copy_pool(p, reinterpret_cast<char*>(pool_smem), p.off_s, pool_global, p.off, gm[CM_NS], gm[CM_NA], gm[CM_ITOP], tid, nt);
        __syncthreads();
    }
    Search<ENV, T, SCRIPTED> ctx(p, tree, (p.pool_in_smem && owner) ? reinterpret_cast<char*>(pool_smem) : pool_global,
                                 p.pool_in_smem && owner);
    const RollShared* const sh0 = CS > 1 ? cg::this_cluster().map_shared_rank(&sh, 0) : &sh;   // the owner's leaf record
    uint32_t sims_done = (uint32_t)ctx.q.meta[CM_SIMS];
    int s_done = 0;
    if (tid == 0) sh.roll_next = 0;   // barriers follow before the first rollout phase reads it
    unsigned long long my_steps = 0;  // rollout env steps played by THIS thread
    int prev_levels = 0;              // recorded path entries of the previous simulation (warp 0, replicated)

    // Widening wave (APW with random expansion, one rollout per leaf).  Whether a simulation widens the ROOT depends on
    // counts only -- len(actions) <= ceil(k * ns**alpha), mcts_action_progressive_widening_hash.py:62 -- and a simulation
    // that widens the root is: sample an action (tree stream, one draw), step the root state, create the child, roll out,
    // back the return up into the new action node, the new state node and the root.  Consecutive such simulations touch
    // disjoint nodes and consume consecutive draws, so a run of them is executed by the whole CTA at once, one thread per
    // simulation, and applied in simulation order: node indices, stream positions, the root's running total (added in
    // order by one lane) and every statistic are exactly those of the serial loop.  The run ends before the first
    // simulation that would not widen, whose sampled action already exists (the reference then re-uses that node and
    // descends), or whose first step terminates; those take the serial path.  BASELINE config C3 (k = 60, alpha = 0.8)
    // widens the root on every one of its 10 000 simulations.
    const bool wave_ok = !SCRIPTED && p.agent == ISLA_AGENT_APW && p.expansion == ISLA_EXPAND_RANDOM && R == 1 && CS == 1 &&
                         p.hash_cap > 0 && nt >= 64 && p.tree_parallel <= 1;
    __shared__ int wv_n, wv_ns0, wv_n0, wv_nS0, wv_nA0, wv_itop, wv_rng0, wv_bad, wv_off, wv_cap;
    __shared__ unsigned wv_sims0;
    constexpr int kWaveWords = ENV == ENV_CURVE ? 2 : 1;   // stream words per sampled action

    // Tree-parallel mode: up to B simulations are selected one after the other -- each descent ends with a PROVISIONAL
    // backup that credits its leaf with `virtual_value` (the virtual loss: visit counts move at once, so the next descent
    // sees the path as taken and explores elsewhere) -- then the rollouts of all B leaves run concurrently on the CTA's
    // threads, and the provisional credits are corrected to the real returns in simulation order.  Statistics differ from
    // the sequential search (simulations of a batch do not see each other's returns); the default (B <= 1) is sequential.
    const int tpB = SCRIPTED ? 1 : max(1, min(p.tree_parallel, kTpMax));
    const bool tp = tpB > 1;
    __shared__ TpRec tp_rec[kTpMax];
    int tp_b = 0, tp_pb = 0;   // simulations in flight, path entries they occupy

#ifdef ISLA_CTA_TIMING
    long long tm_verify = 0, tm_descent = 0, tm_roll = 0, tm_backup = 0, tm_other = 0, tm_mark = clock64();
    long long tm_r_sync1 = 0, tm_r_own = 0, tm_r_wait = 0, tm_r_reduce = 0;   // parts of tm_roll (thread 0 of the owner)
#define CTA_TM(acc) do { const long long now_ = clock64(); acc += now_ - tm_mark; tm_mark = now_; } while (0)
#else
#define CTA_TM(acc)
#endif
    for (int sim = 0; sim < p.n_sim; ++sim) {
        if (wave_ok) {
            Pool& q = ctx.q;
            if (tid == 0) {
                int W = 0;
                if (!ctx.fault) {
                    const int ns0 = q.s_ns[0], n0 = q.s_ccnt[0];
                    auto widens = [&](int ns_, int n_) {
                        const double cap = ns_ < p.ln_count ? p.pw1_table[ns_] : ceil(__dmul_rn(p.k1, pow((double)ns_, p.alpha1)));
                        return n_ == 0 || (double)n_ <= cap;
                    };
                    if (widens(ns0, n0) && widens(ns0 + 1, n0 + 1)) {
                        long long w = nt;
                        w = min(w, (long long)(p.n_sim - sim));
                        w = min(w, p.s_cap - ctx.nS);
                        w = min(w, p.a_cap - ctx.nA);
                        W = (int)max(w, 0ll);
                    }
                    wv_ns0 = ns0; wv_n0 = n0; wv_nS0 = ctx.nS; wv_nA0 = ctx.nA; wv_itop = ctx.itop;
                    wv_rng0 = (int)ctx.ts.consumed(); wv_sims0 = sims_done;
                }
                wv_n = W; wv_bad = W;
            }
            __syncthreads();
            const int W = wv_n;
            if (W >= 2) {
                double* const w_val = red;              // [W] sampled action (first coordinate)
                double* const w_val2 = red + 512;       // [W] second coordinate (CurveEnv box)
                double* const w_g = red + 1024;         // [W] return backed up by simulation i
                const int ns0 = wv_ns0, n0 = wv_n0;
                double val = 0.0, val2 = 0.0, r0 = 0.0;
                T st[S];
                bool bad = false;
                if (tid < W) {
                    const int ns_i = ns0 + tid, n_i = n0 + tid;
                    const double cap = ns_i < p.ln_count ? p.pw1_table[ns_i] : ceil(__dmul_rn(p.k1, pow((double)ns_i, p.alpha1)));
                    bad = !(n_i == 0 || (double)n_i <= cap);
                    // env.action_space.sample() of simulation i: draw number rng0 + i (x kWaveWords) of the tree stream
                    Stream as;
                    as.resume(p.seed, 0u, tree_id, kStreamTree, (uint32_t)(wv_rng0 + tid * kWaveWords));
                    if constexpr (ENV == ENV_CURVE) {
                        val = __dadd_rn(-5.0, __dmul_rn(10.0, u32_to_d01(as.next())));
                        val2 = __dadd_rn(-30.0, __dmul_rn(60.0, u32_to_d01(as.next())));
                    } else {
                        float lo, hi; env_action_bounds(ENV, lo, hi);
                        val = (double)box_sample_f32(as.next(), lo, hi);
                    }
                    w_val[tid] = val; w_val2[tid] = val2;
                    // first step from the root state
                    const T* sp = ctx.state_ptr(0);
#pragma unroll
                    for (int i = 0; i < S; ++i) st[i] = sp[i];
                    T rT;
                    const bool term = env_step_action<ENV, T>(st, val, val2, p.grid, rT);
                    r0 = (double)rT;
                    bad |= term;
                    // the action already exists at the root (the reference re-uses the node)?  -- per-thread probe
                    if (!bad) {
                        const long long want = ctx.key_bits(val, val2);
                        unsigned long long h = ctx.hash_slot(0, want);
                        for (;;) {
                            const int4 e = reinterpret_cast<const int4*>(ctx.hash)[h];
                            if (e.z < 0) break;
                            if (e.z == 0 && e.x == (int)(unsigned)(want & 0xffffffffll) && e.y == (int)(want >> 32) && ctx.confirm(0, e.w, val, val2)) { bad = true; break; }
                            h = (h + 1) & (unsigned long long)(p.hash_cap - 1);
                        }
                    }
                }
                __syncthreads();
                if (tid < W && !bad) {   // ... or is sampled twice within the run?
                    const long long b0 = __double_as_longlong(val), b1 = __double_as_longlong(val2);
                    for (int j = 0; j < tid; ++j)
                        if (__double_as_longlong(w_val[j]) == b0 && __double_as_longlong(w_val2[j]) == b1) { bad = true; break; }
                }
                if (tid < W && bad) atomicMin(&wv_bad, tid);
                __syncthreads();
                const int We = wv_bad;   // simulations sim .. sim + We - 1 widen the root with fresh actions
                if (We >= 2) {
                    const int nS0 = wv_nS0, nA0 = wv_nA0;
                    // the root's child-index array grows once (insertion order preserved)
                    if (tid == 0) {
                        int off = q.s_coff[0], cap = q.s_ccap[0];
                        int ncap = cap;
                        while (ncap < n0 + We) ncap = ncap == 0 ? 4 : ncap * 2;
                        wv_off = off; wv_cap = cap;
                        if (ncap != cap) {
                            if (wv_itop + ncap > p.i_cap) { wv_cap = -1; }
                            else { wv_off = wv_itop; wv_itop += ncap; wv_cap = -ncap - 2; }   // negative: moved, copy the old entries
                        }
                    }
                    __syncthreads();
                    if (wv_cap == -1) {
                        if (leader) ctx.fault = ISLA_ERR_CAPACITY;
                    } else {
                        const int new_off = wv_off;
                        if (wv_cap < -1) {
                            const int old_off = q.s_coff[0];
                            for (int i = tid; i < n0; i += nt) q.ipool[new_off + i] = q.ipool[old_off + i];
                        }
                        if (tid < We) {
                            // simulation i: rollout from the new child (own Philox stream: cumulative simulation index)
                            const int budget = p.budget_mode == ISLA_BUDGET_ZERO ? 0 : max(0, min(p.max_depth - 1, p.disc_count));
                            T ss[S];
#pragma unroll
                            for (int i = 0; i < S; ++i) ss[i] = st[i];
                            int t = 0;
                            T rr = T(0);
                            double Rr = 0.0;
                            if constexpr (EnvTraits<ENV>::kContinuous && EnvTraits<ENV>::AD == 1) {
                                Rr = rollout_fast1d<ENV, T>(p.seed, wv_sims0 + (unsigned)tid, tree_id, 0u, ss, budget, p.disc_table, t, rr);
                            } else {
                                Stream rs;
                                rs.init(p.seed, wv_sims0 + (unsigned)tid, tree_id, 0u);
                                bool done = false;
                                while (!done && t < budget) {
                                    done = rollout_step<ENV, T>(rs, ss, p.grid, rr, p.heur, p.noise, 0);
                                    if constexpr (!Rewards<ENV>::kClosedForm) Rr = __dadd_rn(Rr, __dmul_rn((double)rr, __ldg(p.disc_table + t)));
                                    ++t;
                                }
                                if constexpr (Rewards<ENV>::kClosedForm) Rr = Rewards<ENV>::rollout_return(t, (double)rr, p.disc_table, p.cum_table);
                            }
                            my_steps += t;
                            const double G = __dadd_rn(r0, __dmul_rn(p.gamma, Rr));
                            w_g[tid] = G;
                            // the new action node and its state node, as the serial expansion + backup leave them
                            const int a = nA0 + tid, c = nS0 + tid;
                            q.a_total[a] = G; q.a_value[a] = val; q.a_na[a] = 1; q.a_visits[a] = 1; q.a_child[a] = c;
                            if (p.action_dim == 2) q.a_value2[a] = val2;
                            q.s_total[c] = G; q.s_term_reward[c] = 0.0; q.s_ns[c] = 1; q.s_flags[c] = 0;
                            q.s_coff[c] = 0; q.s_ccap[c] = 0; q.s_ccnt[c] = 0; q.s_next[c] = -1;
                            T* d = ctx.state_ptr(c);
#pragma unroll
                            for (int i = 0; i < S; ++i) d[i] = st[i];
                            q.ipool[new_off + n0 + tid] = a;
                            // action hash set: claim an empty slot (all keys of the run are distinct and new)
                            const long long bits = ctx.key_bits(val, val2);
                            unsigned long long h = ctx.hash_slot(0, bits);
                            int4* const tab = reinterpret_cast<int4*>(ctx.hash);
                            for (;;) {
                                if (atomicCAS(&tab[h].z, -1, 0) == -1) {
                                    tab[h].x = (int)(unsigned)(bits & 0xffffffffll); tab[h].y = (int)(bits >> 32); tab[h].w = n0 + tid;
                                    break;
                                }
                                h = (h + 1) & (unsigned long long)(p.hash_cap - 1);
                            }
                        }
                        __syncthreads();
                        if (tid == 0) {
                            // the root, in simulation order: ns += 1, total += G (mcts_action_...:80-84)
                            double tot = q.s_total[0];
                            for (int i = 0; i < We; ++i) tot = __dadd_rn(tot, w_g[i]);
                            q.s_total[0] = tot; q.s_ns[0] = ns0 + We; q.s_ccnt[0] = n0 + We;
                            if (wv_cap < -1) { q.s_coff[0] = new_off; q.s_ccap[0] = -wv_cap - 2; }
                        }
                        if (leader) {
                            ctx.nS = nS0 + We; ctx.nA = nA0 + We; ctx.itop = wv_itop;
                            ctx.ts.resume(p.seed, 0u, tree_id, kStreamTree, (uint32_t)(wv_rng0 + We * kWaveWords));
                            ctx.c_env += We; ctx.c_levels += We; ctx.c_nodes += We;
                        }
                        sims_done += We; s_done += We;
                        prev_levels = 1;
                        sim += We - 1;
                    }
                    __syncthreads();
                    if (wv_cap != -1) continue;
                }
            }
        }
        // ------------------------------------------------------------ warp 0: descent to the leaf
        int level = 0;           // number of recorded path entries
        bool need_roll = false;  // the leaf is a fresh non-terminal state: roll out from it
        double G = 0.0, r_leaf = 0.0;
        int leaf_a = -1, leaf_s = -1;
        T cur[S];
        const int pb = tp ? tp_pb : 0;   // first path entry of this simulation
        CTA_TM(tm_other);
        if (leader && !ctx.fault) {
            Pool& q = ctx.q;
            int s = 0;
            if constexpr (!SCRIPTED) {
                // Vanilla trees re-walk almost the same path every simulation.  The 32 lanes re-evaluate UCB1 on 32
                // LEVELS of the previous path at once (the statistics are already backed up); the first level that
                // has an unvisited action, an exact tie (needs a draw) or a different arg-max is where the serial
                // descent resumes.  The old path's last level is always re-decided serially.
                if (p.agent == ISLA_AGENT_VANILLA && prev_levels > 1 && !tp) {
                    const int Lv = prev_levels - 1;
                    int event = Lv;
                    for (int l0 = 0; l0 < Lv; l0 += 32) {
                        const int l = l0 + lane;
                        bool my_event = false;
                        if (l < Lv) {
                            const int sl = q.p_s[l], al = q.p_a[l];
                            const int n = q.s_ccnt[sl], off = q.s_coff[sl];
                            if (n < p.n_actions) my_event = true;
                            else {
                                const int ns = q.s_ns[sl];
                                const double lg = ns < p.ln_count ? p.ln_table[ns] : ::log((double)ns);
                                // fp32 screening (MUFU approximations, error ~1e-6 of the score magnitude): a gap above
                                // 1e-5 of that magnitude between the two largest scores settles the arg-max; only
                                // near-ties pay for the float64 scores, so the decision is always the float64 one
                                bool decided = false;
                                {
                                    const float lgf = (float)lg, Cf = (float)p.C;
                                    float top = -INFINITY, second = -INFINITY, scale = 0.f; int argf = -1;
                                    for (int i = 0; i < n; ++i) {
                                        const int a = q.ipool[off + i];
                                        float rn, ef;
                                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"((float)q.a_na[a]));
                                        const float qf = (float)q.a_total[a] * rn;
                                        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ef) : "f"(lgf * rn));
                                        ef *= Cf;
                                        const float sc = qf + ef;
                                        scale = fmaxf(scale, fabsf(qf) + fabsf(ef));
                                        if (sc > top) { second = top; top = sc; argf = a; } else second = fmaxf(second, sc);
                                    }
                                    if (top - second > 1e-5f * scale + 1e-30f) { decided = true; my_event = argf != al; }
                                }
                                if (!decided) {
                                    double best = -INFINITY; int arg = -1, ties = 0;
                                    for (int i = 0; i < n; ++i) {
                                        const int a = q.ipool[off + i];
                                        const double na = (double)q.a_na[a];
                                        const double sc = __dadd_rn(__ddiv_rn(q.a_total[a], na), __dmul_rn(p.C, __dsqrt_rn(__ddiv_rn(lg, na))));
                                        if (sc > best) { best = sc; arg = a; ties = 1; } else if (sc == best) ++ties;
                                    }
                                    my_event = (ties != 1) || (arg != al);
                                }
                            }
                        }
                        const unsigned ev = __ballot_sync(kFull, my_event);
                        if (ev) { event = l0 + __ffs(ev) - 1; break; }
                    }
                    level = event;
                    s = q.p_s[event];
                    ctx.c_levels += (unsigned long long)event;
                }
            }
            CTA_TM(tm_verify);
            for (;;) {
                ++ctx.c_levels;
                // ---------------- state node: choose or create the action node
                int a = -1, a_rank = -1;
                const int n_act = q.s_ccnt[s], off = q.s_coff[s];
                if (p.agent == ISLA_AGENT_VANILLA || p.agent == ISLA_AGENT_SPW) {
                    // unvisited actions = those without a node (visit_actions[a] == 0), mcts.py:55-59
                    // (up to 128 actions: DiscreteCurveEnv grids)
                    const int nun = p.n_actions - n_act;
                    if (nun > 0) {
                        unsigned have[4] = {0u, 0u, 0u, 0u};
                        for (int i = 0; i < n_act; ++i) { const int k = (int)q.a_value[q.ipool[off + i]]; have[k >> 5] |= 1u << (k & 31); }
                        int pos = ctx.choice(nun), disc = 0;
                        for (int k = 0; k < p.n_actions; ++k) if (!((have[k >> 5] >> (k & 31)) & 1u)) { if (pos == 0) disc = k; --pos; }
                        a = ctx.new_action((double)disc);
                        ctx.append_action(s, a);
                    } else {
                        a_rank = ctx.ucb_pick(s);
                        a = q.ipool[off + a_rank];
                    }
                } else {
                    // progressive widening of the action set: len(actions) <= ceil(k * ns**alpha)
                    const int ns_now = q.s_ns[s];
                    const double cap = ns_now < p.ln_count ? p.pw1_table[ns_now] : ceil(__dmul_rn(p.k1, pow((double)ns_now, p.alpha1)));
                    if (n_act == 0 || (double)n_act <= cap) {
                        double val, val2 = 0.0;
                        if (p.agent == ISLA_AGENT_APW && p.expansion == ISLA_EXPAND_VOO) val = ctx.voo_expand(s, val2, red, &sh, nt);
                        else if (p.agent == ISLA_AGENT_APW && p.expansion == ISLA_EXPAND_GENETIC2) val = ctx.genetic2_expand(s, val2);
                        else if (p.agent == ISLA_AGENT_APW && p.expansion == ISLA_EXPAND_GENETIC) val = ctx.genetic_expand(s, val2);
                        else val = ctx.tree_action(val2);  // dpw ignores --ae (mcts_double_...:57)
                        const int old = ctx.find_action(s, val, val2);
                        if (p.agent == ISLA_AGENT_APW) {
                            if (old >= 0) a = q.ipool[q.s_coff[s] + old];                  // reuse (:66-70)
                            else { a = ctx.new_action(val, val2); ctx.append_action(s, a); ctx.hash_put(s, val, q.s_ccnt[s] - 1, val2); }
                        } else {
                            a = ctx.new_action(val);                                      // always a new node (:58-60)
                            if (old >= 0) { if (lane == 0) q.ipool[q.s_coff[s] + old] = a; __syncwarp(); }
                            else { ctx.append_action(s, a); ctx.hash_put(s, val, q.s_ccnt[s] - 1); }
                        }
                    } else {
                        a_rank = ctx.ucb_pick(s);
                        a = q.ipool[off + a_rank];
                    }
                }
                if (ctx.fault || ctx.tc.fault) break;
                // ---------------- action node: step the env from this state node's data
                {
                    const T* sp = ctx.state_ptr(s);
#pragma unroll
                    for (int i = 0; i < S; ++i) cur[i] = sp[i];
                }
                T rT;
                double step_a = q.a_value[a];
                if constexpr (ENV == ENV_FROZENLAKE) {
                    if (p.noise) {
                        // is_slippery=True.  Mcts / MctsHash / apw step a CLONE of the env per simulation: this is draw
                        // `level` of the clone.  spw / dpw step the live env: the draws run on through the whole fit().
                        const bool live_env = p.agent == ISLA_AGENT_SPW || p.agent == ISLA_AGENT_DPW;
                        const int k = live_env ? ctx.noise_pos : level;
                        if (k >= p.noise_n) { ctx.fault = ISLA_ERR_INVALID; break; }
                        step_a = (double)lake_slip((int)step_a, __ldg(p.noise + k));
                        if (live_env) ++ctx.noise_pos;
                    }
                }
                const bool term = env_step_action<ENV, T>(cur, step_a, p.action_dim == 2 ? q.a_value2[a] : 0.0, p.grid, rT);
                const double r = (double)rT;
                ++ctx.c_env;
                const int child = q.a_child[a];
                bool descend = false, update_action = true;
                int resampled = -1;   // spw / dpw: the child picked by the resampling branch
                if (p.agent == ISLA_AGENT_VANILLA || p.agent == ISLA_AGENT_APW) {
                    if (term) {
                        // terminal: get-or-create the flagged child; total += r; na += 1; child.ns += 1
                        int c = child;
                        if (c < 0) { c = ctx.new_state(cur, true, 0.0); if (lane == 0) q.a_child[a] = c; }
                        __syncwarp();
                        if (lane == 0 && !ctx.fault) { q.a_total[a] += r; q.a_na[a] += 1; q.s_ns[c] += 1; }
                        G = r;
                    } else if (child < 0) {
                        const int c = ctx.new_state(cur, false, 0.0);
                        if (lane == 0) q.a_child[a] = c;
                        need_roll = true; leaf_a = a; leaf_s = c; r_leaf = r;
                    } else {
                        descend = true;
                    }
                } else {
                    const double kk = p.agent == ISLA_AGENT_SPW ? p.k1 : p.k2;
                    const double al = p.agent == ISLA_AGENT_SPW ? p.alpha1 : p.alpha2;
                    // children dict of the action node: one entry per distinct observation (deterministic env: one)
                    int prev = -1;
                    const int old = ctx.find_child(a, cur, prev);      // the entry under this observation's key, if any
                    const int n_child = ctx.count_children(a);
                    if (old < 0) { prev = -1; for (int c = q.a_child[a]; c >= 0; c = q.s_next[c]) prev = c; }   // tail
                    const int na_now = q.a_na[a];
                    const double wcap = na_now < p.ln_count ? p.pw2_table[na_now] : __dmul_rn(kk, pow((double)na_now, al));
                    if (p.agent == ISLA_AGENT_SPW && term) {
                        // always a FRESH terminal child that overwrites the key (mcts_state_...:81-95)
                        const int c = ctx.new_state(cur, true, r);
                        if (!ctx.fault) ctx.put_child(a, old, prev, c);
                        if (lane == 0 && !ctx.fault) { q.a_total[a] += r; q.a_na[a] += 1; q.s_ns[c] += 1; }
                        G = r;
                    } else if (n_child == 0 || (double)n_child <= wcap) {
                        // widen: the new node REPLACES the child stored under the same observation key
                        const bool t2 = (p.agent == ISLA_AGENT_DPW) && term;
                        const int c = ctx.new_state(cur, t2, 0.0);
                        if (!ctx.fault) ctx.put_child(a, old, prev, c);
                        if (t2) {
                            if (lane == 0 && !ctx.fault) { q.a_total[a] += r; q.a_na[a] += 1; q.s_ns[c] += 1; }
                            G = r;
                        } else {
                            need_roll = true; leaf_a = a; leaf_s = c; r_leaf = r;
                        }
                    } else {
                        // resample an existing child: random.choices(children, weights = na / ns_c) (mcts_state_...:114-119;
                        // dpw: the non-terminal children only, mcts_double_...:119-129)
                        int n_cand = 0;
                        double wtot = 0.0;
                        for (int c = q.a_child[a]; c >= 0; c = q.s_next[c]) {
                            if (p.agent == ISLA_AGENT_DPW && (q.s_flags[c] & 1)) continue;
                            wtot = __dadd_rn(wtot, __ddiv_rn((double)na_now, (double)q.s_ns[c]));
                            ++n_cand;
                        }
                        if (n_cand == 0) {
                            ctx.fault = ISLA_ERR_INVALID;  // reference: random.choices on an empty population raises
                        } else {
                            int pick_pos;
                            if constexpr (SCRIPTED) {
                                pick_pos = (int)ctx.tc.take(ISLA_K_CHOICES);
                                if (pick_pos < 0 || pick_pos >= n_cand) { if (!ctx.tc.fault) ctx.tc.fault = ISLA_ERR_TRACE; pick_pos = 0; }
                            } else {
                                // one draw; first candidate whose cumulative weight exceeds u * total
                                const double u = __dmul_rn(u32_to_d01(ctx.ts.next()), wtot);
                                double cum = 0.0;
                                int i = 0;
                                pick_pos = n_cand - 1;
                                for (int c = q.a_child[a]; c >= 0; c = q.s_next[c]) {
                                    if (p.agent == ISLA_AGENT_DPW && (q.s_flags[c] & 1)) continue;
                                    cum = __dadd_rn(cum, __ddiv_rn((double)na_now, (double)q.s_ns[c]));
                                    if (cum > u) { pick_pos = i; break; }
                                    ++i;
                                }
                            }
                            int pick = -1, i = 0;
                            for (int c = q.a_child[a]; c >= 0; c = q.s_next[c]) {
                                if (p.agent == ISLA_AGENT_DPW && (q.s_flags[c] & 1)) continue;
                                if (i == pick_pos) { pick = c; break; }
                                ++i;
                            }
                            const bool child_term = q.s_flags[pick] & 1;
                            if (p.agent == ISLA_AGENT_SPW && child_term) {
                                const double tr = q.s_term_reward[pick];
                                if (lane == 0) { q.a_total[a] += tr; q.a_na[a] += 1; q.s_ns[pick] += 1; }
                                G = tr;
                            } else {
                                // the env continues from the resampled child's observation (:126-127); the reward of the
                                // step just taken is kept; self.total / self.na stay untouched (:126-130)
                                descend = true; update_action = false;
                                resampled = pick;
                            }
                        }
                    }
                }
                __syncwarp();
                if (ctx.fault || ctx.tc.fault) break;
                // record this level: the state node is always backed up, the action node unless skipped
                if (pb + level >= p.s_cap) { ctx.fault = ISLA_ERR_CAPACITY; break; }
                if (lane == 0) {
                    q.p_s[pb + level] = s; q.p_a[pb + level] = a; q.p_r[pb + level] = r;
                    q.p_flag[pb + level] = (descend ? 1 : 0) | (update_action ? 2 : 0);
                }
                ++level;
                if (!descend) break;
                s = resampled >= 0 ? resampled : child;
            }
            const bool live_env = p.noise && (p.agent == ISLA_AGENT_SPW || p.agent == ISLA_AGENT_DPW);
            const int noise_base = live_env ? ctx.noise_pos : level;
            if constexpr (ENV == ENV_FROZENLAKE) {   // the noise table must cover every step the rollouts can take
                const int b_ = p.budget_mode == ISLA_BUDGET_ZERO ? 0 : max(0, min(p.max_depth - level, p.disc_count));
                if (p.noise && need_roll && noise_base + b_ > p.noise_n) ctx.fault = ISLA_ERR_INVALID;
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < S; ++i) sh.leaf[i] = (double)cur[i];
                sh.budget = p.budget_mode == ISLA_BUDGET_ZERO ? 0 : max(0, min(p.max_depth - level, p.disc_count));
                sh.depth = noise_base;
                sh.roll_steps = 0;
                sh.need = (need_roll && !ctx.fault && !ctx.tc.fault) ? 1 : 0;
                sh.sim = sims_done;
            }
        } else if (leader && lane == 0) {
            sh.need = 0;
        }
        if (owner) {
        // end of the descent.  Until then the other warps serve voo's try batches (barrier 1 = work posted or descent
        // over, barrier 2 = batch evaluated); the last barrier-1 also publishes sh.need / sh.leaf / sh.budget.
        if (leader) {
            CTA_TM(tm_descent);
            if (lane == 0) sh.vw.cmd = VOO_EXIT;
            __syncwarp();
            bar_sync(1, nt);
        } else {
            for (;;) {
                bar_sync(1, nt);
                if (sh.vw.cmd == VOO_EXIT) break;
                if (sh.vw.cmd == VOO_FIND) voo_find_closest<ENV>(p, sh.vw, tree_id, tid, nt);
                else voo_eval_tries<ENV>(p, ctx.q, sh.vw, tree_id, tid);
                bar_sync(2, nt);
            }
        }
        }
        if (CS > 1) cg::this_cluster().sync();   // the owner's leaf record is final
        CTA_TM(tm_r_sync1);

        if (tp) {
            // ---------------------------------------------------------- tree-parallel: park this simulation, run the batch when full
            if (leader) {
                Pool& q = ctx.q;
                const bool ok = !ctx.fault && !ctx.tc.fault && level > 0;
                if (ok) {
                    const double Gp = need_roll ? __dadd_rn(r_leaf, __dmul_rn(p.gamma, p.virtual_value)) : G;
                    if (need_roll && lane == 0) { q.a_na[leaf_a] += 1; q.s_ns[leaf_s] += 1; q.a_total[leaf_a] += Gp; q.s_total[leaf_s] += Gp; }
                    if (lane == 0) {
                        double g = Gp;
                        q.p_g[pb + level - 1] = g;
                        for (int l = level - 2; l >= 0; --l) { g = __dadd_rn(q.p_r[pb + l], __dmul_rn(p.gamma, g)); q.p_g[pb + l] = g; }
                    }
                    __syncwarp();
                    // (a node can occur once per path only: the updates of one simulation touch distinct nodes)
                    for (int l = lane; l < level; l += 32) {
                        const int s_ = q.p_s[pb + l], a_ = q.p_a[pb + l], fl = q.p_flag[pb + l];
                        const double g = q.p_g[pb + l];
                        if (l < level - 1 && (fl & 2)) { q.a_total[a_] += g; q.a_na[a_] += 1; }
                        q.s_ns[s_] += 1; q.a_visits[a_] += 1; q.s_total[s_] += g;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        TpRec& rc = tp_rec[tp_b];
#pragma unroll
                        for (int i = 0; i < S; ++i) rc.leaf[i] = (double)cur[i];
                        rc.r_leaf = r_leaf; rc.g_prov = Gp;
                        rc.budget = sh.budget; rc.need = need_roll ? 1 : 0; rc.depth = level;
                        rc.leaf_a = leaf_a; rc.leaf_s = leaf_s; rc.level = level; rc.pb = pb; rc.sim = sims_done;
                    }
                    ++sims_done; ++s_done;
                }
                if (lane == 0) { sh.need = ok ? 1 : 0; sh.depth = ok ? level : 0; }
            }
            __syncthreads();
            const int parked = sh.need, parked_levels = sh.depth;
            if (parked) { ++tp_b; tp_pb += parked_levels; }
            const bool flush = tp_b > 0 && (tp_b == tpB || sim == p.n_sim - 1 || !parked);
            if (flush) {
                const int nb = tp_b;
                // every rollout (b, j) of the batch: index b * P + j
                for (int i0 = 0; i0 < nb * P; i0 += nt) {
                    const int i = i0 + tid;
                    if (i < nb * P) {
                        const int b = i / P, j = i - b * P;
                        const TpRec& rc = tp_rec[b];
                        double Rr = 0.0;
                        if (rc.need && j < R) {
                            T st[S];
#pragma unroll
                            for (int k = 0; k < S; ++k) st[k] = T(rc.leaf[k]);
                            int t = 0;
                            T rr = T(0);
                            if constexpr (EnvTraits<ENV>::kContinuous && EnvTraits<ENV>::AD == 1) {
                                Rr = rollout_fast1d<ENV, T>(p.seed, rc.sim, tree_id, (uint32_t)j, st, rc.budget, p.disc_table, t, rr);
                            } else {
                                Stream rs;
                                rs.init(p.seed, rc.sim, tree_id, (uint32_t)j);
                                bool done = false;
                                while (!done && t < rc.budget) {
                                    done = rollout_step<ENV, T>(rs, st, p.grid, rr, p.heur, p.noise, min(rc.depth + t, p.noise_n - 1));
                                    if constexpr (!Rewards<ENV>::kClosedForm) Rr = __dadd_rn(Rr, __dmul_rn((double)rr, __ldg(p.disc_table + t)));
                                    ++t;
                                }
                                if constexpr (Rewards<ENV>::kClosedForm) Rr = Rewards<ENV>::rollout_return(t, (double)rr, p.disc_table, p.cum_table);
                            }
                            my_steps += t;
                        }
                        red[i] = Rr;
                    }
                }
                __syncthreads();
                for (int stride = P >> 1; stride > 0; stride >>= 1) {   // the fixed pairwise tree, per leaf
                    for (int i = tid; i < nb * stride; i += nt) {
                        const int b = i / stride, k = i - b * stride;
                        red[b * P + k] = __dadd_rn(red[b * P + k], red[b * P + k + stride]);
                    }
                    __syncthreads();
                }
                // corrections, in simulation order: totals move from the provisional to the real return, counts stay
                if (leader) {
                    Pool& q = ctx.q;
                    for (int b = 0; b < nb; ++b) {
                        const TpRec& rc = tp_rec[b];
                        if (!rc.need) continue;
                        const double mean = R == 1 ? red[b * P] : __ddiv_rn(red[b * P], (double)R);
                        const double Gr = __dadd_rn(rc.r_leaf, __dmul_rn(p.gamma, mean));
                        const double d0 = __dsub_rn(Gr, rc.g_prov);
                        if (lane == 0) {
                            q.a_total[rc.leaf_a] += d0; q.s_total[rc.leaf_s] += d0;
                            double d = d0;
                            q.p_g[rc.pb + rc.level - 1] = d;
                            for (int l = rc.level - 2; l >= 0; --l) { d = __dmul_rn(p.gamma, d); q.p_g[rc.pb + l] = d; }
                        }
                        __syncwarp();
                        for (int l = lane; l < rc.level; l += 32) {
                            const int s_ = q.p_s[rc.pb + l], a_ = q.p_a[rc.pb + l], fl = q.p_flag[rc.pb + l];
                            const double d = q.p_g[rc.pb + l];
                            if (l < rc.level - 1 && (fl & 2)) q.a_total[a_] += d;
                            q.s_total[s_] += d;
                        }
                        __syncwarp();
                    }
                }
                tp_b = 0; tp_pb = 0;
            }
            prev_levels = 0;
            __syncthreads();   // sh.* and tp_rec are rewritten by the next descent
            continue;
        }

        // ------------------------------------------------------------ all threads: R rollouts from the leaf
        const bool need = sh0->need != 0;
        if (need) {
            const int budget = sh0->budget;
            if constexpr (SCRIPTED) {
                // the trace scripts ONE rollout (R == 1), replayed by every lane of warp 0
                if (leader) {
                    T st[S];
#pragma unroll
                    for (int i = 0; i < S; ++i) st[i] = T(sh.leaf[i]);
                    double Rr = 0.0; int t = 0; bool done = false;
                    while (!done && t < budget && !ctx.tc.fault) {
                        double act2 = 0.0, act;
                        bool heuristic = false;
                        if constexpr (ENV == ENV_FROZENLAKE) heuristic = p.heur != nullptr;
                        if (heuristic) act = (double)lake_heuristic_action(p.heur, (int)st[0], [&](int ties) { return ctx.choice(ties); });
                        else act = ctx.tree_action(act2);
                        if (ctx.tc.fault) break;
                        if constexpr (ENV == ENV_FROZENLAKE) {
                            if (p.noise) {
                                if (sh.depth + t >= p.noise_n) { ctx.fault = ISLA_ERR_INVALID; break; }
                                act = (double)lake_slip((int)act, __ldg(p.noise + sh.depth + t));
                            }
                        }
                        T rr;
                        done = env_step_action<ENV, T>(st, act, act2, p.grid, rr);
                        Rr = __dadd_rn(Rr, __dmul_rn((double)rr, p.disc_table[t]));
                        ++t;
                    }
                    if (tid == 0) { my_steps += t; red[0] = Rr; sh.roll_steps = t; }
                }
                __syncthreads();
            } else {
                // this CTA plays the rollouts j = jl * CS + rank (jl < slice): the single-CTA reduction tree (strides
                // P/2 ... 1) first combines indices with equal low bits, i.e. exactly the elements of one rank, and
                // ends with the CS per-rank sums
                const int slice = P / CS;
                const unsigned sim_id = sh0->sim;
                const int leaf_depth = sh0->depth;
                double leaf[S];
#pragma unroll
                for (int i = 0; i < S; ++i) leaf[i] = sh0->leaf[i];
                // Episode lengths are random (CartPole: mean 22 steps, longest of a warp ~56): the slice is handed out
                // dynamically.  The warp meets every kRollMeet env steps; lanes that finished store their return under
                // the rollout's index (the reduction below does not care who played it) and claim the next unplayed
                // rollouts of the CTA -- one shared-memory atomic per warp and meeting, ranks by ballot.  Between
                // meetings the lanes only step.  (A fixed rollout-per-thread assignment in passes of the CTA size ran 12
                // of 32 lanes on C2 as a batch: every pass waited for its longest episode.)
#ifndef ISLA_CTA_STATIC_BLOCK
#define ISLA_CTA_STATIC_BLOCK 4
#endif
#ifndef ISLA_CTA_MEET
#define ISLA_CTA_MEET 8
#endif
#ifndef ISLA_CTA_UNROLL
#define ISLA_CTA_UNROLL 1
#endif
                constexpr int kRollMeet = ISLA_CTA_MEET, kRollUnroll = ISLA_CTA_UNROLL;
                if (slice <= nt) {
                    // one rollout per thread at most: nothing to hand out (one tree on a cluster: the phase lasts as long as
                    // its longest episode, and the lean loop keeps that chain short)
                    if (tid < slice) {
                        const int j = tid * CS + (int)crank;
                        double Rr = 0.0;
                        if (j < R) {
                            T st[S];
#pragma unroll
                            for (int i = 0; i < S; ++i) st[i] = T(leaf[i]);
                            int t = 0;
                            T rr = T(0);
                            if constexpr (EnvTraits<ENV>::kContinuous && EnvTraits<ENV>::AD == 1) {
                                Rr = rollout_fast1d<ENV, T>(p.seed, sim_id, tree_id, (uint32_t)j, st, budget, p.disc_table, t, rr);
                            } else if constexpr (ENV == ENV_CARTPOLE) {
                                t = rollout_cartpole_blocks<T, ISLA_CTA_STATIC_BLOCK>(p.seed, sim_id, tree_id, (uint32_t)j, st, budget, rr);
                                Rr = Rewards<ENV>::rollout_return(t, (double)rr, p.disc_table, p.cum_table);
                            } else {
                                Stream rs;
                                rs.init(p.seed, sim_id, tree_id, (uint32_t)j);
                                bool done = false;
                                while (!done && t < budget) {
                                    done = rollout_step<ENV, T>(rs, st, p.grid, rr, p.heur, p.noise, min(leaf_depth + t, p.noise_n - 1));
                                    if constexpr (!Rewards<ENV>::kClosedForm) Rr = __dadd_rn(Rr, __dmul_rn((double)rr, __ldg(p.disc_table + t)));
                                    ++t;
                                }
                                if constexpr (Rewards<ENV>::kClosedForm) Rr = Rewards<ENV>::rollout_return(t, (double)rr, p.disc_table, p.cum_table);
                            }
                            my_steps += t;
                            if (j == 0) sh.roll_steps = t;
                        }
                        red[tid] = Rr;
                    }
                } else {
                    const unsigned lt = (1u << lane) - 1u;
                    bool more = true;                   // warp-uniform: unclaimed rollouts may remain
                    bool act = false, pend = false, done = false;
                    int jl = 0, j = 0, t = 0;
                    T st[S];
#pragma unroll
                    for (int i = 0; i < S; ++i) st[i] = T(0);
                    T rr = T(0);
                    double Rr = 0.0;
                    Stream rs;
                    rs.init(p.seed, sim_id, tree_id, 0u);
                    uint32_t dyn_word = 0u;   // CartPole blocks: the stream word holding the rollout's current 32 action bits
                    (void)dyn_word;
                    for (;;) {
                        if (pend) {
                            if constexpr (Rewards<ENV>::kClosedForm) Rr = Rewards<ENV>::rollout_return(t, (double)rr, p.disc_table, p.cum_table);
                            red[jl] = Rr;
                            my_steps += t;
                            if (j == 0) sh.roll_steps = t;
                            pend = false;
                        }
                        const unsigned idle = __ballot_sync(kFull, !act);
                        if (more && idle != 0u) {
                            const int cnt = __popc(idle);
                            int base = 0;
                            if (lane == 0) base = atomicAdd(&sh.roll_next, cnt);
                            base = __shfl_sync(kFull, base, 0);
                            if (!act) {
                                const int mine = base + __popc(idle & lt);
                                if (mine < slice) {
                                    jl = mine; j = jl * CS + (int)crank;
                                    if (j < R) {
#pragma unroll
                                        for (int i = 0; i < S; ++i) st[i] = T(leaf[i]);
                                        rs.init(p.seed, sim_id, tree_id, (uint32_t)j);
                                        t = 0; done = false; rr = T(0); Rr = 0.0;
                                        act = budget > 0;
                                        pend = !act;
                                    } else {
                                        red[jl] = 0.0;   // padding of the power-of-two reduction tree
                                    }
                                }
                            }
                            more = base + cnt < slice;
                        }
                        if (!__any_sync(kFull, act || pend)) {
                            if (!more) break;
                            continue;
                        }
#ifndef ISLA_CTA_DYN_BLOCK
#define ISLA_CTA_DYN_BLOCK 4
#endif
                        if constexpr (ENV == ENV_CARTPOLE && ISLA_CTA_DYN_BLOCK > 1) {
                            // CartPole: blocks of predicated steps (the tail of one step overlaps the pole-angle chain of
                            // the next; 16 warps per SM leave the latency of a dependent chain exposed).  Step t draws
                            // bit t of the rollout's stream; a rollout starts at t = 0, so blocks never straddle a word.
                            constexpr int NB = ISLA_CTA_DYN_BLOCK;
                            static_assert(kRollMeet % NB == 0 && 32 % NB == 0, "blocks per meeting");
#pragma unroll 1
                            for (int u = 0; u < kRollMeet; u += NB) {
                                if (act && (t & 31) == 0) dyn_word = rs.next();
                                uint32_t bits = dyn_word << (t & 31);
#pragma unroll
                                for (int k = 0; k < NB; ++k) {
                                    T ns[S];
#pragma unroll
                                    for (int i = 0; i < S; ++i) ns[i] = st[i];
                                    T r1;
                                    const bool d1 = Env<ENV, T>::step(ns, T(bits >> 31), r1);
                                    bits <<= 1;
                                    if (act) {
#pragma unroll
                                        for (int i = 0; i < S; ++i) st[i] = ns[i];
                                        rr = r1;
                                        ++t;
                                        if (d1 || t >= budget) { act = false; pend = true; }
                                    }
                                }
                            }
                        } else
#pragma unroll kRollUnroll
                        for (int u = 0; u < kRollMeet; ++u) {
                            if (act) {
                                done = rollout_step<ENV, T>(rs, st, p.grid, rr, p.heur, p.noise, min(leaf_depth + t, p.noise_n - 1));
                                // r * pow(gamma, t) summed left to right (abstract_mcts.py:87); envs with a constant
                                // interior reward take the bit-identical closed form from the host prefix table above
                                if constexpr (!Rewards<ENV>::kClosedForm) Rr = __dadd_rn(Rr, __dmul_rn((double)rr, __ldg(p.disc_table + t)));
                                ++t;
                                if (done || t >= budget) { act = false; pend = true; }
                            }
                        }
                    }
                }
                CTA_TM(tm_r_own);
                __syncthreads();
                CTA_TM(tm_r_wait);
                if (tid == 0) sh.roll_next = 0;   // every claim of this phase has been made; barriers follow before the next
                // fixed power-of-two tree: the sum does not depend on the CTA size nor on the cluster size
                for (int stride = slice >> 1; stride > 0; stride >>= 1) {
                    for (int i = tid; i < stride; i += nt) red[i] = __dadd_rn(red[i], red[i + stride]);
                    __syncthreads();
                }
                if (CS > 1 && tid == 0) cg::this_cluster().map_shared_rank(partials, 0)[crank] = red[0];
                CTA_TM(tm_r_reduce);
            }
        }
        if (CS > 1) {
            cg::this_cluster().sync();   // every rank's sum has landed in the owner's `partials`
            if (owner && need && tid == 0) {
                for (int stride = CS >> 1; stride > 0; stride >>= 1)
                    for (int i = 0; i < stride; ++i) partials[i] = __dadd_rn(partials[i], partials[i + stride]);
                red[0] = partials[0];
            }
            __syncthreads();
        }

        // ------------------------------------------------------------ warp 0: leaf update + backup
        CTA_TM(tm_roll);
        if (leader && !ctx.fault && !ctx.tc.fault && level > 0) {
            Pool& q = ctx.q;
            if (need_roll && p.noise && (p.agent == ISLA_AGENT_SPW || p.agent == ISLA_AGENT_DPW)) ctx.noise_pos += sh.roll_steps;
            if (need_roll) {
                const double mean = R == 1 ? red[0] : __ddiv_rn(red[0], (double)R);
                G = __dadd_rn(r_leaf, __dmul_rn(p.gamma, mean));
                if (lane == 0) { q.a_na[leaf_a] += 1; q.s_ns[leaf_s] += 1; q.a_total[leaf_a] += G; q.s_total[leaf_s] += G; }
            }
            // the deepest level's action node was updated above; unwind the recursion for the rest:
            // the returns g_l = r_l + gamma * g_{l+1} are a serial chain (lane 0), the node updates are not
            if (lane == 0) {
                double g = G;
                q.p_g[level - 1] = g;
                int l = level - 2;
                // eight levels at a time: the rewards are loaded before, the returns stored after the eight dependent
                // (multiply, add) pairs, so the chain pays the arithmetic latency only (a deep vanilla path has 200 levels;
                // with one load and one store inside every link the backup was a quarter of a C2 simulation)
                for (; l >= 7; l -= 8) {
                    double r[8], gs[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) r[k] = q.p_r[l - k];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { g = __dadd_rn(r[k], __dmul_rn(p.gamma, g)); gs[k] = g; }   // r + gamma * build_tree_state(child)
#pragma unroll
                    for (int k = 0; k < 8; ++k) q.p_g[l - k] = gs[k];
                }
                for (; l >= 0; --l) {
                    g = __dadd_rn(q.p_r[l], __dmul_rn(p.gamma, g));
                    q.p_g[l] = g;
                }
            }
            __syncwarp();
            for (int l = lane; l < level; l += 32) {
                const int s = q.p_s[l], a = q.p_a[l], fl = q.p_flag[l];
                const double g = q.p_g[l];
                if (l < level - 1 && (fl & 2)) { q.a_total[a] += g; q.a_na[a] += 1; }
                q.s_ns[s] += 1; q.a_visits[a] += 1; q.s_total[s] += g;          // mcts.py:63-66
            }
            __syncwarp();
            ++sims_done; ++s_done;
            prev_levels = level;
        } else {
            prev_levels = 0;
        }
        CTA_TM(tm_backup);
        __syncthreads();
    }
#ifdef ISLA_CTA_TIMING
    if (tree == 0 && owner && tid == 0)
        printf("cta tree 0 cycles: verify %lld descent %lld rollout-phase %lld (+ leaf hand-over %lld, own rollout %lld, wait for the CTA %lld, reduction %lld) backup %lld other %lld\n",
               tm_verify, tm_descent, tm_roll, tm_r_sync1, tm_r_own, tm_r_wait, tm_r_reduce, tm_backup, tm_other);
#endif

    if (my_steps) { atomicAdd(p.counters + CNT_ROLLOUT_STEPS, my_steps); atomicAdd(p.counters + CNT_ENV_STEPS, my_steps); }
    bool copy_back = p.pool_in_smem && owner;
    if (leader) {
        int flt = ctx.fault ? ctx.fault : ctx.tc.fault;
        if (lane == 0) {
            int* m = ctx.q.meta;
            m[CM_NS] = ctx.nS; m[CM_NA] = ctx.nA; m[CM_ITOP] = ctx.itop; m[CM_SIMS] = (int)sims_done;
            m[CM_RNG] = (int)ctx.ts.consumed(); m[CM_FAULT] = flt; m[CM_NOISE_POS] = ctx.noise_pos;
            if constexpr (SCRIPTED) m[CM_TRACE_POS] = (int)ctx.tc.pos;
            atomicAdd(p.counters + CNT_ENV_STEPS, ctx.c_env);
            atomicAdd(p.counters + CNT_ROLLOUT_STEPS, ctx.c_roll);
            atomicAdd(p.counters + CNT_LEVELS, ctx.c_levels);
            atomicAdd(p.counters + CNT_SIMS, (unsigned long long)s_done);
            atomicAdd(p.counters + CNT_NODES, ctx.c_nodes);
            if (flt) atomicAdd(p.counters + CNT_FAULTS, 1ull);
        }
    }
    if (copy_back) {
        __syncthreads();
        const int* sm = reinterpret_cast<const int*>(reinterpret_cast<char*>(pool_smem) + p.off_s[O_META]);
        copy_pool(p, pool_global, p.off, reinterpret_cast<const char*>(pool_smem), p.off_s, sm[CM_NS], sm[CM_NA], sm[CM_ITOP], tid, nt);
    }
}

// root state + empty pools
template <typename T>
__global__ void cta_set_roots_kernel(const CtaParams p, const double* roots, int S) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= p.n_trees) return;
    Pool q = make_pool(p, tree);
    q.s_total[0] = 0.0; q.s_term_reward[0] = 0.0; q.s_ns[0] = 0; q.s_flags[0] = 0; q.s_coff[0] = 0; q.s_ccap[0] = 0; q.s_ccnt[0] = 0;
    q.s_next[0] = -1;
    T* st = reinterpret_cast<T*>(q.s_state);
    for (int i = 0; i < S; ++i) {
        T v = (T)roots[tree * S + i];
        if (p.agent == ISLA_AGENT_SPW || p.agent == ISLA_AGENT_DPW) v = T(float(v));  // root_data is the float32 observation
        st[i] = v;
    }
    for (int i = 0; i < CM_WORDS; ++i) q.meta[i] = 0;
    q.meta[CM_NS] = 1;
}

// discrete agents: root statistics by action index
__global__ void cta_root_stats_kernel(const CtaParams p, int64_t* na_out, double* total_out) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= p.n_trees) return;
    Pool q = make_pool(p, tree);
    const int A = p.n_actions;
    for (int a = 0; a < A; ++a) { na_out[tree * A + a] = 0; total_out[tree * A + a] = 0.0; }
    if (A == 0) return;
    const int off = q.s_coff[0], n = q.s_ccnt[0];
    for (int i = 0; i < n; ++i) {
        const int a = q.ipool[off + i];
        const int k = (int)q.a_value[a];
        if (k >= 0 && k < A) { na_out[tree * A + k] = q.a_na[a]; total_out[tree * A + k] = q.a_total[a]; }
    }
}

// End of fit(): spw/dpw first re-order root.actions by key (persistently), then q = total/na and the
// uniformly random arg-max.  One thread per tree (root fan-out is small where the order matters).
// One WARP per tree: a progressive-widening root has thousands of children (C3: 10 000), and one thread walking them three
// times cost 3.5 ms per fit -- more than the 1.5 ms search.  Same arithmetic and the same tie rule as before: q = total / na
// in float64, exact ties broken by one draw of the tree stream among the tied children in root.actions order.
__global__ void cta_best_kernel(const CtaParams p, int scripted_pos, int32_t* best_rank_out, double* best_action_out) {
    const long long tree = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tree >= p.n_trees) return;   // whole warps leave
    Pool q = make_pool(p, tree);
    const int off = q.s_coff[0], n = q.s_ccnt[0];
    const int AD = p.action_dim;
    if (n == 0) {
        if (lane == 0) { best_rank_out[tree] = -1; if (best_action_out) for (int d = 0; d < AD; ++d) best_action_out[tree * AD + d] = 0.0; }
        return;
    }
    if (p.agent == ISLA_AGENT_SPW || p.agent == ISLA_AGENT_DPW) {
        if (lane == 0) {
            for (int x = 1; x < n; ++x) {  // stable insertion sort == sorted(items)
                const int cur = q.ipool[off + x];
                const double kv = q.a_value[cur];
                int y = x;
                while (y > 0) {
                    const double pv = q.a_value[q.ipool[off + y - 1]];
                    bool less;
                    if (p.agent == ISLA_AGENT_SPW) less = kv < pv;
                    else {  // python bytes order of the little-endian float32 keys
                        const unsigned ku = __float_as_uint((float)kv), pu = __float_as_uint((float)pv);
                        const unsigned kb = __byte_perm(ku, 0, 0x0123), pb = __byte_perm(pu, 0, 0x0123);
                        less = kb < pb;
                    }
                    if (!less) break;
                    q.ipool[off + y] = q.ipool[off + y - 1];
                    --y;
                }
                q.ipool[off + y] = cur;
            }
            if (p.hash_cap > 0) {  // the ranks stored in the action hash set follow the new order
                int4* hash = reinterpret_cast<int4*>(p.hash_base + tree * p.hash_cap * 16);
                for (int i = 0; i < n; ++i) {
                    const long long bits = __double_as_longlong(q.a_value[q.ipool[off + i]]);
                    unsigned long long h = (unsigned long long)bits * 0x9E3779B97F4A7C15ull;
                    h ^= h >> 29;  // state 0: the state term of Search::hash_slot vanishes
                    h &= (unsigned long long)(p.hash_cap - 1);
                    for (;;) {
                        const int4 e = hash[h];
                        if (e.z < 0) break;
                        if (e.z == 0 && e.x == (int)(unsigned)(bits & 0xffffffffll) && e.y == (int)(bits >> 32)) { hash[h].w = i; break; }
                        h = (h + 1) & (unsigned long long)(p.hash_cap - 1);
                    }
                }
            }
        }
        __syncwarp();
    }
    auto qv = [&](int i) { const int a = q.ipool[off + i]; return __ddiv_rn(q.a_total[a], (double)q.a_na[a]); };
    double qmax = -INFINITY;
    for (int i = lane; i < n; i += 32) qmax = fmax(qmax, qv(i));
    qmax = warp_max(qmax);
    int ties = 0;
    for (int base = 0; base < n; base += 32) ties += __popc(__ballot_sync(kFull, base + lane < n && qv(base + lane) == qmax));
    int pos = 0;
    if (scripted_pos >= 0) pos = scripted_pos < ties ? scripted_pos : ties - 1;
    else if (ties > 1) {
        Stream ts;   // every lane draws the same word; lane 0 records the stream position
        ts.resume(p.seed, 0u, (uint32_t)(p.tree_id_base + (unsigned long long)tree), kStreamTree, (uint32_t)q.meta[CM_RNG]);
        pos = (int)mulhi_u32(ts.next(), (uint32_t)ties);
        __syncwarp();
        if (lane == 0) q.meta[CM_RNG] = (int)ts.consumed();
    }
    int pick = 0;
    for (int base = 0; base < n; base += 32) {
        unsigned b = __ballot_sync(kFull, base + lane < n && qv(base + lane) == qmax);
        const int c = __popc(b);
        if (pos < c) {
            for (int k = 0; k < pos; ++k) b &= b - 1;
            pick = base + __ffs(b) - 1;
            break;
        }
        pos -= c;
    }
    if (lane == 0) {
        best_rank_out[tree] = pick;
        if (best_action_out) {
            best_action_out[tree * AD] = q.a_value[q.ipool[off + pick]];
            if (AD == 2) best_action_out[tree * AD + 1] = q.a_value2[q.ipool[off + pick]];
        }
    }
}

__global__ void cta_root_children_kernel(const CtaParams p, long long tree, int cap, double* action, int64_t* na, double* total, int32_t* count) {
    Pool q = make_pool(p, tree);
    const int off = q.s_coff[0], n = q.s_ccnt[0];
    if (threadIdx.x == 0 && blockIdx.x == 0) *count = n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n && i < cap; i += gridDim.x * blockDim.x) {
        const int a = q.ipool[off + i];
        if (p.action_dim == 2) { action[2 * i] = q.a_value[a]; action[2 * i + 1] = q.a_value2[a]; } else action[i] = q.a_value[a];
        na[i] = q.a_na[a]; total[i] = q.a_total[a];
    }
}

// Post-fit depth analytics (agents/abstract_mcts.py:120-142,190-209; sweep.py:199-200) of one tree, level-synchronous:
// state node levels are propagated from the root, every state is tagged with the root child it hangs under, and the
// leaves (state nodes none of whose action nodes has a child) feed max / sum / count per root child.  The path arrays of
// the pool (p_s, p_a, p_flag: only meaningful during a search launch) serve as scratch.
__global__ void __launch_bounds__(256) cta_depth_stats_kernel(const CtaParams p, long long tree, int cap, int32_t* dmax,
                                                              double* dmean, int32_t* count) {
    Pool q = make_pool(p, tree);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nS = q.meta[CM_NS], n_root = q.s_ccnt[0], root_off = q.s_coff[0];
    int* const lvl = q.p_s; int* const own = q.p_a; int* const leaves = q.p_flag;
    __shared__ int changed;
    for (int s = tid; s < nS; s += nt) lvl[s] = -1;
    for (int i = tid; i < n_root && i < cap; i += nt) { dmax[i] = 0; dmean[i] = 0.0; leaves[i] = 0; }
    if (tid == 0) *count = n_root;
    __syncthreads();
    // the root's children (up to 10^4 under progressive widening) are spread over the CTA
    for (int i = tid; i < n_root; i += nt) {
        for (int c = q.a_child[q.ipool[root_off + i]]; c >= 0; c = q.s_next[c]) { lvl[c] = 1; own[c] = i; }
    }
    __syncthreads();
    for (int cur = 1;; ++cur) {
        if (tid == 0) changed = 0;
        __syncthreads();
        for (int s = tid; s < nS; s += nt) {
            if (lvl[s] != cur) continue;
            const int n = q.s_ccnt[s], off = q.s_coff[s], o = own[s];
            int pushed = 0;
            for (int i = 0; i < n; ++i) {
                for (int c = q.a_child[q.ipool[off + i]]; c >= 0; c = q.s_next[c]) { lvl[c] = cur + 1; own[c] = o; ++pushed; }
            }
            if (pushed) changed = 1;
            if (o < cap) {
                atomicMax(dmax + o, cur);
                if (!pushed) { atomicAdd(leaves + o, 1); atomicAdd(dmean + o, (double)cur); }   // integer-valued sums: order-free
            }
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    for (int i = tid; i < n_root && i < cap; i += nt) dmean[i] = leaves[i] ? __ddiv_rn(dmean[i], (double)leaves[i]) : 0.0;
}

// byte offsets of the SoA arrays of one tree's pool for a capacity of `cap` state / action nodes; returns the pool size
static long long pool_layout(long long cap, int state_dim, int rb, int action_dim, long long* o) {
    long long off = 0;
    const long long i_cap = 2 * cap + 4 * cap + 64;
    auto take = [&](int idx, long long bytes) { o[idx] = off; off = (off + bytes + 255) / 256 * 256; };
    take(O_S_TOTAL, cap * 8); take(O_S_TERMR, cap * 8); take(O_S_NS, cap * 4); take(O_S_FLAGS, cap * 4);
    take(O_S_COFF, cap * 4); take(O_S_CCAP, cap * 4); take(O_S_CCNT, cap * 4); take(O_S_STATE, cap * state_dim * rb);
    take(O_A_TOTAL, cap * 8); take(O_A_VALUE, cap * 8); take(O_A_NA, cap * 4); take(O_A_VISITS, cap * 4);
    take(O_A_CHILD, cap * 4); take(O_IPOOL, i_cap * 4);
    take(O_P_S, cap * 4); take(O_P_A, cap * 4); take(O_P_R, cap * 8); take(O_P_FLAG, cap * 4);
    take(O_META, CM_WORDS * 4);
    take(O_P_G, cap * 8);
    take(O_A_VALUE2, action_dim == 2 ? cap * 8 : 0);
    take(O_S_NEXT, cap * 4);
    o[O_COUNT] = off;
    return off;
}

// n_sim_launch: simulations of the launch being prepared (0: not a search launch).  The shared-memory view of a pool is
// sized for the nodes that can exist at the END of that launch (<= 2 + simulations run since set_roots + n_sim_launch),
// not for the engine's capacity, so a tree keeps living in shared memory while it is small even if the engine was
// sized for many more simulations.
CtaParams make_params(const isla_engine* e, int n_sim_launch = 0) {
    CtaParams p;
    const isla_layout& L = e->lay;
    p.pool_base = e->ws + L.pool_offset; p.pool_stride = L.pool_stride;
    p.s_cap = L.s_cap; p.a_cap = L.a_cap; p.i_cap = 2 * L.a_cap + 4 * L.s_cap + 64;
    const int rb = L.real_bytes;
    p.grid = ActionGrid{e->cfg.env_id == ISLA_ENV_CURVE ? e->cfg.action_grid0 : 0, e->cfg.env_id == ISLA_ENV_CURVE ? e->cfg.action_grid1 : 0};
    p.action_dim = (e->cfg.env_id == ISLA_ENV_CURVE && p.grid.g0 == 0) ? 2 : 1;
    p.state_bytes = L.state_dim * rb;
    long long off = pool_layout(L.s_cap, L.state_dim, rb, p.action_dim, p.off);
    long long cap_s = e->sims_run + (long long)n_sim_launch + 2;
    if (cap_s > L.s_cap) cap_s = L.s_cap;
    p.stride_s = pool_layout(cap_s, L.state_dim, rb, p.action_dim, p.off_s);
    p.counters = (unsigned long long*)(e->ws + L.counters_offset);
    p.trace_kinds = nullptr; p.trace_values = nullptr; p.trace_len = 0;
    p.n_trees = e->cfg.n_trees; p.scripted_tree = -1;
    p.seed = e->cfg.seed; p.tree_id_base = e->cfg.tree_id_base;
    p.C = e->cfg.C; p.gamma = e->cfg.gamma; p.alpha1 = e->cfg.alpha1; p.k1 = e->cfg.k1; p.alpha2 = e->cfg.alpha2; p.k2 = e->cfg.k2;
    p.epsilon = e->cfg.epsilon;
    p.disc_table = (const double*)(e->ws + L.disc_offset); p.ln_table = (const double*)(e->ws + L.ln_offset);
    p.cum_table = (const double*)(e->ws + L.cumdisc_offset);
    p.pw1_table = (const double*)(e->ws + L.pw1_offset); p.pw2_table = (const double*)(e->ws + L.pw2_offset);
    p.ln_count = (int)L.ln_count; p.disc_count = (int)L.disc_count;
    p.n_sim = 0; p.max_depth = e->cfg.max_depth; p.budget_mode = e->cfg.budget_mode; p.agent = e->cfg.agent;
    p.n_actions = e->cfg.n_actions; p.R = e->cfg.rollouts_per_leaf > 0 ? e->cfg.rollouts_per_leaf : 1;
    p.expansion = e->cfg.expansion; p.voo_n = e->cfg.voo_n_samples; p.voo_max_try = e->cfg.voo_max_try; p.real_bytes = rb;
    // A pool that fits beside the static rollout-reduction buffer (227 KB per SM) can live in shared memory for the launch:
    // the descent then costs ~30 cycles per access instead of an L2 round trip.  That pays while ALL trees are resident
    // at once; with more trees than that, shared-memory pools cap the CTAs per SM and pools in HBM/L2 with up to 32 CTAs
    // per SM are 3-4x faster (measured: 4096 DiscreteCurveEnv trees 4.8 -> 17.2 M sims/s).
    // leaf-parallel batches beyond one CTA's 512 threads are spread over a thread-block cluster (portable size <= 8)
    // -- while the clusters of all trees still fit on the SMs at once; batches of many trees are throughput-bound and
    // run best with one CTA per tree (296 trees x 4096 rollouts: 1.48 M sims/s without clusters, 0.29 M with)
    const long long sms_ = e->sm_count > 0 ? e->sm_count : 148;
    int cs = e->cfg.cluster_ctas;
    if (cs <= 0) { cs = 1; while (cs < 8 && p.R > 512 * cs && e->cfg.n_trees * cs * 2 <= sms_) cs <<= 1; }
    int Pw = 1; while (Pw < p.R) Pw <<= 1;
    while (cs > 1 && (Pw / cs) < 1) cs >>= 1;
    p.cluster = cs;
    (void)off;
    const long long fit_per_sm = p.stride_s > 0 ? (190 * 1024) / p.stride_s : 0;
    const long long sms = e->sm_count > 0 ? e->sm_count : 148;
    // (every CTA of a launch reserves the dynamic shared memory, the cluster's helper CTAs included)
    // (tree-parallel batches record several paths at once: their entries need the full-size arrays in HBM)
    p.pool_in_smem = (fit_per_sm >= 1 && e->cfg.n_trees * cs <= sms * fit_per_sm && !(e->cfg.tuning & 16) && e->cfg.tree_parallel <= 1) ? 1 : 0;
    p.hash_base = e->ws ? e->ws + L.hash_offset : nullptr; p.hash_cap = L.hash_cap;
    p.heur = e->heur_dev;
    p.noise = e->noise_dev; p.noise_n = e->noise_n;
    p.tree_parallel = e->cfg.tree_parallel; p.virtual_value = e->cfg.virtual_value;
    return p;
}

long long pool_bytes(const isla_config& cfg, long long s_cap, long long a_cap, int S, int rb) {
    isla_engine tmp;
    tmp.cfg = cfg; tmp.ws = nullptr; tmp.sm_count = 0;
    tmp.lay.pool_offset = 0; tmp.lay.pool_stride = 0; tmp.lay.s_cap = s_cap; tmp.lay.a_cap = a_cap; tmp.lay.real_bytes = rb;
    tmp.lay.state_dim = S; tmp.lay.counters_offset = 0; tmp.lay.disc_offset = 0; tmp.lay.ln_offset = 0; tmp.lay.ln_count = 0; tmp.lay.disc_count = 0; tmp.lay.pw1_offset = 0; tmp.lay.pw2_offset = 0; tmp.lay.cumdisc_offset = 0; tmp.lay.hash_offset = 0; tmp.lay.hash_cap = 0;
    return make_params(&tmp).off[O_COUNT];
}

int pick_threads(const isla_config& cfg) {
    int nt = cfg.block_threads;
    if (nt <= 0) {
        const int R = cfg.rollouts_per_leaf > 0 ? cfg.rollouts_per_leaf : 1;
        nt = 32;
        while (nt < R && nt < 512) nt <<= 1;
        // a batch that puts two or more trees on every SM: two 256-thread CTAs per SM overlap one tree's serial descent
        // with the other's rollouts (the 128-register kernel fits 512 threads per SM either way; C2 x 296: 184 -> 208 G env-steps/s)
        if (nt > 256 && cfg.n_trees >= 296) nt = 256;
        if (cfg.expansion == ISLA_EXPAND_VOO) nt = 512;   // voo's rejection tries are evaluated CTA-wide
        if (cfg.tree_parallel > 1) { while (nt < cfg.tree_parallel * R && nt < 512) nt <<= 1; }   // one thread per in-flight rollout
        // apw with random expansion and one rollout per leaf: runs of root-widening simulations are played one thread per
        // simulation (the widening wave of cta_search_kernel); worth a wide CTA while the trees do not fill the GPU anyway
        if (cfg.agent == ISLA_AGENT_APW && cfg.expansion == ISLA_EXPAND_RANDOM && R == 1 && cfg.n_trees <= 296) nt = 512;
    }
    nt = (nt + 31) / 32 * 32;
    return nt > 512 ? 512 : nt;
}

constexpr int kClusterRefused = 1;   // launch(): a non-portable cluster size does not fit this device (nothing was launched)

template <int ENV, typename T>
int launch(const CtaParams& p, int nt, bool scripted, cudaStream_t st) {
    const size_t smem = p.pool_in_smem ? (size_t)p.stride_s : 0;
    if (scripted) {
        if (smem) cudaFuncSetAttribute(cta_search_kernel<ENV, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cta_search_kernel<ENV, T, true><<<1, nt, smem, st>>>(p);
    } else {
        if (smem) cudaFuncSetAttribute(cta_search_kernel<ENV, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (p.cluster > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(p.n_trees * p.cluster)); cfg.blockDim = dim3((unsigned)nt);
            cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = (unsigned)p.cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            if (p.cluster > 8) {   // beyond the portable size: opt in (sm_100 schedules 16-CTA clusters inside one GPC)
                const cudaError_t ae = cudaFuncSetAttribute(cta_search_kernel<ENV, T, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                int fit = 0;
                if (ae != cudaSuccess || cudaOccupancyMaxActiveClusters(&fit, cta_search_kernel<ENV, T, false>, &cfg) != cudaSuccess || fit < 1) {
                    cudaGetLastError();
                    return kClusterRefused;   // this device / partition cannot place such a cluster: the caller falls back
                }
            }
            const cudaError_t le = cudaLaunchKernelEx(&cfg, cta_search_kernel<ENV, T, false>, p);
            if (le != cudaSuccess) return cuda_fail(le, "cta_search_kernel cluster launch");
        } else {
            cta_search_kernel<ENV, T, false><<<(unsigned)p.n_trees, nt, smem, st>>>(p);
        }
    }
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : cuda_fail(err, "cta_search_kernel launch");
}

template <typename T>
int launch_env(const isla_engine* e, const CtaParams& p, int nt, bool scripted, cudaStream_t st) {
    switch (e->cfg.env_id) {
    case ISLA_ENV_CARTPOLE: return launch<ENV_CARTPOLE, T>(p, nt, scripted, st);
    case ISLA_ENV_PENDULUM: return launch<ENV_PENDULUM, T>(p, nt, scripted, st);
    case ISLA_ENV_MOUNTAINCAR: return launch<ENV_MOUNTAINCAR, T>(p, nt, scripted, st);
    case ISLA_ENV_FROZENLAKE: return launch<ENV_FROZENLAKE, T>(p, nt, scripted, st);
    case ISLA_ENV_CURVE: return launch<ENV_CURVE, T>(p, nt, scripted, st);
    }
    set_error("unknown env %d", e->cfg.env_id);
    return ISLA_ERR_INVALID;
}

}  // namespace

int cta_fill_layout(const isla_config& cfg, isla_layout& lay) {
    const int S = env_state_dim(cfg.env_id);
    const int rb = cfg.precision == ISLA_PRECISION_F64 ? 8 : 4;
    const int R = cfg.rollouts_per_leaf > 0 ? cfg.rollouts_per_leaf : 1;
    if (R > kMaxRollSmem) { set_error("rollouts_per_leaf %d exceeds %d", R, kMaxRollSmem); return ISLA_ERR_INVALID; }
    const int env_actions = cfg.env_id == ISLA_ENV_CURVE ? cfg.action_grid0 * cfg.action_grid1 : env_n_actions(cfg.env_id);
    if ((cfg.agent == ISLA_AGENT_VANILLA || cfg.agent == ISLA_AGENT_SPW) && (cfg.n_actions < 1 || cfg.n_actions > 128 ||
                                                                           cfg.n_actions != env_actions)) {
        set_error("n_actions=%d does not match env (%d)", cfg.n_actions, env_actions); return ISLA_ERR_INVALID;
    }
    if (cfg.expansion != ISLA_EXPAND_RANDOM && cfg.agent != ISLA_AGENT_APW) { set_error("voo / genetic2 expansion are APW options"); return ISLA_ERR_INVALID; }
    if (cfg.tree_parallel > 1) {
        int P2 = 1;
        while (P2 < R) P2 <<= 1;
        if (cfg.tree_parallel > kTpMax || (long long)cfg.tree_parallel * P2 > kMaxRollSmem) {
            set_error("tree_parallel %d: at most %d simulations in flight and tree_parallel * rollouts_per_leaf <= %d", cfg.tree_parallel, kTpMax, kMaxRollSmem);
            return ISLA_ERR_INVALID;
        }
        if (cfg.cluster_ctas > 1) { set_error("tree_parallel needs cluster_ctas <= 1"); return ISLA_ERR_INVALID; }
    }
    if (cfg.cluster_ctas != 0 && cfg.cluster_ctas != 1 && cfg.cluster_ctas != 2 && cfg.cluster_ctas != 4 && cfg.cluster_ctas != 8 && cfg.cluster_ctas != 16) {
        set_error("cluster_ctas must be 0 (auto), 1, 2, 4, 8 or 16 (16: non-portable cluster size)"); return ISLA_ERR_INVALID;
    }
    if (cfg.expansion < ISLA_EXPAND_RANDOM || cfg.expansion > ISLA_EXPAND_GENETIC) { set_error("unknown expansion %d", cfg.expansion); return ISLA_ERR_INVALID; }
    if (cfg.expansion == ISLA_EXPAND_GENETIC && cfg.voo_n_samples < 1) { set_error("genetic_policy: n_samples must be >= 1"); return ISLA_ERR_INVALID; }
    if (cfg.expansion == ISLA_EXPAND_VOO && (cfg.voo_n_samples < 1 || cfg.voo_n_samples > kMaxRollSmem / 3 || cfg.voo_max_try < 1)) {
        set_error("voo: n_samples must be in [1, %d] and max_try >= 1", kMaxRollSmem / 3); return ISLA_ERR_INVALID;
    }
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    lay.kernel = ISLA_KERNEL_CTA_PER_TREE;
    lay.state_dim = S; lay.real_bytes = rb; lay.n_actions = cfg.n_actions;
    lay.action_dim = (cfg.env_id == ISLA_ENV_CURVE && cfg.action_grid0 == 0) ? 2 : 1;
    lay.s_cap = (int64_t)cfg.sim_capacity + 2; lay.a_cap = (int64_t)cfg.sim_capacity + 2;
    lay.slots_per_tree = 0; lay.rec_offset = lay.rec_stride = lay.state_offset = lay.state_stride = lay.meta_offset = lay.path_offset = 0;
    int64_t off = 0;
    lay.pool_offset = off; lay.pool_stride = pool_bytes(cfg, lay.s_cap, lay.a_cap, S, rb); off = up(off + cfg.n_trees * lay.pool_stride);
    lay.hash_cap = 0;
    if (cfg.agent == ISLA_AGENT_APW || cfg.agent == ISLA_AGENT_DPW) { lay.hash_cap = 64; while (lay.hash_cap < 2 * lay.a_cap) lay.hash_cap <<= 1; }
    lay.hash_offset = off; off = up(off + cfg.n_trees * lay.hash_cap * 16);
    lay.counters_offset = off; off = up(off + 64);
    add_table_layout(cfg, lay, off);
    lay.total_bytes = off;
    return 0;
}

int cta_set_roots(isla_engine* e, const double* roots_dev, cudaStream_t st) {
    e->sims_run = 0;
    CtaParams p = make_params(e);
    ISLA_CUDA_OK(cudaMemsetAsync(e->ws + e->lay.counters_offset, 0, 64, st));
    if (e->lay.hash_cap > 0)  // empty hash sets: state field < 0
        ISLA_CUDA_OK(cudaMemsetAsync(e->ws + e->lay.hash_offset, 0xFF, (size_t)(e->cfg.n_trees * e->lay.hash_cap * 16), st));
    const unsigned grid = (unsigned)((e->cfg.n_trees + 127) / 128);
    if (e->lay.real_bytes == 4) cta_set_roots_kernel<float><<<grid, 128, 0, st>>>(p, roots_dev, e->lay.state_dim);
    else cta_set_roots_kernel<double><<<grid, 128, 0, st>>>(p, roots_dev, e->lay.state_dim);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int cta_run(isla_engine* e, int n_sim, long long scripted_tree, const uint8_t* kinds, const double* values,
            long long trace_len, cudaStream_t st) {
    CtaParams p = make_params(e, n_sim);
    e->sims_run += n_sim;
    p.n_sim = n_sim; p.scripted_tree = scripted_tree; p.trace_kinds = kinds; p.trace_values = values; p.trace_len = trace_len;
    const bool scripted = scripted_tree >= 0;
    if (scripted && p.R != 1) { set_error("scripted runs need rollouts_per_leaf == 1"); return ISLA_ERR_INVALID; }
    if (p.noise && (p.agent == ISLA_AGENT_SPW || p.agent == ISLA_AGENT_DPW) && (p.R != 1 || p.cluster > 1 || p.tree_parallel > 1)) {
        set_error("spw / dpw on the slippery lake step the live env: its noise runs sequentially through the fit, so "
                  "rollouts_per_leaf, cluster_ctas and tree_parallel must be 1");
        return ISLA_ERR_INVALID;
    }
    const int nt = scripted ? 32 : pick_threads(e->cfg);
    auto go = [&](const CtaParams& q, int threads) {
        return e->lay.real_bytes == 8 ? launch_env<double>(e, q, threads, scripted, st) : launch_env<float>(e, q, threads, scripted, st);
    };
    // one (or a few) trees with >= 4096 rollouts per leaf: a 16-CTA cluster of 256-thread CTAs halves the warps every SM
    // has to issue for (C2 on one tree: 25.5 -> 24.2 us per simulation; results are independent of the cluster size).  16 is
    // a non-portable cluster size: if this device cannot place it, the engine stays with 8.
    const long long sms = e->sm_count > 0 ? e->sm_count : 148;
    if (!scripted && e->cfg.cluster_ctas == 0 && p.cluster == 8 && p.R >= 4096 && !e->no_cluster16 && e->cfg.block_threads <= 0 &&
        e->cfg.expansion != ISLA_EXPAND_VOO && (long long)e->cfg.n_trees * 16 * 2 <= sms) {
        CtaParams p16 = p;
        p16.cluster = 16;
        int P = 1;
        while (P < p.R) P <<= 1;
        const int slice = P / 16;
        const int rc = go(p16, nt < slice ? nt : slice);
        if (rc != kClusterRefused) return rc;
        e->no_cluster16 = true;
    }
    const int rc = go(p, nt);
    if (rc == kClusterRefused) { set_error("cluster_ctas %d: this device cannot place a cluster of that size", p.cluster); return ISLA_ERR_INVALID; }
    return rc;
}

int cta_root_stats(isla_engine* e, int64_t* na_dev, double* total_dev, cudaStream_t st) {
    if (e->cfg.n_actions < 1) { set_error("root_stats by action index needs a discrete agent; use isla_engine_root_children"); return ISLA_ERR_INVALID; }
    CtaParams p = make_params(e);
    cta_root_stats_kernel<<<(unsigned)((e->cfg.n_trees + 127) / 128), 128, 0, st>>>(p, na_dev, total_dev);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int cta_best(isla_engine* e, int scripted_pos, int32_t* best_dev, double* best_action_dev, cudaStream_t st) {
    CtaParams p = make_params(e);
    cta_best_kernel<<<(unsigned)((e->cfg.n_trees + 3) / 4), 128, 0, st>>>(p, scripted_pos, best_dev, best_action_dev);   // 4 warps = 4 trees
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int cta_root_children(isla_engine* e, long long tree, int cap, double* action, int64_t* na, double* total, int32_t* count,
                      cudaStream_t st) {
    CtaParams p = make_params(e);
    cta_root_children_kernel<<<8, 256, 0, st>>>(p, tree, cap, action, na, total, count);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int cta_depth_stats(isla_engine* e, long long tree, int cap, int32_t* dmax, double* dmean, int32_t* count, cudaStream_t st) {
    CtaParams p = make_params(e);
    cta_depth_stats_kernel<<<1, 256, 0, st>>>(p, tree, cap, dmax, dmean, count);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

void cta_pool_offsets(const isla_engine* e, long long out[24]) {
    CtaParams p = make_params(e);
    for (int i = 0; i < 24; ++i) out[i] = i <= O_COUNT ? p.off[i] : 0;
}

}  // namespace isla
