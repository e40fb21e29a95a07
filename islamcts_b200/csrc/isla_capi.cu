// C ABI of libisla_b200.so (declared in include/isla_b200.h): engine handle, dispatch between the
// thread-per-tree and CTA-per-tree kernels, the root-statistics merge and error reporting.
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <new>
#include <atomic>
#include <vector>

#include "isla_envs.cuh"
#include "isla_internal.h"

namespace isla {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return ISLA_ERR_CUDA;
}

// ln(n) and gamma**t are computed on the HOST with the C library -- the same libm calls numpy's np.log
// (utils/action_selection_functions.py:21) and CPython's pow(gamma, t) (agents/abstract_mcts.py:87) end in --
// so mathematically tied UCB scores stay bit-identical ties on the device.
void add_table_layout(const isla_config& cfg, isla_layout& lay, int64_t& off) {
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    lay.ln_count = (int64_t)cfg.sim_capacity + 8;
    lay.disc_count = (int64_t)cfg.max_depth + 2;
    lay.ln_offset = off; off = up(off + lay.ln_count * 8);
    lay.disc_offset = off; off = up(off + lay.disc_count * 8);
    lay.pw1_offset = off; off = up(off + lay.ln_count * 8);
    lay.pw2_offset = off; off = up(off + lay.ln_count * 8);
    lay.cumdisc_offset = off; off = up(off + lay.disc_count * 8);
}
// Return table of the thread-per-tree engine.  Its envs pay a constant reward r on every non-terminal step (CartPole 1,
// FrozenLake 0), so the value a simulation backs up l levels above its leaf is h^l(G_leaf), h(x) = r + gamma * x evaluated
// exactly like the recursion unwind `reward + gamma * child.build_tree_state(...)` (mcts_hash.py:126-127: one multiply,
// one add, both rounded), and G_leaf is one of max_depth + 3 values (Rewards<ENV>::leaf_row_*): a terminal reward, or
// r + gamma * rollout(t) with the t-step rollout return accumulated as in abstract_mcts.py:87.  The kernel therefore
// writes nothing per level on the way up: the next simulation reads ret[row][levels above the leaf].
// `volatile` keeps every product and sum a separately rounded double (no contraction), as in CPython.
static int build_return_table(isla_engine* e, const std::vector<double>& disc, const std::vector<double>& cum) {
    const int rows = e->cfg.max_depth + 3;
    // a path has at most sim_capacity + 1 levels; deeper columns than the table holds are continued on the device
    long long K = (long long)e->cfg.sim_capacity + 2;
    const long long cap_entries = (64ll << 20) / 8;
    if (K > 4096) K = 4096;
    while (K > 64 && K * rows > cap_entries) K /= 2;
    const bool lake = e->cfg.env_id == ISLA_ENV_FROZENLAKE;
    const double r = lake ? 0.0 : 1.0, gamma = e->cfg.gamma;
    std::vector<double> tab((size_t)rows * (size_t)K);
    for (int j = 0; j < rows; ++j) {
        volatile double g;
        if (lake) {
            if (j == 0) g = 0.0;
            else if (j == 1) g = 1.0;
            else { volatile double R = 1.0 * disc[(size_t)j - 2]; volatile double m = gamma * R; g = 0.0 + m; }
        } else {
            if (j == rows - 1) g = 0.0;
            else { volatile double m = gamma * cum[(size_t)j]; g = 1.0 + m; }
        }
        double* row = tab.data() + (size_t)j * (size_t)K;
        for (long long k = 0; k < K; ++k) {
            row[k] = g;
            volatile double m = gamma * g;
            g = r + m;
        }
    }
    if (e->ret_dev) { cudaFree(e->ret_dev); e->ret_dev = nullptr; }
    // kRetTablePad zero entries on both sides: a bulk-copied window of a round may start before row 0 / end past the last row
    ISLA_CUDA_OK(cudaMalloc((void**)&e->ret_dev, (tab.size() + 2 * kRetTablePad) * 8));
    ISLA_CUDA_OK(cudaMemset(e->ret_dev, 0, (tab.size() + 2 * kRetTablePad) * 8));
    ISLA_CUDA_OK(cudaMemcpy(e->ret_dev + kRetTablePad, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    e->ret_rows = rows; e->ret_K = (int)K;
    return 0;
}

int upload_tables(isla_engine* e) {
    std::vector<double> ln((size_t)e->lay.ln_count), disc((size_t)e->lay.disc_count);
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    for (size_t t = 0; t < disc.size(); ++t) disc[t] = std::pow(e->cfg.gamma, (double)t);
    ISLA_CUDA_OK(cudaMemcpy(e->ws + e->lay.ln_offset, ln.data(), ln.size() * 8, cudaMemcpyHostToDevice));
    ISLA_CUDA_OK(cudaMemcpy(e->ws + e->lay.disc_offset, disc.data(), disc.size() * 8, cudaMemcpyHostToDevice));
    // sum_{t<n} 1.0 * gamma**t accumulated left to right, exactly like `reward += vals[1] * pow(gamma, t)` with
    // reward 1 on every step (agents/abstract_mcts.py:87): the return of an n-step CartPole rollout
    std::vector<double> cum(disc.size());
    { double acc = 0.0; for (size_t n = 0; n < cum.size(); ++n) { cum[n] = acc; acc = acc + 1.0 * disc[n]; } }
    ISLA_CUDA_OK(cudaMemcpy(e->ws + e->lay.cumdisc_offset, cum.data(), cum.size() * 8, cudaMemcpyHostToDevice));
    // progressive-widening caps, again with the host libm pow (python: k * ns ** alpha):
    //   pw1[n] = ceil(k1 * n**alpha1)  state-node test  (mcts_action_...:62, mcts_double_...:56)
    //   pw2[n] = k * n**alpha          action-node test (mcts_state_...:97 with (k, alpha); mcts_double_...:87 with (kSpw, alphaSpw))
    std::vector<double> pw1(ln.size()), pw2(ln.size());
    const double kk = e->cfg.agent == ISLA_AGENT_SPW ? e->cfg.k1 : e->cfg.k2;
    const double al = e->cfg.agent == ISLA_AGENT_SPW ? e->cfg.alpha1 : e->cfg.alpha2;
    for (size_t n = 0; n < ln.size(); ++n) {
        pw1[n] = std::ceil(e->cfg.k1 * std::pow((double)n, e->cfg.alpha1));
        pw2[n] = kk * std::pow((double)n, al);
    }
    ISLA_CUDA_OK(cudaMemcpy(e->ws + e->lay.pw1_offset, pw1.data(), pw1.size() * 8, cudaMemcpyHostToDevice));
    ISLA_CUDA_OK(cudaMemcpy(e->ws + e->lay.pw2_offset, pw2.data(), pw2.size() * 8, cudaMemcpyHostToDevice));
    if (e->kernel == ISLA_KERNEL_THREAD_PER_TREE) {
        const int rc = build_return_table(e, disc, cum);
        if (rc) return rc;
    }
    ISLA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

static int resolve_kernel(const isla_config& c) {
    if (c.kernel == ISLA_KERNEL_THREAD_PER_TREE || c.kernel == ISLA_KERNEL_CTA_PER_TREE) return c.kernel;
    const bool tpt_ok = c.agent == ISLA_AGENT_VANILLA && env_n_actions(c.env_id) > 0 && c.rollouts_per_leaf <= 1 && c.tree_parallel <= 1;
    // narrow vanilla trees -> one thread each (the kernel spreads a small batch over many warps, see kTreesPerWarp; it is
    // at least as fast as the CTA engine even for ONE tree: CartPole nsim 1000 in 9.9 vs 12.9 ms);
    // wide (progressive widening, grids) / leaf-parallel trees -> one CTA (or cluster) each
    return tpt_ok ? ISLA_KERNEL_THREAD_PER_TREE : ISLA_KERNEL_CTA_PER_TREE;
}

static int validate(const isla_config& c) {
    if (c.env_id < 0 || c.env_id >= ENV_COUNT) { set_error("unknown env_id %d", c.env_id); return ISLA_ERR_INVALID; }
    if (c.agent < ISLA_AGENT_VANILLA || c.agent > ISLA_AGENT_DPW) { set_error("unknown agent %d", c.agent); return ISLA_ERR_INVALID; }
    if (c.n_trees < 1) { set_error("n_trees must be >= 1"); return ISLA_ERR_INVALID; }
    if (c.sim_capacity < 1) { set_error("sim_capacity must be >= 1"); return ISLA_ERR_INVALID; }
    if (c.max_depth < 0) { set_error("max_depth must be >= 0"); return ISLA_ERR_INVALID; }
    if (c.precision != ISLA_PRECISION_F32 && c.precision != ISLA_PRECISION_F64) { set_error("bad precision"); return ISLA_ERR_INVALID; }
    if (c.env_id == ISLA_ENV_CURVE) {
        const bool grid = c.action_grid0 != 0 || c.action_grid1 != 0;
        if (grid && (c.action_grid0 < 1 || c.action_grid1 < 1 || (int64_t)c.action_grid0 * c.action_grid1 > 128)) {
            set_error("DiscreteCurveEnv grid %d x %d: need 1 <= g0*g1 <= 128", c.action_grid0, c.action_grid1); return ISLA_ERR_INVALID;
        }
        if (c.agent == ISLA_AGENT_SPW || c.agent == ISLA_AGENT_DPW) {
            set_error("spw/dpw poke the observation into env.state; CurveEnv keeps its state in env.car"); return ISLA_ERR_INVALID;
        }
    }
    const int A = c.env_id == ISLA_ENV_CURVE ? c.action_grid0 * c.action_grid1 : env_n_actions(c.env_id);
    if ((c.agent == ISLA_AGENT_VANILLA || c.agent == ISLA_AGENT_SPW) && A == 0) {
        set_error("agent %d needs a discrete action space; env %d is continuous", c.agent, c.env_id); return ISLA_ERR_INVALID;
    }
    if ((c.agent == ISLA_AGENT_APW || c.agent == ISLA_AGENT_DPW) && A != 0) {
        set_error("agent %d samples a continuous Box; env %d is discrete", c.agent, c.env_id); return ISLA_ERR_INVALID;
    }
    if ((c.agent == ISLA_AGENT_SPW || c.agent == ISLA_AGENT_DPW) && c.env_id == ISLA_ENV_PENDULUM) {
        set_error("spw/dpw poke the observation into the env state; Pendulum's observation is not its state"); return ISLA_ERR_INVALID;
    }
    return 0;
}

static int fill_layout(const isla_config& cfg, isla_layout& lay) {
    memset(&lay, 0, sizeof lay);
    const int rc = validate(cfg);
    if (rc) return rc;
    return resolve_kernel(cfg) == ISLA_KERNEL_THREAD_PER_TREE ? tpt_fill_layout(cfg, lay) : cta_fill_layout(cfg, lay);
}

// Per-root sums in a FIXED order (bit-reproducible totals, unlike float atomics): the trees of a root are summed by one
// CTA -- every thread adds the root's trees first + tid, first + tid + T, ... in ascending order, then a pairwise tree over
// the T partial sums in shared memory.  `group` (root of each tree) must be ASCENDING, i.e. the trees of a root are
// contiguous (what sharding.plan() lays out); the CTA finds its range by binary search.  group == nullptr: tree i = root i.
__global__ void __launch_bounds__(256) merge_root_stats_kernel(const int64_t* na, const double* total, const int32_t* group, long long n_trees,
                                                               int A, long long* na_out, double* total_out) {
    extern __shared__ double red[];   // [blockDim.x] totals | [blockDim.x] counts (as int64)
    long long* const redn = reinterpret_cast<long long*>(red + blockDim.x);
    const long long root = blockIdx.x;
    long long first = root, last = root + 1;
    if (group) {
        auto lower = [&](long long key) { long long lo = 0, hi = n_trees; while (lo < hi) { const long long mid = (lo + hi) >> 1; if (group[mid] < key) lo = mid + 1; else hi = mid; } return lo; };
        first = lower(root); last = lower(root + 1);
    }
    if (last > n_trees) last = n_trees;
    for (int a = 0; a < A; ++a) {
        double t = 0.0; long long n = 0;
        for (long long tree = first + threadIdx.x; tree < last; tree += blockDim.x) { t = __dadd_rn(t, total[tree * A + a]); n += na[tree * A + a]; }
        red[threadIdx.x] = t; redn[threadIdx.x] = n;
        __syncthreads();
        for (int s_ = blockDim.x >> 1; s_ > 0; s_ >>= 1) {
            if ((int)threadIdx.x < s_) { red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + s_]); redn[threadIdx.x] += redn[threadIdx.x + s_]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) { total_out[root * A + a] = red[0]; na_out[root * A + a] = redn[0]; }
        __syncthreads();
    }
}

// The same sums straight from the thread-per-tree engine's root blocks, for CONTIGUOUS equal groups (tree i belongs to
// root i / per_root, what sharding.plan() lays out), packed for ONE all-reduce: out[root][a] = {na as double, total}.
// Visit counts are far below 2**53, so the float64 sum of counts is exact.
__global__ void __launch_bounds__(256) tpt_root_merge_kernel(const char* blocks, long long slots_per_tree, int A, long long per_root,
                                                             double* out) {
    extern __shared__ double red[];   // [blockDim.x] totals | [blockDim.x] counts
    double* const redn = red + blockDim.x;
    const long long root = blockIdx.x;
    for (int a = 0; a < A; ++a) {
        double t = 0.0, n = 0.0;
        for (long long i = threadIdx.x; i < per_root; i += blockDim.x) {
            const char* b = blocks + (root * per_root + i) * slots_per_tree * 16 + (long long)A * 16;   // block 1 = the root's edges
            t = __dadd_rn(t, reinterpret_cast<const double*>(b)[a]);
            n += (double)reinterpret_cast<const int32_t*>(b + A * 8)[a];
        }
        red[threadIdx.x] = t; redn[threadIdx.x] = n;
        __syncthreads();
        for (int s_ = blockDim.x >> 1; s_ > 0; s_ >>= 1) {
            if ((int)threadIdx.x < s_) { red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + s_]); redn[threadIdx.x] += redn[threadIdx.x + s_]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) { out[(root * A + a) * 2] = redn[0]; out[(root * A + a) * 2 + 1] = red[0]; }
        __syncthreads();
    }
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace isla

using namespace isla;

extern "C" {

const char* isla_last_error(void) { return g_err; }
int isla_abi_version(void) { return ISLA_ABI_VERSION; }
int isla_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int64_t isla_engine_workspace_bytes(const isla_config* cfg) {
    if (!cfg) { set_error("null config"); return ISLA_ERR_INVALID; }
    isla_layout lay;
    const int rc = fill_layout(*cfg, lay);
    return rc ? rc : lay.total_bytes;
}

int isla_engine_create(const isla_config* cfg, void* workspace_dev, int64_t workspace_bytes, isla_engine** out) {
    if (!cfg || !out) { set_error("null argument"); return ISLA_ERR_INVALID; }
    isla_layout lay;
    int rc = fill_layout(*cfg, lay);
    if (rc) return rc;
    if (isla_device_count() < 1) { set_error("no CUDA device: libisla_b200 has no CPU fallback"); return ISLA_ERR_CUDA; }
    if (!workspace_dev || workspace_bytes < lay.total_bytes) {
        set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)lay.total_bytes);
        return ISLA_ERR_INVALID;
    }
    if (((uintptr_t)workspace_dev & 255u) != 0) { set_error("workspace must be 256-byte aligned"); return ISLA_ERR_INVALID; }
    isla_engine* e = new (std::nothrow) isla_engine();
    if (!e) { set_error("out of host memory"); return ISLA_ERR_INVALID; }
    e->cfg = *cfg; e->lay = lay; e->ws = (char*)workspace_dev; e->kernel = lay.kernel; e->roots_set = false;
    int dev = 0;
    ISLA_CUDA_OK(cudaGetDevice(&dev));
    ISLA_CUDA_OK(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, dev));
    rc = upload_tables(e);
    if (rc) { delete e; return rc; }
    *out = e;
    return ISLA_OK;
}

int isla_engine_destroy(isla_engine* e) {
    if (e && e->heur_dev) cudaFree(e->heur_dev);
    if (e && e->noise_dev) cudaFree(e->noise_dev);
    if (e && e->ret_dev) cudaFree(e->ret_dev);
    if (e && e->reroot_tmp) cudaFree(e->reroot_tmp);
    delete e;
    return ISLA_OK;
}

int isla_engine_set_transition_noise(isla_engine* e, const double* u_host, int32_t n) {
    if (!e) { set_error("null engine"); return ISLA_ERR_INVALID; }
    if (!u_host && !e->noise_dev) return ISLA_OK;          // deterministic lake, nothing to clear
    ISLA_CUDA_OK(cudaDeviceSynchronize());
    if (e->noise_dev && (!u_host || n != e->noise_n)) { cudaFree(e->noise_dev); e->noise_dev = nullptr; e->noise_n = 0; }
    if (!u_host) return ISLA_OK;
    if (e->cfg.env_id != ISLA_ENV_FROZENLAKE || n < 1) { set_error("transition noise: FrozenLake (is_slippery=True) only, n >= 1"); return ISLA_ERR_INVALID; }
    if (!e->noise_dev) ISLA_CUDA_OK(cudaMalloc((void**)&e->noise_dev, (size_t)n * sizeof(double)));
    ISLA_CUDA_OK(cudaMemcpy(e->noise_dev, u_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    e->noise_n = n;
    return ISLA_OK;
}

int isla_engine_set_rollout_heuristic(isla_engine* e, const double* table_host, int32_t n_states, int32_t n_actions) {
    if (!e) { set_error("null engine"); return ISLA_ERR_INVALID; }
    if (!table_host) {
        if (e->heur_dev) { ISLA_CUDA_OK(cudaDeviceSynchronize()); cudaFree(e->heur_dev); e->heur_dev = nullptr; }
        return ISLA_OK;
    }
    if (e->cfg.env_id != ISLA_ENV_FROZENLAKE || n_states != 16 || n_actions != 4) {
        set_error("grid_policy: only FrozenLake (16 states x 4 actions) has a heuristic rollout on the device"); return ISLA_ERR_INVALID;
    }
    if (!e->heur_dev) ISLA_CUDA_OK(cudaMalloc((void**)&e->heur_dev, 16 * 4 * sizeof(double)));
    ISLA_CUDA_OK(cudaMemcpy(e->heur_dev, table_host, 16 * 4 * sizeof(double), cudaMemcpyHostToDevice));
    return ISLA_OK;
}

int isla_engine_layout(const isla_engine* e, isla_layout* out) {
    if (!e || !out) { set_error("null argument"); return ISLA_ERR_INVALID; }
    *out = e->lay;
    return ISLA_OK;
}

int isla_engine_set_roots(isla_engine* e, const double* roots_dev, void* stream) {
    if (!e || !roots_dev) { set_error("null argument"); return ISLA_ERR_INVALID; }
    const int rc = e->kernel == ISLA_KERNEL_THREAD_PER_TREE ? tpt_set_roots(e, roots_dev, (cudaStream_t)stream)
                                                            : cta_set_roots(e, roots_dev, (cudaStream_t)stream);
    if (rc == 0) e->roots_set = true;
    if (rc == 0) count_launch(1);
    return rc;
}

int isla_engine_run(isla_engine* e, int32_t n_sim, void* stream) {
    if (!e) { set_error("null engine"); return ISLA_ERR_INVALID; }
    if (!e->roots_set) { set_error("isla_engine_set_roots must be called before run"); return ISLA_ERR_INVALID; }
    if (n_sim < 0) { set_error("n_sim < 0"); return ISLA_ERR_INVALID; }
    if (n_sim == 0) return ISLA_OK;
    const int rc_ = e->kernel == ISLA_KERNEL_THREAD_PER_TREE ? tpt_run(e, n_sim, -1, nullptr, nullptr, 0, (cudaStream_t)stream)
                                                    : cta_run(e, n_sim, -1, nullptr, nullptr, 0, (cudaStream_t)stream);
    if (rc_ == 0) count_launch(1);
    return rc_;
}

int isla_engine_run_scripted(isla_engine* e, int64_t tree, int32_t n_sim, const uint8_t* kinds_dev, const double* values_dev,
                             int64_t trace_len, void* stream) {
    if (!e) { set_error("null engine"); return ISLA_ERR_INVALID; }
    if (!e->roots_set) { set_error("isla_engine_set_roots must be called before run"); return ISLA_ERR_INVALID; }
    if (tree < 0 || tree >= e->cfg.n_trees) { set_error("tree index out of range"); return ISLA_ERR_INVALID; }
    if (trace_len > 0 && (!kinds_dev || !values_dev)) { set_error("null trace"); return ISLA_ERR_INVALID; }
    if (n_sim <= 0) return ISLA_OK;
    const int rc_ = e->kernel == ISLA_KERNEL_THREAD_PER_TREE ? tpt_run(e, n_sim, tree, kinds_dev, values_dev, trace_len, (cudaStream_t)stream)
                                                    : cta_run(e, n_sim, tree, kinds_dev, values_dev, trace_len, (cudaStream_t)stream);
    if (rc_ == 0) count_launch(1);
    return rc_;
}

int isla_engine_root_stats(isla_engine* e, int64_t* na_dev, double* total_dev, void* stream) {
    if (!e || !na_dev || !total_dev) { set_error("null argument"); return ISLA_ERR_INVALID; }
    const int rc_ = e->kernel == ISLA_KERNEL_THREAD_PER_TREE ? tpt_root_stats(e, na_dev, total_dev, (cudaStream_t)stream)
                                                    : cta_root_stats(e, na_dev, total_dev, (cudaStream_t)stream);
    if (rc_ == 0) count_launch(1);
    return rc_;
}

int isla_engine_best(isla_engine* e, int32_t scripted_pos, int32_t* best_dev, double* best_action_dev, void* stream) {
    if (!e || !best_dev) { set_error("null argument"); return ISLA_ERR_INVALID; }
    const int rc_ = e->kernel == ISLA_KERNEL_THREAD_PER_TREE ? tpt_best(e, scripted_pos, best_dev, best_action_dev, (cudaStream_t)stream)
                                                    : cta_best(e, scripted_pos, best_dev, best_action_dev, (cudaStream_t)stream);
    if (rc_ == 0) count_launch(1);
    return rc_;
}

int isla_engine_root_children(isla_engine* e, int64_t tree, int32_t cap, double* action_dev, int64_t* na_dev, double* total_dev,
                              int32_t* count_dev, void* stream) {
    if (!e || !count_dev || (cap > 0 && (!action_dev || !na_dev || !total_dev))) { set_error("null argument"); return ISLA_ERR_INVALID; }
    if (tree < 0 || tree >= e->cfg.n_trees) { set_error("tree index out of range"); return ISLA_ERR_INVALID; }
    if (e->kernel != ISLA_KERNEL_CTA_PER_TREE) { set_error("root_children: CTA-per-tree engine only (use root_stats)"); return ISLA_ERR_INVALID; }
    const int rc_ = cta_root_children(e, tree, cap, action_dev, na_dev, total_dev, count_dev, (cudaStream_t)stream);
    if (rc_ == 0) count_launch(1);
    return rc_;
}

int isla_engine_depth_stats(isla_engine* e, int64_t tree, int32_t cap, int32_t* depth_max_dev, double* depth_mean_dev,
                            int32_t* count_dev, void* stream) {
    if (!e || !count_dev || cap < 0 || (cap > 0 && (!depth_max_dev || !depth_mean_dev))) { set_error("bad argument"); return ISLA_ERR_INVALID; }
    if (tree < 0 || tree >= e->cfg.n_trees) { set_error("tree index out of range"); return ISLA_ERR_INVALID; }
    if (!e->roots_set) { set_error("isla_engine_set_roots must be called first"); return ISLA_ERR_INVALID; }
    const int rc_ = e->kernel == ISLA_KERNEL_THREAD_PER_TREE
               ? tpt_depth_stats(e, tree, cap, depth_max_dev, depth_mean_dev, count_dev, (cudaStream_t)stream)
               : cta_depth_stats(e, tree, cap, depth_max_dev, depth_mean_dev, count_dev, (cudaStream_t)stream);
    if (rc_ == 0) count_launch(1);
    return rc_;
}

int isla_engine_pool_offsets(const isla_engine* e, int64_t out[24]) {
    if (!e || !out) { set_error("null argument"); return ISLA_ERR_INVALID; }
    if (e->kernel != ISLA_KERNEL_CTA_PER_TREE) { set_error("pool_offsets: CTA-per-tree engine only"); return ISLA_ERR_INVALID; }
    long long tmp[24];
    cta_pool_offsets(e, tmp);
    for (int i = 0; i < 24; ++i) out[i] = tmp[i];
    return ISLA_OK;
}

int isla_engine_reseed(isla_engine* e, uint64_t seed, uint64_t tree_id_base) {
    if (!e) { set_error("null engine"); return ISLA_ERR_INVALID; }
    e->cfg.seed = seed; e->cfg.tree_id_base = tree_id_base; e->roots_set = false;
    return ISLA_OK;
}

int isla_engine_counters(isla_engine* e, int64_t out_host[8], void* stream) {
    if (!e || !out_host) { set_error("null argument"); return ISLA_ERR_INVALID; }
    ISLA_CUDA_OK(cudaMemcpyAsync(out_host, e->ws + e->lay.counters_offset, 64, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ISLA_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return ISLA_OK;
}

int isla_merge_root_stats(const int64_t* na_dev, const double* total_dev, const int32_t* group_dev, int64_t n_trees,
                          int32_t n_actions, int64_t n_roots, int64_t* na_out_dev, double* total_out_dev, void* stream) {
    if (!na_dev || !total_dev || !na_out_dev || !total_out_dev || n_trees < 0 || n_actions < 1 || n_roots < 1) {
        set_error("bad argument"); return ISLA_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // one CTA per root; threads = trees per root rounded up to a power of two (32..256)
    long long per = n_roots > 0 ? (n_trees + n_roots - 1) / n_roots : 1;
    int threads = 32;
    while (threads < 256 && threads < per) threads <<= 1;
    merge_root_stats_kernel<<<(unsigned)n_roots, threads, (size_t)threads * 16, st>>>(na_dev, total_dev, group_dev, n_trees, n_actions,
                                                                                     (long long*)na_out_dev, total_out_dev);
    ISLA_CUDA_OK(cudaGetLastError());
    count_launch(1);
    return ISLA_OK;
}

int isla_engine_root_merge(isla_engine* e, int64_t n_roots, double* packed_out_dev, void* stream) {
    if (!e || !packed_out_dev || n_roots < 1) { set_error("bad argument"); return ISLA_ERR_INVALID; }
    if (e->kernel != ISLA_KERNEL_THREAD_PER_TREE) { set_error("root_merge: thread-per-tree engine only (use root_stats + isla_merge_root_stats)"); return ISLA_ERR_INVALID; }
    if (e->cfg.n_trees % n_roots) { set_error("root_merge: n_trees must be a multiple of n_roots (contiguous equal groups)"); return ISLA_ERR_INVALID; }
    const long long per = e->cfg.n_trees / n_roots;
    int threads = 32;
    while (threads < 256 && threads < per) threads <<= 1;
    tpt_root_merge_kernel<<<(unsigned)n_roots, threads, (size_t)threads * 16, (cudaStream_t)stream>>>(
        e->ws + e->lay.rec_offset, e->lay.slots_per_tree, e->lay.n_actions, per, packed_out_dev);
    ISLA_CUDA_OK(cudaGetLastError());
    count_launch(1);
    return ISLA_OK;
}

int isla_engine_reroot(isla_engine* e, const int32_t* action_dev, const uint8_t* alive_dev, int32_t n_sim_next, void* stream) {
    if (!e || !action_dev) { set_error("null argument"); return ISLA_ERR_INVALID; }
    if (e->kernel != ISLA_KERNEL_THREAD_PER_TREE) { set_error("reroot: thread-per-tree engine only"); return ISLA_ERR_INVALID; }
    if (!e->roots_set) { set_error("isla_engine_set_roots must be called first"); return ISLA_ERR_INVALID; }
    if (n_sim_next < 0) { set_error("n_sim_next < 0"); return ISLA_ERR_INVALID; }
    const int rc = tpt_reroot(e, action_dev, alive_dev, n_sim_next, (cudaStream_t)stream);
    if (rc == 0) count_launch(1);
    return rc;
}

int64_t isla_kernel_launches(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
