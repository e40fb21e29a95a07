/* isla_b200.h -- C ABI of the B200-native MCTS engine (libisla_b200.so).
 *
 * Drop-in boundary for ONE hot path of LorenzoBonanni/IslaMcts: the simulate loop inside
 * Mcts / MctsHash / MctsApw / MctsSpw / MctsDpw .fit()  (UCB1 selection, progressive-widening
 * expansion, random rollout with discounted return, value backup).  The reference has no FFI; its
 * boundary is the Python class API, so each entry point cites the Python it replaces
 * (paths relative to the reference tree).  Host binding: islamcts_b200/_lib.py (ctypes);
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every `*_dev` pointer is DEVICE memory owned by the caller;
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); no call synchronises the device unless its comment says so;
 *   - every function returns 0 on success or a negative isla_status; isla_last_error() gives text;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     ISLA_ERR_CUDA.
 */
#ifndef ISLA_B200_H
#define ISLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISLA_ABI_VERSION 3

typedef enum {
    ISLA_OK = 0,
    ISLA_ERR_INVALID = -1,      /* bad argument / unsupported combination (Python: NotImplementedError) */
    ISLA_ERR_CUDA = -2,         /* CUDA runtime error, text in isla_last_error() */
    ISLA_ERR_CAPACITY = -3,     /* node capacity of the device tree exhausted */
    ISLA_ERR_TRACE = -4         /* scripted run: trace exhausted / kind mismatch / value out of range */
} isla_status;

/* environments: the gym ids the Python layer recognises, and the reference's own CurveEnv / DiscreteCurveEnv
 * (islaMcts/environments/curve_env.py:59-199; state = x, y, velocity, angle, n_step) */
enum { ISLA_ENV_CARTPOLE = 0, ISLA_ENV_PENDULUM = 1, ISLA_ENV_MOUNTAINCAR = 2, ISLA_ENV_FROZENLAKE = 3, ISLA_ENV_CURVE = 4 };
/* agents: islaMcts/agents/{mcts,mcts_hash}.py ; mcts_action_... ; mcts_state_... ; mcts_double_... */
enum { ISLA_AGENT_VANILLA = 0, ISLA_AGENT_APW = 1, ISLA_AGENT_SPW = 2, ISLA_AGENT_DPW = 3 };
/* rollout budget: MctsHash depth accounting (mcts_hash.py:34,116,126) or the zero-length rollouts the
 * other four agents perform at HEAD (mcts.py:27 + abstract_mcts.py:83) */
enum { ISLA_BUDGET_TO_MAX_DEPTH = 0, ISLA_BUDGET_ZERO = 1 };
enum { ISLA_EXPAND_RANDOM = 0, ISLA_EXPAND_VOO = 1, ISLA_EXPAND_GENETIC2 = 2, ISLA_EXPAND_GENETIC = 3 };   /* utils/action_selection_functions.py:46, :139, :104, :50 */
enum { ISLA_PRECISION_F32 = 0, ISLA_PRECISION_F64 = 1 };   /* env arithmetic; tree statistics are always f64 */
enum { ISLA_KERNEL_AUTO = 0, ISLA_KERNEL_THREAD_PER_TREE = 1, ISLA_KERNEL_CTA_PER_TREE = 2 };
/* decision-trace kinds of the scripted (parity) mode */
enum { ISLA_K_CHOICE = 0, ISLA_K_SAMPLE_D = 1, ISLA_K_SAMPLE_C = 2, ISLA_K_CHOICES = 3, ISLA_K_U01 = 4, ISLA_K_UNIFORM = 5 };

/* Replaces the MctsParameters / PwParameters / DpwParameters records
 * (agents/parameters/mcts_parameters.py:8-31, pw_parameters.py:7-11, dpw_parameters.py:6-13). */
typedef struct {
    int32_t env_id;             /* ISLA_ENV_*   (param.env, identified by the Python layer) */
    int32_t agent;              /* ISLA_AGENT_* */
    int32_t n_actions;          /* param.n_actions (discrete agents) */
    int32_t max_depth;          /* param.max_depth */
    int32_t budget_mode;        /* ISLA_BUDGET_* */
    int32_t rollouts_per_leaf;  /* build extension: mean of R rollouts backed up as ONE visit */
    int32_t expansion;          /* ISLA_EXPAND_*  (param.action_expansion_function) */
    int32_t voo_n_samples;      /* voo: samples per expansion; genetic_policy: n_samples */
    int32_t voo_max_try;
    int32_t precision;          /* ISLA_PRECISION_* */
    int32_t kernel;             /* ISLA_KERNEL_*; AUTO = thread-per-tree for vanilla trees on CartPole / FrozenLake with
                                 * rollouts_per_leaf <= 1, CTA-per-tree otherwise */
    int32_t block_threads;      /* CTA-per-tree kernel: threads per tree (multiple of 32); 0 = auto */
    int32_t sim_capacity;       /* max cumulative simulations per tree the workspace is sized for */
    int32_t tuning;             /* experiment flags for A/B timing and tests; 0 = defaults */
    int32_t action_grid0;       /* ISLA_ENV_CURVE: DiscreteCurveEnv([g0, g1]) = g0 accelerations x g1 steering angles */
    int32_t action_grid1;       /*   (curve_env.py:184-199; n_actions = g0*g1 <= 128); both 0 = the continuous 2-D CurveEnv box */
    int32_t cluster_ctas;       /* CTA-per-tree kernel: thread blocks (one thread-block cluster) that share one tree's leaf-parallel
                                 * rollouts; 0 = auto (up to 8 when rollouts_per_leaf > 512; 16 -- a non-portable cluster size -- for a few
                                 * trees with >= 4096 rollouts per leaf where the device can place it, else 8), or 1, 2, 4, 8, 16 */
    int32_t tree_parallel;      /* build extension, CTA-per-tree kernel: > 1 = that many simulations of ONE tree in flight at once
                                 * (virtual loss: each descent provisionally credits its leaf with virtual_value, the rollouts of
                                 * the batch run concurrently, the credits are then corrected in order); 0 / 1 = the reference's
                                 * sequential loop (agents/mcts_hash.py:30-34).  Statistics differ from the sequential search. */
    double C;                   /* param.C */
    double gamma;               /* param.gamma */
    double alpha1, k1;          /* PwParameters.alpha/k ; DpwParameters.alphaApw/kApw */
    double alpha2, k2;          /* DpwParameters.alphaSpw/kSpw */
    double epsilon;             /* voo epsilon */
    uint64_t seed;              /* Philox key */
    uint64_t tree_id_base;      /* global id of tree 0 of this engine (rank offset under sharding) */
    int64_t n_trees;            /* independent trees held by this engine */
    double virtual_value;       /* tree_parallel > 1: the rollout return a leaf is credited with until its real rollouts arrive */
} isla_config;

typedef struct isla_engine isla_engine;

/* Byte layout of the workspace so that host code can decode the device tree (lazy node proxies). */
typedef struct {
    int32_t kernel;             /* resolved ISLA_KERNEL_* */
    int32_t state_dim;
    int32_t real_bytes;         /* 4 or 8 */
    int32_t n_actions;
    int32_t action_dim;         /* reals per stored action: 1, or 2 for the continuous CurveEnv box */
    int32_t reserved0;
    int64_t slots_per_tree;     /* thread-per-tree: edge slots per tree = blocks * A; slot = block * A + action */
    int64_t rec_offset, rec_stride;         /* thread-per-tree: node blocks {double total[A]; int32 na[A]; uint32 link[A]} */
    int64_t state_offset, state_stride;     /* env states, state_dim reals per slot */
    int64_t gfirst_offset, gfirst_stride;   /* double per slot: return backed up when the child state was created */
    int64_t meta_offset;                    /* per-tree int32[8]: alloc, sims, rng words consumed, fault, trace pos, pending */
    int64_t path_offset, pathg_offset;      /* PV cache: per tree, chunks of 32 path levels {sibling statistics | own totals | pending returns G | edge ids} */
    int64_t pvown_offset;                   /* (pathg_offset and pvown_offset equal path_offset: those rows live inside the chunks) */
    int64_t counters_offset;                /* int64[8] device counters */
    int64_t ln_offset, ln_count;            /* double ln(n), n < ln_count  (host libm log, like numpy)  */
    int64_t disc_offset, disc_count;        /* double gamma**t, t < disc_count (host libm pow, like CPython) */
    int64_t pw1_offset, pw2_offset;         /* double[ln_count]: ceil(k1*n**alpha1) and k*n**alpha widening caps (host libm) */
    int64_t cumdisc_offset;                 /* double[disc_count]: sum_{t<n} gamma**t accumulated like `reward += 1.0 * pow(gamma, t)` */
    /* CTA-per-tree (general) pools, offsets of the per-tree blocks and their strides */
    int64_t s_cap, a_cap;                   /* state / action node capacity per tree */
    int64_t pool_offset, pool_stride;       /* bytes per tree of the general node pool */
    int64_t hash_offset, hash_cap;          /* apw/dpw: per-tree action hash set, hash_cap 16-byte entries per tree (0 = none) */
    int64_t total_bytes;
} isla_layout;

const char* isla_last_error(void);
int isla_abi_version(void);
/* number of visible CUDA devices (0 without a GPU); never fails */
int isla_device_count(void);

/* Sizing + creation.  The workspace is caller-owned device memory (e.g. a torch uint8 tensor),
 * >= isla_engine_workspace_bytes(cfg), 256-byte aligned. */
int64_t isla_engine_workspace_bytes(const isla_config* cfg);
int isla_engine_create(const isla_config* cfg, void* workspace_dev, int64_t workspace_bytes, isla_engine** out);
int isla_engine_destroy(isla_engine* e);
int isla_engine_layout(const isla_engine* e, isla_layout* out);

/* Agent(param) + param.root_data / param.env:  root env states, double [n_trees][state_dim]
 * (FrozenLake: the cell index as a double).  Clears the trees (replaces constructing a new agent,
 * sweep.py:186-189). */
int isla_engine_set_roots(isla_engine* e, const double* roots_dev, void* stream);

/* param.rollout_selection_fn = grid_policy(prior_knowledge, n_actions) (utils/action_selection_functions.py:258-282;
 * examples/example_lake.py:56-66): rollouts on ISLA_ENV_FROZENLAKE take an arg-min of table_host[cell][action] (HOST
 * doubles, [16][4], lower = better), ties broken uniformly.  NULL restores env.action_space.sample().  Synchronous. */
int isla_engine_set_rollout_heuristic(isla_engine* e, const double* table_host, int32_t n_states, int32_t n_actions);

/* FrozenLake-v1 with is_slippery=True as the reference runs it: every simulation steps a pickled clone of param.env
 * INCLUDING the env's generator state (utils/mcts_utils.py:8-16), so inside one fit() the k-th env.step of every
 * simulation draws the same number.  u_host[k] (HOST doubles, n >= the deepest step index a simulation can reach) is
 * that sequence, e.g. copy.deepcopy(env.np_random).random(n); the action carried out at step k is (a - 1 + i) mod 4
 * with i = argmax(cumsum([1/3, 1/3, 1/3]) > u[k]) (gym's categorical_sample).  NULL = deterministic.  Synchronous. */
int isla_engine_set_transition_noise(isla_engine* e, const double* u_host, int32_t n);

/* fit(): the `for s in range(n_sim)` loop (mcts.py:24-27, mcts_hash.py:30-34, mcts_action_...:25-30,
 * mcts_state_...:23-26, mcts_double_...:24-27) for all trees.  Calling it again continues the trees. */
int isla_engine_run(isla_engine* e, int32_t n_sim, void* stream);

/* Parity mode: tree `tree` consumes a recorded decision trace (kinds uint8, values double, device). */
int isla_engine_run_scripted(isla_engine* e, int64_t tree, int32_t n_sim, const uint8_t* kinds_dev,
                             const double* values_dev, int64_t trace_len, void* stream);

/* End of fit(): q = total/na per root child and the uniformly random arg-max
 * (mcts_hash.py:42-50).  Discrete agents: na/total are [n_trees][n_actions] by ACTION index.
 * best_dev[n_trees] = chosen action index.  scripted_pos >= 0 replaces the tie-break draw. */
int isla_engine_root_stats(isla_engine* e, int64_t* na_dev, double* total_dev, void* stream);
/* best_rank_dev[n_trees]: position of the chosen child in root.actions (insertion order; spw/dpw: after the
 * persistent key-sorted re-ordering of mcts_state_...:29 / mcts_double_...:30); best_action_dev[n_trees][action_dim]
 * (optional, may be NULL): its action value (discrete agents: the action index). */
int isla_engine_best(isla_engine* e, int32_t scripted_pos, int32_t* best_rank_dev, double* best_action_dev, void* stream);
/* Root children of one tree in root.actions order (continuous-action agents: apw/dpw): action value
 * (action_dev[cap][action_dim]), ActionNode.na, ActionNode.total; count_dev[0] = number of children (may exceed cap). */
int isla_engine_root_children(isla_engine* e, int64_t tree, int32_t cap, double* action_dev, int64_t* na_dev,
                              double* total_dev, int32_t* count_dev, void* stream);
/* Post-fit tree analytics of the sweep loop (sweep.py:199-200): for every root action node of one tree, in root.actions
 * order, AbstractActionNode.get_depth_max(0) (agents/abstract_mcts.py:190-200) and get_depth_mean(0, True) (:202-209):
 * the deepest state level below it and the mean level of its leaf states.  depth_max_dev int32[cap], depth_mean_dev
 * double[cap], count_dev[0] = number of root children (may exceed cap).  Uses per-tree scratch of the workspace: do not
 * run it concurrently with a search on the same engine. */
int isla_engine_depth_stats(isla_engine* e, int64_t tree, int32_t cap, int32_t* depth_max_dev, double* depth_mean_dev,
                            int32_t* count_dev, void* stream);
/* CTA-per-tree engine: byte offsets of the SoA arrays inside one tree's pool, for host-side decoding
 * (order: s_total, s_term_reward, s_ns, s_flags, s_coff, s_ccap, s_ccnt, s_state, a_total, a_value, a_na,
 *  a_visits, a_child, ipool, path_s, path_a, path_r, path_flag, meta, path_g, a_value2, end). */
int isla_engine_pool_offsets(const isla_engine* e, int64_t out[24]);

/* New Philox key / tree-id base for the searches that follow (takes effect at the next isla_engine_set_roots): the
 * episode loop builds a fresh tree per real step (sweep.py:186) and gives every step its own streams without
 * re-creating the engine. */
int isla_engine_reseed(isla_engine* e, uint64_t seed, uint64_t tree_id_base);

/* Real-step half of the episode loop (sweep.py:190,202-213) for n episodes at once, no host round trip: applies
 * actions_dev[n][action_dim] (ISLA_ENV_CURVE with a grid: the DiscreteCurveEnv index) to states_dev[n][state_dim] in
 * float64 for every episode with alive_dev[i] != 0, adds the reward to total_reward_dev[i], counts the step in
 * n_steps_dev[i] and clears alive_dev[i] when the env terminates or max_steps is reached; finished episodes are frozen.
 * timeout_reward: the reward booked for the step that reaches max_steps (the reference books -1000, sweep.py:206-209);
 * NaN = the env's own reward. */
int isla_episode_step(int32_t env_id, int32_t action_grid0, int32_t action_grid1, double* states_dev,
                      const double* actions_dev, int64_t n, int32_t max_steps, double timeout_reward, uint8_t* alive_dev,
                      double* total_reward_dev, int32_t* n_steps_dev, void* stream);

/* Device counters int64[8]: [0] env steps executed, [1] rollout steps, [2] tree levels visited,
 * [3] simulations, [4] nodes created, [5] faults.  SYNCHRONISES the stream. */
int isla_engine_counters(isla_engine* e, int64_t out_host[8], void* stream);

/* Root-parallel merge (north_star (4); replaces the reference's only parallel construct, joblib over whole
 * experiments, islaMcts/experiment_runner.py:305-306): per-root sums over the trees that share a root, in a FIXED
 * order (bit-reproducible totals).  group_dev[n_trees] = root id in [0, n_roots), ASCENDING (the trees of a root are
 * contiguous); NULL = tree i is root i.  Outputs [n_roots][n_actions].  The caller then all-reduces over NCCL. */
int isla_merge_root_stats(const int64_t* na_dev, const double* total_dev, const int32_t* group_dev,
                          int64_t n_trees, int32_t n_actions, int64_t n_roots, int64_t* na_out_dev,
                          double* total_out_dev, void* stream);

/* root_stats + merge in ONE launch for the thread-per-tree engine, packed for ONE all-reduce: tree i belongs to root
 * i / (n_trees / n_roots); packed_out_dev[n_roots][n_actions][2] = {na as float64 (exact: counts << 2**53), total}.
 * q = total / na and the arg-max follow on every rank (mcts_hash.py:42-50). */
int isla_engine_root_merge(isla_engine* e, int64_t n_roots, double* packed_out_dev, void* stream);

/* Tree reuse between real steps (build extension; the reference builds a new agent -- a fresh tree -- for every real step,
 * sweep.py:186): the subtree below root action action_dev[tree] of every tree (alive_dev[tree] != 0, or all if NULL)
 * becomes the tree; its statistics are kept bit for bit, the chosen child's state becomes the root state, and the next
 * isla_engine_run continues from there.  A subtree that would not leave room in the workspace for the n_sim_next
 * simulations of the coming step (node blocks, or visit counts beyond the ln(n) table: both sized by sim_capacity) is
 * dropped and the search restarts from the bare child state, as the per-step rebuild would.  Thread-per-tree engine. */
int isla_engine_reroot(isla_engine* e, const int32_t* action_dev, const uint8_t* alive_dev, int32_t n_sim_next, void* stream);

/* Kernels this library has launched in this process (bench.py's gpu_launches is a difference of two readings). */
int64_t isla_kernel_launches(void);

/* env.step over a batch (parity of the dynamics): states [n][state_dim], actions [n][action_dim] (discrete:
 * index as a real; ISLA_ENV_CURVE: (acceleration, steering) -- DiscreteCurveEnv indices are decoded by the caller),
 * precision selects float/double buffers for states/actions/rewards. */
int isla_env_step(int32_t env_id, int32_t precision, const void* states_dev, const void* actions_dev, int64_t n,
                  void* next_states_dev, void* rewards_dev, uint8_t* terminated_dev, void* stream);

/* AbstractStateNode.rollout (abstract_mcts.py:72-91) over a batch: trajectory i starts at
 * states[i] (or, if states_dev == NULL, at the gym reset distribution, Philox reset stream, index i),
 * plays uniformly random actions (stream (seed, tree=i, sim, j=0)) for at most `budget` steps,
 * returns sum r*gamma^t (double) and the steps taken.  action_grid0/1: the config fields of the same name; both 0
 * unless ISLA_ENV_CURVE is to sample the DiscreteCurveEnv grid. */
int isla_rollout(int32_t env_id, int32_t precision, const void* states_dev, int64_t n, int32_t budget, double gamma,
                 uint64_t seed, uint32_t sim, int32_t action_grid0, int32_t action_grid1, double* returns_dev,
                 int32_t* steps_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISLA_B200_H */
