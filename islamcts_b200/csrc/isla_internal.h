// Internal declarations shared by the translation units of libisla_b200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/isla_b200.h"

namespace isla {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n);   // kernels launched by this library, process-wide (isla_kernel_launches)

#define ISLA_CUDA_OK(expr)                                                   \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) return ::isla::cuda_fail(_e, #expr);          \
    } while (0)

// ---------------------------------------------------------------- thread-per-tree (vanilla) layout
// One NODE BLOCK per expanded state node: the statistics of its A outgoing edges, i.e. the reference's
// A ActionNodes (agents/abstract_mcts.py:145-156), each fused with its single StateNode child
// (deterministic envs: exactly one child per action node).  16 bytes per edge:
//   double total[A]   ActionNode.total
//   int32  na[A]      ActionNode.na
//   uint32 link[A]    bits 0..27 block index of the child state's own block (0 = none yet),
//                     bit 28 child is terminal, bit 29 the terminal step's reward is 1, bits 30..31 insertion rank
// A = 2 -> one 32-byte sector per UCB scan, A = 4 -> 64 bytes.  Cold per-edge data (the env state of the
// child and g_first, the return backed up at its creation) live in separate arrays indexed by
// slot = block * A + action.  Rewards are not stored: the envs of this kernel have a constant reward on
// non-terminal steps (CartPole 1, FrozenLake 0) and a 0/1 reward on the terminal step.
constexpr uint32_t kLinkBlockMask = 0x0fffffffu;
constexpr uint32_t kLinkTerminal = 1u << 28;
constexpr uint32_t kLinkRewardOne = 1u << 29;
constexpr int kLinkRankShift = 30;
constexpr int kRetTablePad = 64;   // zero doubles before and after the return table (isla_capi.cu: build_return_table)

// per-tree meta words
enum : int { META_ALLOC = 0, META_SIMS = 1, META_RNG = 2, META_FAULT = 3, META_TRACE_POS = 4, META_PENDING = 5, META_WORDS = 8 };
// device counters
enum : int { CNT_ENV_STEPS = 0, CNT_ROLLOUT_STEPS = 1, CNT_LEVELS = 2, CNT_SIMS = 3, CNT_NODES = 4, CNT_FAULTS = 5 };

struct TptParams {
    char* blocks;          // node blocks, A * 16 bytes each
    void* states;
    double* gfirst;
    int32_t* meta;
    char* pv;              // PV cache: per tree, chunks of 32 levels (layout: PvChunk in isla_tree_tpt.cu)
    long long pv_stride;   // bytes of one tree's chunks
    unsigned long long* counters;
    const uint8_t* trace_kinds;
    const double* trace_values;
    long long trace_len;
    long long n_trees;
    long long slots_per_tree;
    long long scripted_tree;
    unsigned long long seed;
    unsigned long long tree_id_base;
    double C, gamma;
    const double* ln_table;    // host-computed ln(n)
    const double* disc_table;  // host-computed gamma**t
    const double* cum_table;   // host-computed sum_{t<n} gamma**t (left-to-right)
    const double* ret_table;   // host-computed returns: [max_depth + 3 leaf values][ret_K levels above the leaf] (isla_capi.cu)
    int ret_K;
    int n_sim, max_depth, budget_mode, block_cap, ln_count, disc_count, ln_in_smem, tuning;
    const double* heur;        // grid_policy table (FrozenLake), nullptr = random rollouts
    const double* noise;       // slippery FrozenLake: noise[k] of the k-th env.step of a simulation, nullptr = deterministic
    int noise_n;
    int tpw;                   // trees owned by one warp (its first tpw lanes)
};

}  // namespace isla

struct isla_engine {
    isla_config cfg;
    isla_layout lay;
    char* ws;
    int kernel;
    int sm_count;
    bool roots_set;
    long long sims_run = 0;   // simulations requested since the last set_roots (bounds the nodes a tree can hold)
    double* heur_dev = nullptr;   // grid_policy prior knowledge [16][4] (engine-owned), nullptr = random rollouts
    double* noise_dev = nullptr;  // slippery FrozenLake: per-step transition noise of a fit (engine-owned)
    int noise_n = 0;
    char* reroot_tmp = nullptr;   // thread-per-tree engine: scratch arena of isla_engine_reroot (engine-owned, lazily sized)
    size_t reroot_bytes = 0;
    bool no_cluster16 = false;   // a 16-CTA (non-portable) cluster launch was refused on this device: stay with 8
    double* ret_dev = nullptr;    // thread-per-tree engine: return table (engine-owned), ret_rows x ret_K
    int ret_rows = 0, ret_K = 0;
};

namespace isla {
// thread-per-tree engine (isla_tree_tpt.cu)
int upload_tables(isla_engine* e);
void add_table_layout(const isla_config& cfg, isla_layout& lay, int64_t& off);
int tpt_fill_layout(const isla_config& cfg, isla_layout& lay);
int tpt_set_roots(isla_engine* e, const double* roots_dev, cudaStream_t st);
int tpt_run(isla_engine* e, int n_sim, long long scripted_tree, const uint8_t* kinds, const double* values,
            long long trace_len, cudaStream_t st);
int tpt_root_stats(isla_engine* e, int64_t* na_dev, double* total_dev, cudaStream_t st);
int tpt_reroot(isla_engine* e, const int32_t* action_dev, const uint8_t* alive_dev, int n_sim_next, cudaStream_t st);
int tpt_best(isla_engine* e, int scripted_pos, int32_t* best_dev, double* best_action_dev, cudaStream_t st);
// CTA-per-tree engine (isla_tree_cta.cu)
int cta_fill_layout(const isla_config& cfg, isla_layout& lay);
int cta_set_roots(isla_engine* e, const double* roots_dev, cudaStream_t st);
int cta_run(isla_engine* e, int n_sim, long long scripted_tree, const uint8_t* kinds, const double* values,
            long long trace_len, cudaStream_t st);
int cta_root_stats(isla_engine* e, int64_t* na_dev, double* total_dev, cudaStream_t st);
int cta_best(isla_engine* e, int scripted_pos, int32_t* best_dev, double* best_action_dev, cudaStream_t st);
int cta_root_children(isla_engine* e, long long tree, int cap, double* action, int64_t* na, double* total, int32_t* count,
                      cudaStream_t st);
void cta_pool_offsets(const isla_engine* e, long long out[24]);
int cta_depth_stats(isla_engine* e, long long tree, int cap, int32_t* dmax, double* dmean, int32_t* count, cudaStream_t st);
int tpt_depth_stats(isla_engine* e, long long tree, int cap, int32_t* dmax, double* dmean, int32_t* count, cudaStream_t st);
}  // namespace isla
