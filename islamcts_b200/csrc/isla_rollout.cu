// Batched env.step and random-policy rollouts: one thread per trajectory, env state in registers.
//
// isla_env_step  <-> gym env.step(a)                       (call sites: islaMcts/agents/mcts_hash.py:93,
//                                                           agents/abstract_mcts.py:85)
// isla_rollout   <-> AbstractStateNode.rollout             (agents/abstract_mcts.py:72-91) with
//                    continuous_default_policy             (utils/action_selection_functions.py:46-47)
// Roofline: FP32 issue rate (state never leaves registers; HBM traffic is one state read and one
// 12-byte result write per trajectory).
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "isla_envs.cuh"
#include "isla_internal.h"
#include "isla_philox.cuh"

namespace isla {
namespace {

template <int ENV, typename T>
__global__ void env_step_kernel(const T* __restrict__ states, const T* __restrict__ actions, long long n,
                                T* __restrict__ next_states, T* __restrict__ rewards, uint8_t* __restrict__ term) {
    constexpr int S = EnvTraits<ENV>::S;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T s[S];
#pragma unroll
    for (int k = 0; k < S; ++k) s[k] = states[i * S + k];
    T r;
    bool t;
    if constexpr (ENV == ENV_CURVE) t = Env<ENV, T>::step2(s, actions[2 * i], actions[2 * i + 1], r);
    else t = Env<ENV, T>::step(s, actions[i], r);
#pragma unroll
    for (int k = 0; k < S; ++k) next_states[i * S + k] = s[k];
    rewards[i] = r;
    term[i] = t ? 1 : 0;
}

// One thread per trajectory, one launch-sized grid: for envs whose random episodes (almost) never end before the
// budget (Pendulum, MountainCarContinuous) every lane runs the same number of steps and nothing is gained by refilling.
template <int ENV, typename T>
__global__ void __launch_bounds__(256) rollout_kernel_simple(const T* __restrict__ states, long long n, int budget,
                                                             unsigned long long seed, unsigned sim, ActionGrid grid,
                                                             const double* __restrict__ disc_table, double* __restrict__ returns,
                                                             int32_t* __restrict__ steps) {
    constexpr int S = EnvTraits<ENV>::S;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T s[S];
    if (states) {
#pragma unroll
        for (int k = 0; k < S; ++k) s[k] = states[i * S + k];
    } else {
        env_reset<ENV, T>(seed, (uint32_t)i, s);
    }
    double R = 0.0;
    int t = 0;
    if constexpr (EnvTraits<ENV>::kContinuous && EnvTraits<ENV>::AD == 1) {
        T last_r;
        R = rollout_fast1d<ENV, T>(seed, sim, (uint32_t)i, 0u, s, budget, disc_table, t, last_r);   // four steps per Philox block, branch-free
    } else {
        Stream rs;
        rs.init(seed, sim, (uint32_t)i, 0u);
        bool done = false;
        while (!done && t < budget) {
            T r;
            done = rollout_step<ENV, T>(rs, s, grid, r);
            R = __dadd_rn(R, __dmul_rn((double)r, __ldg(disc_table + t)));  // r * pow(gamma, t), host libm pow
            ++t;
        }
    }
    returns[i] = R;
    steps[i] = t;
}

// Persistent warps for random-length episodes (CartPole under the random policy: mean 22 steps, maximum over a warp ~56).
// A lane that finishes takes another trajectory instead of idling.  Starting a trajectory costs two Philox blocks (the
// reset draw and the first 128 action bits) -- more than two env steps -- so starts are PRODUCED with all 32 lanes busy,
// 32 at a time, into a per-warp queue in shared memory {start state, first action block, trajectory id}; taking one is two
// LDS.  The warp meets every kInner env steps: finished lanes store their result and take the next queued start (rank by
// ballot, no atomics), and the queue is topped up when it has room for 32 more (one global atomic per 32 trajectories).
// Between meetings there is no warp-level instruction at all; a lane that finishes mid-way idles < kInner steps.
// (Round 1 met after EVERY step and started trajectories with whatever lanes were waiting, >= 12: 104 warp instructions
// per env step at 22 of 32 lanes busy; see profiles/README.md for the ncu numbers of both.)
#ifndef ISLA_ROLL_INNER
#define ISLA_ROLL_INNER 8
#endif
#ifndef ISLA_ROLL_INNER_LAKE
#define ISLA_ROLL_INNER_LAKE 4
#endif
// env steps between two meetings of a warp: CartPole episodes last 22 steps on average, FrozenLake's 7.7
template <int ENV> constexpr int kInnerOf = ENV == ENV_FROZENLAKE ? ISLA_ROLL_INNER_LAKE : ISLA_ROLL_INNER;
constexpr int kQueue = 64;   // queued starts per warp (two productions)

template <int ENV, typename T>
__global__ void __launch_bounds__(256) rollout_kernel(const T* __restrict__ states, long long n, int budget, double gamma,
                                                      unsigned long long seed, unsigned sim, const double* __restrict__ disc_table,
                                                      const double* __restrict__ cum_table, double* __restrict__ returns,
                                                      int32_t* __restrict__ steps, unsigned long long* __restrict__ next) {
    constexpr int S = EnvTraits<ENV>::S;
    constexpr int kInner = kInnerOf<ENV>;
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ T q_state[8][kQueue][S];
    __shared__ uint32_t q_bits[8][kQueue][4];
    __shared__ uint32_t q_idx[8][kQueue];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned head = 0, tail = 0;          // warp-uniform: queue positions (slot = position % kQueue)
    bool exhausted = false;               // warp-uniform: no trajectory left to claim
    bool active = false, pending = false; // pending: finished, result not stored yet
    uint32_t idx = 0;
    T s[S];
#pragma unroll
    for (int k = 0; k < S; ++k) s[k] = T(0);
    Stream rs;
    rs.init(seed, sim, 0u, 0u);
    int t = 0;
    double R = 0.0;
    T last_r = T(0);
    for (;;) {
        if (pending) {
            if constexpr (Rewards<ENV>::kClosedForm) R = Rewards<ENV>::rollout_return(t, (double)last_r, disc_table, cum_table);
            returns[idx] = R;
            steps[idx] = t;
            pending = false;
        }
        if (!exhausted && tail - head <= (unsigned)(kQueue - 32)) {
            __syncwarp();                                                    // every lane is done reading the slots reused below
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(next, 32ull);
            base = __shfl_sync(kFull, base, 0);
            const long long mine = (long long)base + lane;
            if (mine < n) {
                const int slot = (int)((tail + (unsigned)lane) % kQueue);
                T st0[S];
                if (states) {
#pragma unroll
                    for (int k = 0; k < S; ++k) st0[k] = states[mine * S + k];
                } else {
                    env_reset<ENV, T>(seed, (uint32_t)mine, st0);
                }
                uint32_t b[4];
                philox4x32_10(0u, sim, (uint32_t)mine, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), b);   // block 0 of its action stream
#pragma unroll
                for (int k = 0; k < S; ++k) q_state[w][slot][k] = st0[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) q_bits[w][slot][k] = b[k];
                q_idx[w][slot] = (uint32_t)mine;
            }
            const long long left = n - (long long)base;
            tail += left >= 32 ? 32u : left > 0 ? (unsigned)left : 0u;
            exhausted = left <= 32;
            __syncwarp();
        }
        const unsigned want = __ballot_sync(kFull, !active);
        if (want != 0u && head != tail) {
            if (!active) {
                const unsigned k = head + __popc(want & lt);
                if ((int)(tail - k) > 0) {
                    const int slot = (int)(k % kQueue);
#pragma unroll
                    for (int q = 0; q < S; ++q) s[q] = q_state[w][slot][q];
                    idx = q_idx[w][slot];
                    rs.init(seed, sim, idx, 0u);
                    rs.blk = 1u; rs.idx = 0;                                 // its first block was drawn by the producer
#pragma unroll
                    for (int q = 0; q < 4; ++q) rs.buf[q] = q_bits[w][slot][q];
                    t = 0; R = 0.0; last_r = T(0);
                    active = budget > 0;
                    pending = !active;
                }
            }
            const unsigned avail = tail - head, nw = (unsigned)__popc(want);
            head += nw < avail ? nw : avail;
        }
        if (__ballot_sync(kFull, active || pending) == 0u) {
            if (exhausted && head == tail) break;
            continue;
        }
#if ISLA_ROLL_NOUNROLL
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int u = 0; u < kInner; ++u) {
            if (active) {
                const T a = sample_action<ENV, T>(rs);
                const bool done = Env<ENV, T>::step(s, a, last_r);
                if constexpr (!Rewards<ENV>::kClosedForm) R = __dadd_rn(R, __dmul_rn((double)last_r, __ldg(disc_table + t)));  // r * pow(gamma, t)
                ++t;
                if (done || t >= budget) { active = false; pending = true; }
            }
        }
    }
}

// (FrozenLake, 7.7-step episodes, runs the same kernel with a meeting every 4 steps.  Its results are stored by the finished
//  lanes of a meeting together, and the queue hands out consecutive ids, so the 12-byte results of a warp land in the same
//  sectors within a few hundred cycles and merge in L2: DRAM reads 8.7 MB per 2^24 trajectories.  Round 1 claimed single
//  trajectories per idle lane at arbitrary times -- 8 GB of read-modify-write traffic -- and an intermediate design parked
//  units of 8 consecutive results in shared memory for full-sector stores: same traffic as now, 168 G env-steps/s against 289 G.)

// real-step half of the episode loop: env.step(best action) for every live episode, float64, in place
template <int ENV>
__global__ void episode_step_kernel(double* __restrict__ states, const double* __restrict__ actions, long long n, ActionGrid grid,
                                    int max_steps, double timeout_reward, uint8_t* __restrict__ alive, double* __restrict__ total_reward,
                                    int32_t* __restrict__ n_steps) {
    constexpr int S = EnvTraits<ENV>::S;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !alive[i]) return;
    double s[S];
#pragma unroll
    for (int k = 0; k < S; ++k) s[k] = states[i * S + k];
    const int AD = (ENV == ENV_CURVE && grid.g0 == 0) ? 2 : 1;
    double r;
    const bool term = env_step_action<ENV, double>(s, actions[i * AD], AD == 2 ? actions[i * AD + 1] : 0.0, grid, r);
#pragma unroll
    for (int k = 0; k < S; ++k) states[i * S + k] = s[k];
    const int steps = n_steps[i] + 1;
    n_steps[i] = steps;
    // the reference's sweep loop replaces the reward of the step that reaches the limit (sweep.py:206-209: reward = -1000)
    if (steps >= max_steps && !isnan(timeout_reward)) r = timeout_reward;
    total_reward[i] = __dadd_rn(total_reward[i], r);
    if (term || steps >= max_steps) alive[i] = 0;
}

template <typename T>
int env_step_dispatch(int env, const void* s, const void* a, long long n, void* ns, void* r, uint8_t* t, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    switch (env) {
    case ENV_CARTPOLE: env_step_kernel<ENV_CARTPOLE, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    case ENV_PENDULUM: env_step_kernel<ENV_PENDULUM, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    case ENV_MOUNTAINCAR: env_step_kernel<ENV_MOUNTAINCAR, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    case ENV_FROZENLAKE: env_step_kernel<ENV_FROZENLAKE, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    case ENV_CURVE: env_step_kernel<ENV_CURVE, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    default: set_error("unknown env %d", env); return ISLA_ERR_INVALID;
    }
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

// gamma**t and its left-to-right prefix sums from the host libm (CPython's pow, agents/abstract_mcts.py:87).  The
// device copies are cached per (device, gamma, length): later calls enqueue no copy and do not synchronise.
struct DiscTables { int dev; double gamma; size_t m; double* ptr; };
static std::mutex g_tab_mu;
static std::vector<DiscTables> g_tabs;

static int get_tables(double gamma, int budget, const double** disc, const double** cum) {
    int dev = 0;
    ISLA_CUDA_OK(cudaGetDevice(&dev));
    const size_t m = (size_t)(budget > 0 ? budget : 1) + 1;
    std::lock_guard<std::mutex> lk(g_tab_mu);
    for (const DiscTables& t : g_tabs)
        if (t.dev == dev && t.m >= m && std::memcmp(&t.gamma, &gamma, sizeof gamma) == 0) { *disc = t.ptr; *cum = t.ptr + t.m; return 0; }
    std::vector<double> h(2 * m);
    double acc = 0.0;
    for (size_t t = 0; t < m; ++t) { h[t] = std::pow(gamma, (double)t); h[m + t] = acc; acc = acc + 1.0 * h[t]; }
    double* p = nullptr;
    ISLA_CUDA_OK(cudaMalloc((void**)&p, h.size() * 8));
    ISLA_CUDA_OK(cudaMemcpy(p, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    ISLA_CUDA_OK(cudaDeviceSynchronize());
    g_tabs.push_back(DiscTables{dev, gamma, m, p});
    *disc = p; *cum = p + m;
    return 0;
}

// Claim counters of the persistent rollout kernels: one ring of kClaimSlots counters per device, allocated once.  A launch
// takes the next slot and zeroes it on its own stream (so up to kClaimSlots rollout launches may be in flight at once).
// A cudaMallocAsync / cudaFreeAsync pair per call cost 0.3-4 ms whenever the pool had been trimmed at a synchronise --
// more than the 1.1 ms kernel it served.
constexpr int kClaimSlots = 256;
struct ClaimRing { int dev; unsigned long long* ptr; unsigned nextslot; };
static std::vector<ClaimRing> g_claims;

static int claim_counter(unsigned long long** out, cudaStream_t st) {
    int dev = 0;
    ISLA_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    ClaimRing* ring = nullptr;
    for (ClaimRing& r : g_claims)
        if (r.dev == dev) ring = &r;
    if (!ring) {
        unsigned long long* p = nullptr;
        ISLA_CUDA_OK(cudaMalloc((void**)&p, kClaimSlots * sizeof(unsigned long long)));
        g_claims.push_back(ClaimRing{dev, p, 0u});
        ring = &g_claims.back();
    }
    *out = ring->ptr + (ring->nextslot++ % kClaimSlots);
    ISLA_CUDA_OK(cudaMemsetAsync(*out, 0, sizeof(unsigned long long), st));
    return 0;
}

template <int ENV, typename T>
int rollout_launch(const void* s, long long n, int budget, double gamma, unsigned long long seed, unsigned sim, ActionGrid ag,
                   const double* disc, const double* cum, double* ret, int32_t* steps, int sms, cudaStream_t st) {
    if constexpr (Rewards<ENV>::kClosedForm) {
        // random episode lengths: persistent grid (every resident CTA slot of the GPU), lane refill
        unsigned long long* next = nullptr;
        if (int rc = claim_counter(&next, st)) return rc;
        {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rollout_kernel<ENV, T>, 256, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 4; }
            const long long want = (n + 255) / 256;
            const unsigned grid = (unsigned)(want < (long long)sms * occ ? want : (long long)sms * occ);
            rollout_kernel<ENV, T><<<grid, 256, 0, st>>>((const T*)s, n, budget, gamma, seed, sim, disc, cum, ret, steps, next);
        }
        ISLA_CUDA_OK(cudaGetLastError());
    } else {
        rollout_kernel_simple<ENV, T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T*)s, n, budget, seed, sim, ag, disc, ret, steps);
        ISLA_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

template <typename T>
int rollout_dispatch(int env, const void* s, long long n, int budget, double gamma, unsigned long long seed, unsigned sim,
                     ActionGrid ag, double* ret, int32_t* steps, cudaStream_t st) {
    int dev = 0, sms = 148;
    ISLA_CUDA_OK(cudaGetDevice(&dev));
    ISLA_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const double *disc = nullptr, *cum = nullptr;
    const int rc = get_tables(gamma, budget, &disc, &cum);
    if (rc) return rc;
    switch (env) {
    case ENV_CARTPOLE: return rollout_launch<ENV_CARTPOLE, T>(s, n, budget, gamma, seed, sim, ag, disc, cum, ret, steps, sms, st);
    case ENV_PENDULUM: return rollout_launch<ENV_PENDULUM, T>(s, n, budget, gamma, seed, sim, ag, disc, cum, ret, steps, sms, st);
    case ENV_MOUNTAINCAR: return rollout_launch<ENV_MOUNTAINCAR, T>(s, n, budget, gamma, seed, sim, ag, disc, cum, ret, steps, sms, st);
    case ENV_FROZENLAKE: return rollout_launch<ENV_FROZENLAKE, T>(s, n, budget, gamma, seed, sim, ag, disc, cum, ret, steps, sms, st);
    case ENV_CURVE: return rollout_launch<ENV_CURVE, T>(s, n, budget, gamma, seed, sim, ag, disc, cum, ret, steps, sms, st);
    }
    set_error("unknown env %d", env);
    return ISLA_ERR_INVALID;
}

}  // namespace
}  // namespace isla

extern "C" int isla_env_step(int32_t env_id, int32_t precision, const void* states_dev, const void* actions_dev, int64_t n,
                             void* next_states_dev, void* rewards_dev, uint8_t* terminated_dev, void* stream) {
    if (n <= 0) return ISLA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = precision == ISLA_PRECISION_F64
               ? isla::env_step_dispatch<double>(env_id, states_dev, actions_dev, n, next_states_dev, rewards_dev, terminated_dev, st)
               : isla::env_step_dispatch<float>(env_id, states_dev, actions_dev, n, next_states_dev, rewards_dev, terminated_dev, st);
    if (rc == 0) isla::count_launch(1);
    return rc;
}

extern "C" int isla_episode_step(int32_t env_id, int32_t action_grid0, int32_t action_grid1, double* states_dev,
                                 const double* actions_dev, int64_t n, int32_t max_steps, double timeout_reward, uint8_t* alive_dev,
                                 double* total_reward_dev, int32_t* n_steps_dev, void* stream) {
    using namespace isla;
    if (n <= 0) return ISLA_OK;
    if (!states_dev || !actions_dev || !alive_dev || !total_reward_dev || !n_steps_dev) { set_error("null argument"); return ISLA_ERR_INVALID; }
    ActionGrid ag{0, 0};
    if (env_id == ISLA_ENV_CURVE && (action_grid0 != 0 || action_grid1 != 0)) {
        if (action_grid0 < 1 || action_grid1 < 1 || (int64_t)action_grid0 * action_grid1 > 128) {
            set_error("DiscreteCurveEnv grid %d x %d: need 1 <= g0*g1 <= 128", action_grid0, action_grid1); return ISLA_ERR_INVALID;
        }
        ag = ActionGrid{action_grid0, action_grid1};
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + 127) / 128);
    switch (env_id) {
    case ENV_CARTPOLE: episode_step_kernel<ENV_CARTPOLE><<<grid, 128, 0, st>>>(states_dev, actions_dev, n, ag, max_steps, timeout_reward, alive_dev, total_reward_dev, n_steps_dev); break;
    case ENV_PENDULUM: episode_step_kernel<ENV_PENDULUM><<<grid, 128, 0, st>>>(states_dev, actions_dev, n, ag, max_steps, timeout_reward, alive_dev, total_reward_dev, n_steps_dev); break;
    case ENV_MOUNTAINCAR: episode_step_kernel<ENV_MOUNTAINCAR><<<grid, 128, 0, st>>>(states_dev, actions_dev, n, ag, max_steps, timeout_reward, alive_dev, total_reward_dev, n_steps_dev); break;
    case ENV_FROZENLAKE: episode_step_kernel<ENV_FROZENLAKE><<<grid, 128, 0, st>>>(states_dev, actions_dev, n, ag, max_steps, timeout_reward, alive_dev, total_reward_dev, n_steps_dev); break;
    case ENV_CURVE: episode_step_kernel<ENV_CURVE><<<grid, 128, 0, st>>>(states_dev, actions_dev, n, ag, max_steps, timeout_reward, alive_dev, total_reward_dev, n_steps_dev); break;
    default: set_error("unknown env %d", env_id); return ISLA_ERR_INVALID;
    }
    ISLA_CUDA_OK(cudaGetLastError());
    count_launch(1);
    return ISLA_OK;
}

extern "C" int isla_rollout(int32_t env_id, int32_t precision, const void* states_dev, int64_t n, int32_t budget, double gamma,
                            uint64_t seed, uint32_t sim, int32_t action_grid0, int32_t action_grid1, double* returns_dev,
                            int32_t* steps_dev, void* stream) {
    if (n <= 0) return ISLA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    isla::ActionGrid ag{0, 0};
    if (env_id == ISLA_ENV_CURVE && (action_grid0 != 0 || action_grid1 != 0)) {
        if (action_grid0 < 1 || action_grid1 < 1 || (int64_t)action_grid0 * action_grid1 > 128) {
            isla::set_error("DiscreteCurveEnv grid %d x %d: need 1 <= g0*g1 <= 128", action_grid0, action_grid1); return ISLA_ERR_INVALID;
        }
        ag = isla::ActionGrid{action_grid0, action_grid1};
    }
    const int rc = precision == ISLA_PRECISION_F64
               ? isla::rollout_dispatch<double>(env_id, states_dev, n, budget, gamma, seed, sim, ag, returns_dev, steps_dev, st)
               : isla::rollout_dispatch<float>(env_id, states_dev, n, budget, gamma, seed, sim, ag, returns_dev, steps_dev, st);
    if (rc == 0) isla::count_launch(1);
    return rc;
}
