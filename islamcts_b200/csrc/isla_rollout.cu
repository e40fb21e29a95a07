// Batched env.step and random-policy rollouts: one thread per trajectory, env state in registers.
//
// isla_env_step  <-> gym env.step(a)                       (call sites: islaMcts/agents/mcts_hash.py:93,
//                                                           agents/abstract_mcts.py:85)
// isla_rollout   <-> AbstractStateNode.rollout             (agents/abstract_mcts.py:72-91) with
//                    continuous_default_policy             (utils/action_selection_functions.py:46-47)
// Roofline: FP32 issue rate (state never leaves registers; HBM traffic is one state read and one
// 12-byte result write per trajectory).
#include <cmath>
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
    const bool t = Env<ENV, T>::step(s, actions[i], r);
#pragma unroll
    for (int k = 0; k < S; ++k) next_states[i * S + k] = s[k];
    rewards[i] = r;
    term[i] = t ? 1 : 0;
}

template <int ENV, typename T>
__global__ void __launch_bounds__(256) rollout_kernel(const T* __restrict__ states, long long n, int budget, double gamma,
                                                      unsigned long long seed, unsigned sim, const double* __restrict__ disc_table,
                                                      double* __restrict__ returns, int32_t* __restrict__ steps,
                                                      unsigned long long* __restrict__ total_steps) {
    constexpr int S = EnvTraits<ENV>::S;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int t = 0;
    if (i < n) {
        T s[S];
        if (states) {
#pragma unroll
            for (int k = 0; k < S; ++k) s[k] = states[i * S + k];
        } else {
            env_reset<ENV, T>(seed, (uint32_t)i, s);
        }
        Stream rs;
        rs.init(seed, sim, (uint32_t)i, 0u);
        double R = 0.0;
        bool done = false;
        while (!done && t < budget) {
            const T a = sample_action<ENV, T>(rs);
            T r;
            done = Env<ENV, T>::step(s, a, r);
            R = __dadd_rn(R, __dmul_rn((double)r, __ldg(disc_table + t)));  // r * pow(gamma, t), host libm pow
            ++t;
        }
        returns[i] = R;
        steps[i] = t;
    }
    if (total_steps) {
        // block-level sum of executed steps -> one atomic per warp
        unsigned v = (unsigned)t;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(total_steps, (unsigned long long)v);
    }
}

template <typename T>
int env_step_dispatch(int env, const void* s, const void* a, long long n, void* ns, void* r, uint8_t* t, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    switch (env) {
    case ENV_CARTPOLE: env_step_kernel<ENV_CARTPOLE, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    case ENV_PENDULUM: env_step_kernel<ENV_PENDULUM, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    case ENV_MOUNTAINCAR: env_step_kernel<ENV_MOUNTAINCAR, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    case ENV_FROZENLAKE: env_step_kernel<ENV_FROZENLAKE, T><<<grid, 256, 0, st>>>((const T*)s, (const T*)a, n, (T*)ns, (T*)r, t); break;
    default: set_error("unknown env %d", env); return ISLA_ERR_INVALID;
    }
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
int rollout_dispatch(int env, const void* s, long long n, int budget, double gamma, unsigned long long seed, unsigned sim,
                     double* ret, int32_t* steps, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    // gamma**t table from the host libm (CPython's pow, agents/abstract_mcts.py:87), stream-ordered scratch
    std::vector<double> h((size_t)(budget > 0 ? budget : 1));
    for (size_t t = 0; t < h.size(); ++t) h[t] = std::pow(gamma, (double)t);
    double* disc = nullptr;
    ISLA_CUDA_OK(cudaMallocAsync((void**)&disc, h.size() * 8, st));
    ISLA_CUDA_OK(cudaMemcpyAsync(disc, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
    ISLA_CUDA_OK(cudaStreamSynchronize(st));  // h is pageable host memory
    switch (env) {
    case ENV_CARTPOLE: rollout_kernel<ENV_CARTPOLE, T><<<grid, 256, 0, st>>>((const T*)s, n, budget, gamma, seed, sim, disc, ret, steps, nullptr); break;
    case ENV_PENDULUM: rollout_kernel<ENV_PENDULUM, T><<<grid, 256, 0, st>>>((const T*)s, n, budget, gamma, seed, sim, disc, ret, steps, nullptr); break;
    case ENV_MOUNTAINCAR: rollout_kernel<ENV_MOUNTAINCAR, T><<<grid, 256, 0, st>>>((const T*)s, n, budget, gamma, seed, sim, disc, ret, steps, nullptr); break;
    case ENV_FROZENLAKE: rollout_kernel<ENV_FROZENLAKE, T><<<grid, 256, 0, st>>>((const T*)s, n, budget, gamma, seed, sim, disc, ret, steps, nullptr); break;
    default: cudaFreeAsync(disc, st); set_error("unknown env %d", env); return ISLA_ERR_INVALID;
    }
    ISLA_CUDA_OK(cudaGetLastError());
    ISLA_CUDA_OK(cudaFreeAsync(disc, st));
    return 0;
}

}  // namespace
}  // namespace isla

extern "C" int isla_env_step(int32_t env_id, int32_t precision, const void* states_dev, const void* actions_dev, int64_t n,
                             void* next_states_dev, void* rewards_dev, uint8_t* terminated_dev, void* stream) {
    if (n <= 0) return ISLA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return precision == ISLA_PRECISION_F64
               ? isla::env_step_dispatch<double>(env_id, states_dev, actions_dev, n, next_states_dev, rewards_dev, terminated_dev, st)
               : isla::env_step_dispatch<float>(env_id, states_dev, actions_dev, n, next_states_dev, rewards_dev, terminated_dev, st);
}

extern "C" int isla_rollout(int32_t env_id, int32_t precision, const void* states_dev, int64_t n, int32_t budget, double gamma,
                            uint64_t seed, uint32_t sim, double* returns_dev, int32_t* steps_dev, void* stream) {
    if (n <= 0) return ISLA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return precision == ISLA_PRECISION_F64
               ? isla::rollout_dispatch<double>(env_id, states_dev, n, budget, gamma, seed, sim, returns_dev, steps_dev, st)
               : isla::rollout_dispatch<float>(env_id, states_dev, n, budget, gamma, seed, sim, returns_dev, steps_dev, st);
}
