// On-device environment dynamics: CartPole-v1, Pendulum-v1, MountainCarContinuous-v0, FrozenLake-v1 and the
// reference's own CurveEnv / DiscreteCurveEnv.
//
// Replaces env.step() on the hot path (reference call sites: islaMcts/agents/mcts.py:79,
// mcts_hash.py:93, abstract_mcts.py:85).  The gym arithmetic is third-party Gymnasium, absent from
// /root/reference and unpinned by it; the definitions followed are the ones SURVEY.md section 8c restates.
// CurveEnv is reference code (islaMcts/environments/curve_env.py) and is pinned by step vectors generated from it
// (tests/golden/env/curve_env_steps.npz).
//
// Two arithmetic modes, chosen at compile time by the state type T:
//   float  : the fast path (state in registers, FMA contraction allowed, fp32 trig).
//   double : "faithful" mode.  Every product/sum is rounded separately (no FMA contraction) in the
//            same order as the float64 reference arithmetic, so that whole searches can be compared
//            count-for-count with the CPU oracle.
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#include "isla_philox.cuh"

namespace isla {

enum : int { ENV_CARTPOLE = 0, ENV_PENDULUM = 1, ENV_MOUNTAINCAR = 2, ENV_FROZENLAKE = 3, ENV_CURVE = 4, ENV_COUNT = 5 };

// DiscreteCurveEnv([g0, g1]) (curve_env.py:184-199): g0 accelerations x g1 steering angles; g0 == 0 = the continuous CurveEnv
struct ActionGrid { int g0, g1; };

template <typename T> struct Ar;
template <> struct Ar<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ void sincos(double x, double& s, double& c) { s = ::sin(x); c = ::cos(x); }
    static __device__ __forceinline__ double cosine(double x) { return ::cos(x); }
    static __device__ __forceinline__ double sine(double x) { return ::sin(x); }
    static __device__ __forceinline__ double fmodulo(double a, double b) { return ::fmod(a, b); }
    static __device__ __forceinline__ double fmin_(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double fmax_(double a, double b) { return ::fmax(a, b); }
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
    static __device__ __forceinline__ void sincos(float x, float& s, float& c) { ::sincosf(x, &s, &c); }
    static __device__ __forceinline__ float cosine(float x) { return ::cosf(x); }
    static __device__ __forceinline__ float sine(float x) { return ::sinf(x); }
    static __device__ __forceinline__ float fmodulo(float a, float b) { return ::fmodf(a, b); }
    static __device__ __forceinline__ float fmin_(float a, float b) { return ::fminf(a, b); }
    static __device__ __forceinline__ float fmax_(float a, float b) { return ::fmaxf(a, b); }
};

template <int ENV> struct EnvTraits;
// S: env state reals; A: discrete actions (0: Box, or decided at run time by the ActionGrid); AD: action dimensions
template <> struct EnvTraits<ENV_CARTPOLE>    { static constexpr int S = 4, A = 2, AD = 1; static constexpr bool kContinuous = false; };
template <> struct EnvTraits<ENV_PENDULUM>    { static constexpr int S = 2, A = 0, AD = 1; static constexpr bool kContinuous = true;  };
template <> struct EnvTraits<ENV_MOUNTAINCAR> { static constexpr int S = 2, A = 0, AD = 1; static constexpr bool kContinuous = true;  };
template <> struct EnvTraits<ENV_FROZENLAKE>  { static constexpr int S = 1, A = 4, AD = 1; static constexpr bool kContinuous = false; };
// state = (x, y, velocity, angle [deg], n_step): the step counter is hidden state of the env object (it scales rewards)
template <> struct EnvTraits<ENV_CURVE>       { static constexpr int S = 5, A = 0, AD = 2; static constexpr bool kContinuous = true;  };

__host__ __device__ inline int env_state_dim(int env) { return env == ENV_CARTPOLE ? 4 : env == ENV_FROZENLAKE ? 1 : env == ENV_CURVE ? 5 : 2; }
// fixed discrete action count of the gym envs (0: continuous, or CurveEnv whose count comes from the ActionGrid)
__host__ __device__ inline int env_n_actions(int env) { return env == ENV_CARTPOLE ? 2 : env == ENV_FROZENLAKE ? 4 : 0; }
__host__ __device__ inline int env_action_dim(int env) { return env == ENV_CURVE ? 2 : 1; }
__host__ __device__ inline void env_action_bounds(int env, float& lo, float& hi) {
    if (env == ENV_PENDULUM) { lo = -2.0f; hi = 2.0f; } else { lo = -1.0f; hi = 1.0f; }
}

// step(): advances s in place, returns `terminated`, writes the reward.
template <int ENV, typename T> struct Env;

template <typename T> struct Env<ENV_CARTPOLE, T> {
    using A_ = Ar<T>;
    static __device__ __forceinline__ bool step(T (&s)[4], T action, T& reward) {
        constexpr T gravity = T(9.8), masspole = T(0.1), total_mass = T(0.1 + 1.0), length = T(0.5),
                    polemass_length = T(0.1 * 0.5), force_mag = T(10.0), tau = T(0.02),
                    th_lim = T(12 * 2 * 3.141592653589793 / 360), x_lim = T(2.4), four_thirds = T(4.0 / 3.0);
        const T x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
        const T force = (action >= T(0.5)) ? force_mag : -force_mag;
        T sintheta, costheta;
        A_::sincos(theta, sintheta, costheta);
        const T temp = A_::div(A_::add(force, A_::mul(A_::mul(polemass_length, A_::mul(theta_dot, theta_dot)), sintheta)), total_mass);
        const T thetaacc = A_::div(A_::sub(A_::mul(gravity, sintheta), A_::mul(costheta, temp)),
                                   A_::mul(length, A_::sub(four_thirds, A_::div(A_::mul(masspole, A_::mul(costheta, costheta)), total_mass))));
        const T xacc = A_::sub(temp, A_::div(A_::mul(A_::mul(polemass_length, thetaacc), costheta), total_mass));
        s[0] = A_::add(x, A_::mul(tau, x_dot));
        s[1] = A_::add(x_dot, A_::mul(tau, xacc));
        s[2] = A_::add(theta, A_::mul(tau, theta_dot));
        s[3] = A_::add(theta_dot, A_::mul(tau, thetaacc));
        reward = T(1.0);
        return (s[0] < -x_lim) || (s[0] > x_lim) || (s[2] < -th_lim) || (s[2] > th_lim);
    }
};

// Fast float32 CartPole (the rollout inner loop): same equations, but constants folded, divisions by
// constants turned into multiplications, the one true division done with the fast reciprocal, and
// sin/cos from short polynomials (|theta| <= 0.5 rad: truncation < 1e-8, i.e. below float32 rounding; a
// non-terminated CartPole state has |theta| <= 0.21).  Within 1e-5 relative of the float64 definition
// (tests/test_gpu_envs.py); the double instantiation above stays operation-exact.
template <> struct Env<ENV_CARTPOLE, float> {
    static __device__ __forceinline__ bool step(float (&s)[4], float action, float& reward) {
        constexpr float gravity = 9.8f, masspole = 0.1f, inv_total_mass = float(1.0 / (0.1 + 1.0)), length = 0.5f,
                        polemass_length = 0.05f, force_mag = 10.0f, tau = 0.02f,
                        th_lim = float(12 * 2 * 3.141592653589793 / 360), x_lim = 2.4f, four_thirds = float(4.0 / 3.0);
        const float x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
        const float force = (action >= 0.5f) ? force_mag : -force_mag;
        float sintheta, costheta;
        if (fabsf(theta) <= 0.5f) {
            const float t2 = theta * theta;
            sintheta = theta * fmaf(t2, fmaf(t2, fmaf(t2, fmaf(t2, 2.7557319e-6f, -1.9841270e-4f), 8.3333333e-3f), -1.6666667e-1f), 1.0f);
            costheta = fmaf(t2, fmaf(t2, fmaf(t2, fmaf(t2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.0f);
        } else {
            ::sincosf(theta, &sintheta, &costheta);
        }
        const float temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) * inv_total_mass;
        const float denom = length * (four_thirds - masspole * (costheta * costheta) * inv_total_mass);
        const float thetaacc = __fdividef(gravity * sintheta - costheta * temp, denom);
        const float xacc = temp - polemass_length * thetaacc * costheta * inv_total_mass;
        s[0] = x + tau * x_dot;
        s[1] = x_dot + tau * xacc;
        s[2] = theta + tau * theta_dot;
        s[3] = theta_dot + tau * thetaacc;
        reward = 1.0f;
        return (fabsf(s[0]) > x_lim) || (fabsf(s[2]) > th_lim);
    }
};

template <typename T> struct Env<ENV_PENDULUM, T> {
    using A_ = Ar<T>;
    static __device__ __forceinline__ bool step(T (&s)[2], T action, T& reward) {
        constexpr T max_speed = T(8.0), max_torque = T(2.0), dt = T(0.05), pi = T(3.141592653589793),
                    two_pi = T(2 * 3.141592653589793), c_sin = T(3 * 10.0 / (2 * 1.0)), c_u = T(3.0 / (1.0 * (1.0 * 1.0)));
        const T th = s[0], thdot = s[1];
        const T u = A_::fmin_(A_::fmax_(action, -max_torque), max_torque);
        T ang = A_::fmodulo(A_::add(th, pi), two_pi);
        if (ang < T(0)) ang = A_::add(ang, two_pi);
        ang = A_::sub(ang, pi);
        const T costs = A_::add(A_::add(A_::mul(ang, ang), A_::mul(T(0.1), A_::mul(thdot, thdot))), A_::mul(T(0.001), A_::mul(u, u)));
        T newthdot = A_::add(thdot, A_::mul(A_::add(A_::mul(c_sin, A_::sine(th)), A_::mul(c_u, u)), dt));
        newthdot = A_::fmin_(A_::fmax_(newthdot, -max_speed), max_speed);
        s[0] = A_::add(th, A_::mul(newthdot, dt));
        s[1] = newthdot;
        reward = -costs;
        return false;
    }
};

// Fast float32 Pendulum / MountainCar (rollout inner loops).  The trigonometric argument is first reduced to
// [-pi, pi] with a two-constant Cody-Waite step, where the hardware approximation (MUFU) has an absolute error
// below 5e-7; everything else follows the float64 definition.  Within 1e-5 relative of it (tests/test_gpu_envs.py).
__device__ __forceinline__ float reduce_pi(float x) {
    const float k = rintf(x * 0.15915494309189535f);           // x / (2 pi)
    return fmaf(k, -1.7484556e-7f, fmaf(k, -6.2831855f, x));     // x - k * 2pi, 2pi = 6.2831855 + (-1.7484556e-7)
}
template <> struct Env<ENV_PENDULUM, float> {
    static __device__ __forceinline__ bool step(float (&s)[2], float action, float& reward) {
        constexpr float max_speed = 8.0f, max_torque = 2.0f, dt = 0.05f;
        const float th = s[0], thdot = s[1];
        const float u = fminf(fmaxf(action, -max_torque), max_torque);
        // angle_normalize(th) = ((th + pi) % (2 pi)) - pi: the same reduction to [-pi, pi] the sine needs (the two differ only
        // in which end of the interval th = (2k+1) pi lands on, and the cost uses ang * ang)
        const float ang = reduce_pi(th);
        const float costs = ang * ang + 0.1f * (thdot * thdot) + 0.001f * (u * u);
        float newthdot = thdot + (15.0f * __sinf(ang) + 3.0f * u) * dt;
        newthdot = fminf(fmaxf(newthdot, -max_speed), max_speed);
        s[0] = th + newthdot * dt;
        s[1] = newthdot;
        reward = -costs;
        return false;
    }
};
template <> struct Env<ENV_MOUNTAINCAR, float> {
    static __device__ __forceinline__ bool step(float (&s)[2], float action, float& reward) {
        constexpr float min_position = -1.2f, max_position = 0.6f, max_speed = 0.07f, goal_position = 0.45f, power = 0.0015f;
        float position = s[0], velocity = s[1];
        const float force = fminf(fmaxf(action, -1.0f), 1.0f);
        // 3 * position lies in [-3.6, 1.8]: the approximation's own range reduction (one multiply by 1/2pi, |x/2pi| < 0.58)
        // adds a phase error below 4e-7 there, so no explicit reduction to [-pi, pi] (4 dependent instructions of the chain)
        // The loop-carried chain of a rollout runs through the position: one multiply into the cosine unit, one fused
        // multiply-add out of it (force * power + velocity is ready before the cosine is).
        float c;
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(position * 3.0f));
        velocity = fmaf(c, -0.0025f, force * power + velocity);
        velocity = fminf(fmaxf(velocity, -max_speed), max_speed);
        position += velocity;
        position = fminf(fmaxf(position, min_position), max_position);
        if (position == min_position && velocity < 0.f) velocity = 0.f;
        const bool term = (position >= goal_position) && (velocity >= 0.f);
        reward = (term ? 100.0f : 0.0f) - (action * action) * 0.1f;
        s[0] = position;
        s[1] = velocity;
        return term;
    }
};

template <typename T> struct Env<ENV_MOUNTAINCAR, T> {
    using A_ = Ar<T>;
    static __device__ __forceinline__ bool step(T (&s)[2], T action, T& reward) {
        constexpr T min_position = T(-1.2), max_position = T(0.6), max_speed = T(0.07), goal_position = T(0.45),
                    goal_velocity = T(0.0), power = T(0.0015);
        T position = s[0], velocity = s[1];
        const T force = A_::fmin_(A_::fmax_(action, T(-1.0)), T(1.0));
        velocity = A_::add(velocity, A_::sub(A_::mul(force, power), A_::mul(T(0.0025), A_::cosine(A_::mul(T(3), position)))));
        velocity = A_::fmin_(A_::fmax_(velocity, -max_speed), max_speed);
        position = A_::add(position, velocity);
        position = A_::fmin_(A_::fmax_(position, min_position), max_position);
        if (position == min_position && velocity < T(0)) velocity = T(0);
        const bool term = (position >= goal_position) && (velocity >= goal_velocity);
        reward = A_::sub(term ? T(100.0) : T(0.0), A_::mul(A_::mul(action, action), T(0.1)));
        // gym keeps the state as float32
        s[0] = T(float(position));
        s[1] = T(float(velocity));
        return term;
    }
};

template <typename T> struct Env<ENV_FROZENLAKE, T> {
    // 4x4 map SFFF/FHFH/FFFH/HFFG, is_slippery=False.  holes {5,7,11,12}, goal 15.
    static __device__ __forceinline__ bool step(T (&s)[1], T action, T& reward) {
        constexpr unsigned kHole = (1u << 5) | (1u << 7) | (1u << 11) | (1u << 12), kGoal = 1u << 15;
        int cell = (int)s[0];
        if (((kHole | kGoal) >> cell) & 1u) { reward = T(0); return true; }  // stepping from a terminal cell
        int row = cell >> 2, col = cell & 3;
        const int a = (int)action;
        if (a == 0) col = max(col - 1, 0);
        else if (a == 1) row = min(row + 1, 3);
        else if (a == 2) col = min(col + 1, 3);
        else row = max(row - 1, 0);
        cell = row * 4 + col;
        s[0] = T(cell);
        reward = ((kGoal >> cell) & 1u) ? T(1) : T(0);
        return ((kHole | kGoal) >> cell) & 1u;
    }
};

// ---------------------------------------------------------------- CurveEnv (islaMcts/environments/curve_env.py)
// Car.update :34-56 and CurveEnv.step :121-177.  The car must stay between two parabolas; the reference tests
// 100 points of np.linspace(last_pos, new_pos) against them (:136-143).
//   upper(x) - y is CONCAVE along the segment, so its minimum over the 100 points is at an end point;
//   y - lower(x) is a CONVEX quadratic in the point index, so its minimum over the integers 0..99 is next to the
//   real-valued vertex.
// Hence at most six points decide containment.  The double instantiation evaluates exactly the reference's
// expressions at those points and falls back to the full 100-point loop whenever a decisive margin is within 1e-9 of
// zero, so its verdict is always the one of the 100-point test; the float instantiation (rollout fast path) uses the
// six points alone.
template <typename T> struct Env<ENV_CURVE, T> {
    using A_ = Ar<T>;
    static __device__ __forceinline__ T lower(T x) { return A_::add(A_::mul(T(-(1.0 / 50)), A_::mul(T(3), A_::mul(x, x))), T(27.3)); }
    static __device__ __forceinline__ T upper(T x) { return A_::add(A_::mul(T(-(1.0 / 50)), A_::mul(x, x)), T(27.5)); }
    // np.linspace(a, b, 100)[i] with step = (b - a) / 99
    static __device__ __forceinline__ T point(T a, T b, T step, int i) { return i == 99 ? b : A_::add(A_::mul(T(i), step), a); }
    static __device__ __forceinline__ T gradient(bool second, int idx) {
        // gradient_1 / gradient_2 of CurveEnv.__init__ (:62-74): 1.8*idx up to 26 (mirrored to 1.8*(55-idx) for x >= 0), 50 above
        if (idx >= 27) return T(50);
        if constexpr (sizeof(T) == 8) return __dmul_rn(1.80, (double)(second ? 55 - idx : idx));
        else return 1.8f * (float)(second ? 55 - idx : idx);
    }
    static __device__ __forceinline__ bool step2(T (&s)[5], T acceleration, T steer, T& reward) {
        constexpr T dt = T(0.5);
        T x = s[0], y = s[1], velocity = s[2], angle = s[3];
        const T n_step = A_::add(s[4], T(1));
        const T last_x = x, last_y = y;
        velocity = A_::add(velocity, A_::mul(acceleration, dt));
        if (velocity > T(20)) velocity = T(20);
        if (velocity < T(0)) velocity = T(0);
        angle = A_::fmodulo(A_::add(angle, steer), T(360));
        if (angle < T(0)) angle = A_::add(angle, T(360));                     // python float %
        const T rad = A_::mul(angle, T(3.141592653589793 / 180.0));          // math.radians
        T sn, cs;
        A_::sincos(rad, sn, cs);
        x = A_::add(x, A_::mul(A_::mul(cs, velocity), dt));
        y = A_::add(y, A_::mul(A_::mul(sn, velocity), dt));
        const T step_x = A_::div(A_::sub(x, last_x), T(99)), step_y = A_::div(A_::sub(y, last_y), T(99));
        // candidate points: the two ends and the neighbourhood of the vertex of y - lower(x)
        //   g(i) = 0.06*(lx + i*sx)^2 + ly + i*sy - 27.3  ->  i* = -(0.12*lx*sx + sy) / (0.12*sx^2)
        int iv = 0;
        {
            const T den = T(0.12) * step_x * step_x;
            if (den > T(0)) {
                const T v = -(T(0.12) * last_x * step_x + step_y) / den;
                iv = v <= T(0) ? 0 : v >= T(99) ? 99 : (int)v;
            }
        }
        T margin = T(1e30);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int i = k == 0 ? 0 : k == 1 ? 99 : min(max(iv + k - 3, 0), 99);   // iv-1 .. iv+2
            const T px = point(last_x, x, step_x, i), py = point(last_y, y, step_y, i);
            margin = A_::fmin_(margin, A_::fmin_(A_::sub(py, lower(px)), A_::sub(upper(px), py)));
        }
        bool inside = margin >= T(0);
        if constexpr (sizeof(T) == 8) {
            if (fabs(margin) < 1e-9) {
                inside = true;
                for (int i = 0; i < 100 && inside; ++i) {
                    const T px = point(last_x, x, step_x, i), py = point(last_y, y, step_y, i);
                    inside = (py >= lower(px)) && (py <= upper(px));
                }
            }
        }
        bool term;
        if (!inside) { term = true; reward = T(-1000); }
        else if (x >= T(9) && T(20) <= y && y <= T(26)) { term = true; reward = A_::div(T(10000), n_step); }
        else {
            term = false;
            const int idx = y <= T(21) ? 20 : (y >= T(27) ? 27 : (int)y);
            reward = A_::div(gradient(!(x < T(0)), idx), n_step);
        }
        s[0] = x; s[1] = y; s[2] = velocity; s[3] = angle; s[4] = n_step;
        return term;
    }
};

// np.linspace(lo, hi, n)[i] in float64
__device__ __forceinline__ double linspace_at(double lo, double hi, int n, int i) {
    if (n <= 1) return lo;
    if (i == n - 1) return hi;
    return __dadd_rn(__dmul_rn((double)i, __ddiv_rn(__dsub_rn(hi, lo), (double)(n - 1))), lo);
}
// DiscreteCurveEnv: itertools.product(linspace(-5, 5, g0), linspace(-30, 30, g1))[k]  (curve_env.py:190-199)
__device__ __forceinline__ void curve_decode(ActionGrid g, int k, double& acceleration, double& steer) {
    acceleration = linspace_at(-5.0, 5.0, g.g0, k / g.g1);
    steer = linspace_at(-30.0, 30.0, g.g1, k % g.g1);
}

// env.step(action) with the action as stored in the tree: (a0) for the gym envs, the grid index or (a0, a1) for CurveEnv
template <int ENV, typename T>
__device__ __forceinline__ bool env_step_action(T (&s)[EnvTraits<ENV>::S], double a0, double a1, ActionGrid g, T& reward) {
    if constexpr (ENV == ENV_CURVE) {
        if (g.g0 > 0) curve_decode(g, (int)a0, a0, a1);
        return Env<ENV_CURVE, T>::step2(s, T(a0), T(a1), reward);
    } else {
        return Env<ENV, T>::step(s, T(a0), reward);
    }
}

// Reward structure.  Envs with kClosedForm pay a constant reward kInterior on every non-terminal step and 0/1 on the
// terminal one; rollout_return() is then the value of `reward += r_t * pow(gamma, t)` (agents/abstract_mcts.py:87)
// after n steps in closed form, bit-identical to the left-to-right float64 sum: CartPole pays 1 on every step
// (host-accumulated prefix table `cum`), FrozenLake pays 0 until the last step (0.0 + ... + r * gamma**(n-1) is exact).
template <int ENV> struct Rewards { static constexpr bool kClosedForm = false; };
// Rows of the thread-per-tree engine's return table (isla_capi.cu: build_return_table): the value a simulation backs up
// from its leaf is one of max_depth + 3 numbers, indexed by leaf_row_*().
template <> struct Rewards<ENV_CARTPOLE> {
    static constexpr bool kClosedForm = true;
    static constexpr double kInterior = 1.0;
    static __device__ __forceinline__ double rollout_return(int n, double, const double*, const double* cum) { return __ldg(cum + n); }
    // row t (0..max_depth+1): 1 + gamma * cum[t], a leaf created with a t-step rollout; a terminal step pays 1 = row 0
    // (cum[0] = 0); row max_depth + 2: 0 (unused by CartPole)
    static __host__ __device__ __forceinline__ int leaf_row_rollout(int t, double) { return t; }
    static __host__ __device__ __forceinline__ int leaf_row_terminal(bool reward_one, int max_depth) { return reward_one ? 0 : max_depth + 2; }
};
template <> struct Rewards<ENV_FROZENLAKE> {
    static constexpr bool kClosedForm = true;
    static constexpr double kInterior = 0.0;
    static __device__ __forceinline__ double rollout_return(int n, double last_r, const double* disc, const double*) {
        return n > 0 ? __dmul_rn(last_r, __ldg(disc + (n - 1))) : 0.0;
    }
    // row 0: 0 (hole / budget exhausted / zero-step rollout); row 1: 1 (the terminal step reaches the goal);
    // row 1 + t (t = 1..max_depth): 0 + gamma * (1 * gamma**(t-1)), a t-step rollout that ends on the goal
    static __host__ __device__ __forceinline__ int leaf_row_rollout(int t, double last_r) { return (t > 0 && last_r != 0.0) ? 1 + t : 0; }
    static __host__ __device__ __forceinline__ int leaf_row_terminal(bool reward_one, int) { return reward_one ? 1 : 0; }
};

// gym reset distributions driven by the Philox reset stream (synthetic roots; bench + tests)
template <int ENV, typename T>
__device__ __forceinline__ void env_reset(uint64_t seed, uint32_t index, T (&s)[EnvTraits<ENV>::S]) {
    uint32_t r[4];
    philox4x32_10(0u, 0u, index, kStreamReset, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    if constexpr (ENV == ENV_CARTPOLE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = T(__dadd_rn(-0.05, __dmul_rn(0.1, u32_to_d01(r[i]))));
    } else if constexpr (ENV == ENV_PENDULUM) {
        s[0] = T(__dadd_rn(-3.141592653589793, __dmul_rn(2 * 3.141592653589793, u32_to_d01(r[0]))));
        s[1] = T(__dadd_rn(-1.0, __dmul_rn(2.0, u32_to_d01(r[1]))));
    } else if constexpr (ENV == ENV_MOUNTAINCAR) {
        s[0] = T(float(__dadd_rn(-0.6, __dmul_rn(0.2, u32_to_d01(r[0])))));
        s[1] = T(0);
    } else if constexpr (ENV == ENV_CURVE) {
        s[0] = T(-11); s[1] = T(21); s[2] = T(10); s[3] = T(90); s[4] = T(0);   // Car(-11, 21), curve_env.py:25-31,181
    } else {
        s[0] = T(0);
    }
}

// env.action_space.sample() from a stream: Discrete -> index, Box -> float32 value
template <int ENV, typename T>
__device__ __forceinline__ T sample_action(Stream& st) {
    if constexpr (EnvTraits<ENV>::kContinuous) {
        float lo, hi;
        env_action_bounds(ENV, lo, hi);
        return T(box_sample_f32(st.next(), lo, hi));
    } else {
        constexpr int A = EnvTraits<ENV>::A;
        if constexpr (A == 2) return T(st.next_bits(1));
        else if constexpr (A == 4) return T(st.next_bits(2));
        else return T(mulhi_u32(st.next(), (uint32_t)A));
    }
}

// FrozenLake is_slippery=True: the action carried out is a-1, a or a+1 (mod 4), chosen like gym's categorical_sample,
// argmax(cumsum([1/3, 1/3, 1/3]) > u)
__device__ __forceinline__ int lake_slip(int a, double u) {
    constexpr double c0 = 1.0 / 3.0, c1 = 1.0 / 3.0 + 1.0 / 3.0;
    const int i = c0 > u ? 0 : (c1 > u ? 1 : 2);
    return (a - 1 + i + 4) & 3;
}

// grid_policy (utils/action_selection_functions.py:258-282): an arg-min of the prior knowledge of the current cell;
// `pos_of(ties)` picks among the tied minima (np.random.choice)
template <typename PosFn>
__device__ __forceinline__ int lake_heuristic_action(const double* __restrict__ table, int cell, PosFn pos_of) {
    double ks[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ks[a] = __ldg(table + cell * 4 + a);
    const double m = fmin(fmin(ks[0], ks[1]), fmin(ks[2], ks[3]));
    int ties = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) ties += (ks[a] == m);
    int pos = pos_of(ties), pick = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) if (ks[a] == m) { if (pos == 0) pick = a; --pos; }
    return pick;
}

// One random-policy rollout step: action_space.sample() from the stream, then env.step
// (agents/abstract_mcts.py:84-85).  CurveEnv samples Box([-5,-30],[5,30], float64) or Discrete(g0*g1).
template <int ENV, typename T>
__device__ __forceinline__ bool rollout_step(Stream& st, T (&s)[EnvTraits<ENV>::S], ActionGrid g, T& reward,
                                             const double* heur = nullptr, const double* noise = nullptr, int step_index = 0) {
    if constexpr (ENV == ENV_FROZENLAKE) {
        if (heur || noise) {
            int a;   // rollout_selection_fn = grid_policy: one stream word only when several actions tie
            if (heur) a = lake_heuristic_action(heur, (int)s[0], [&](int ties) { return ties > 1 ? (int)mulhi_u32(st.next(), (uint32_t)ties) : 0; });
            else a = (int)st.next_bits(2);
            if (noise) a = lake_slip(a, __ldg(noise + step_index));
            return Env<ENV, T>::step(s, T(a), reward);
        }
    }
    if constexpr (ENV == ENV_CURVE) {
        double a0, a1;
        if (g.g0 > 0) {
            const int nd = g.g0 * g.g1;
            const int k = nd == 2 ? (int)st.next_bits(1) : nd == 4 ? (int)st.next_bits(2) : (int)mulhi_u32(st.next(), (uint32_t)nd);
            curve_decode(g, k, a0, a1);
        } else {
            a0 = __dadd_rn(-5.0, __dmul_rn(10.0, u32_to_d01(st.next())));
            a1 = __dadd_rn(-30.0, __dmul_rn(60.0, u32_to_d01(st.next())));
        }
        return Env<ENV_CURVE, T>::step2(s, T(a0), T(a1), reward);
    } else {
        return Env<ENV, T>::step(s, sample_action<ENV, T>(st), reward);
    }
}

// One whole random-policy rollout of a continuous env with a one-dimensional Box action (Pendulum, MountainCarContinuous)
// played by ONE thread, arranged for instruction-level parallelism.  A lone rollout is a dependent chain (state -> cos ->
// velocity -> state ...) that issues one instruction every ~4.5 cycles; the action draws do not depend on the state.  So
// the Philox block of the NEXT four steps is computed in the same straight-line block as the current four steps, and the
// steps are predicated (computed on a copy, committed by selects) instead of branched, which lets the scheduler fill the
// chain's latency slots with the generator.  Same stream words, same operations on the committed steps, same left-to-right
// return sum as the step-by-step loop (rollout_step + r * disc[t]): bit-identical results.
#ifndef ISLA_FAST1D_SPECULATE
#define ISLA_FAST1D_SPECULATE 1
#endif
template <int ENV, typename T>
__device__ __forceinline__ double rollout_fast1d(uint64_t seed, uint32_t c1, uint32_t c2, uint32_t c3, T (&st)[2], int budget,
                                                 const double* __restrict__ disc_table, int& t_out, T& last_r) {
    static_assert(EnvTraits<ENV>::kContinuous && EnvTraits<ENV>::AD == 1 && EnvTraits<ENV>::S == 2, "one-dimensional Box actions");
    float lo, hi;
    env_action_bounds(ENV, lo, hi);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t cur[4], nxt[4];
    philox4x32_10(0u, c1, c2, c3, k0, k1, cur);
    int t = 0;
    double Rr = 0.0;
    T rr = T(0);
    bool go = budget > 0;
    uint32_t blk = 1;
#if ISLA_FAST1D_SPECULATE
    // Common case -- no step of a block ends the episode and the budget covers all four: step without looking, then
    // commit the block at once.  The state chain then carries no select and no predicate, and the loop body is ONE
    // straight-line region: the generator of the next block, the four float64 additions of the PREVIOUS block's terms
    // (27 cycles each on this part) and the state chain interleave.  The first block in which an episode ends (rare: these
    // envs run to the budget) and the last budget % 4 steps are played by the predicated loop below.
    {
        double pend[4] = {0.0, 0.0, 0.0, 0.0};   // terms of the last committed block, not yet added (adding 0.0 changes nothing)
        while (go && t + 4 <= budget) {
            philox4x32_10(blk, c1, c2, c3, k0, k1, nxt);
#pragma unroll
            for (int k = 0; k < 4; ++k) Rr = __dadd_rn(Rr, pend[k]);
            T s4[2] = {st[0], st[1]};
            T rk[4];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) any |= Env<ENV, T>::step(s4, T(box_sample_f32(cur[k], lo, hi)), rk[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) pend[k] = 0.0;
            if (any) break;                        // redone step by step below (cur and blk still name this block)
            st[0] = s4[0]; st[1] = s4[1];
#pragma unroll
            for (int k = 0; k < 4; ++k) pend[k] = __dmul_rn((double)rk[k], __ldg(disc_table + t + k));   // r * pow(gamma, t): added left to right, one block later
            rr = rk[3];
            t += 4;
            go = t < budget;
#pragma unroll
            for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
            ++blk;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) Rr = __dadd_rn(Rr, pend[k]);
    }
#endif
    for (; go; ++blk) {
        philox4x32_10(blk, c1, c2, c3, k0, k1, nxt);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            T ns[2] = {st[0], st[1]};
            T r1;
            const bool d = Env<ENV, T>::step(ns, T(box_sample_f32(cur[k], lo, hi)), r1);
            const double term = __dmul_rn((double)r1, __ldg(disc_table + min(t, budget - 1)));   // r * pow(gamma, t)
            if (go) {
                st[0] = ns[0]; st[1] = ns[1];
                rr = r1;
                Rr = __dadd_rn(Rr, term);
                ++t;
                go = !d && t < budget;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
    }
    t_out = t;
    last_r = rr;
    return Rr;
}

// One whole random-policy CartPole rollout by one thread in blocks of NB predicated steps (computed on a copy, committed
// by selects): the tail of one step overlaps the pole-angle chain of the next.  Step t draws bit t of the stream (most
// significant bit of word t / 32 first, exactly next_bits(1)); a block never straddles a word.  Returns the step count;
// the return follows from it in closed form (Rewards<ENV_CARTPOLE>).
template <typename T, int NB>
__device__ __forceinline__ int rollout_cartpole_blocks(uint64_t seed, uint32_t c1, uint32_t c2, uint32_t c3, T (&st)[4], int budget, T& last_r) {
    static_assert(32 % NB == 0, "a block of steps must not straddle a stream word");
    Stream rs;
    rs.init(seed, c1, c2, c3);
    int t = 0;
    bool go = budget > 0;
    uint32_t word = 0u;
    T rr = T(0);
    while (go) {
        if ((t & 31) == 0) word = rs.next();
        uint32_t bits = word << (t & 31);
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            T ns[4] = {st[0], st[1], st[2], st[3]};
            T r1;
            const bool done = Env<ENV_CARTPOLE, T>::step(ns, T(bits >> 31), r1);
            bits <<= 1;
            if (go) {
                st[0] = ns[0]; st[1] = ns[1]; st[2] = ns[2]; st[3] = ns[3];
                rr = r1;
                ++t;
                go = !done && t < budget;
            }
        }
    }
    last_r = rr;
    return t;
}

}  // namespace isla
