/* TEST INFRASTRUCTURE ONLY -- see isla_oracle.h.
 *
 * Plain-C restatement of the IslaMcts simulate loop (reference = pure Python under
 * /root/reference/islaMcts).  Build: `make -C oracle/c` (-ffp-contract=off so that
 * every product/sum rounds separately, like CPython/numpy float64 arithmetic).
 *
 * Structure mirrors the reference's mutual recursion so that each branch can be
 * checked against the cited lines:
 *   state_visit()  <-> StateNode*.build_tree_state   (mcts.py:46-67, mcts_hash.py:59-83,
 *                      mcts_action_...:52-84, mcts_state_...:47-67, mcts_double_...:49-75)
 *   action_visit() <-> ActionNode*.build_tree_action (mcts.py:72-120, mcts_hash.py:87-131,
 *                      mcts_action_...:89-136, mcts_state_...:72-130, mcts_double_...:79-135)
 *   do_rollout()   <-> AbstractStateNode.rollout     (abstract_mcts.py:72-91)
 *   ucb_pick()     <-> ucb1                          (utils/action_selection_functions.py:9-25)
 *
 * Randomness comes from one of two sources: a recorded decision trace of the
 * reference (scripted mode) or Philox4x32-10 counter streams (free-running mode; the
 * CUDA engine uses the same streams so whole searches can be compared bit for bit).
 */
#include "isla_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_DISCRETE 128
#define ORC_MAX_S 8

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                    uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t c1, c2, c3, k0, k1; /* fixed words */
    uint32_t blk;                /* c0: block counter */
    uint32_t buf[4];
    int idx;
    uint32_t bitbuf; int bits_left; /* reservoir of stream_bits() */
} stream_t;

static void stream_init(stream_t* s, uint64_t seed, uint32_t c1, uint32_t c2, uint32_t c3) {
    s->k0 = (uint32_t)seed; s->k1 = (uint32_t)(seed >> 32);
    s->c1 = c1; s->c2 = c2; s->c3 = c3; s->blk = 0; s->idx = 4; s->bitbuf = 0; s->bits_left = 0;
}
static uint32_t stream_u32(stream_t* s) {
    if (s->idx == 4) { orc_philox4x32(s->blk++, s->c1, s->c2, s->c3, s->k0, s->k1, s->buf); s->idx = 0; }
    return s->buf[s->idx++];
}
/* k bits (k divides 32), most significant first */
static uint32_t stream_bits(stream_t* s, int k) {
    if (s->bits_left == 0) { s->bitbuf = stream_u32(s); s->bits_left = 32; }
    uint32_t v = s->bitbuf >> (32 - k);
    s->bitbuf <<= k; s->bits_left -= k;
    return v;
}
#define STREAM_TREE  0x80000000u
#define STREAM_RESET 0xC0000000u

static inline uint32_t mulhi32(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * n) >> 32); }
static inline float u32_to_f01(uint32_t u) { return (float)(u >> 8) * 5.9604644775390625e-08f; } /* 2^-24 */
static inline double u32_to_d01(uint32_t u) { return (double)u * 2.3283064365386963e-10; }       /* 2^-32 */
static inline float box_sample_f32(uint32_t u, float low, float high) {
    float w = high - low; float m = u32_to_f01(u) * w; return low + m;
}

/* ------------------------------------------------------------------ environments */
/* Gymnasium definitions restated in oracle/pyenvs.py; float64. */
int orc_state_dim(int env_id) {
    switch (env_id) { case ORC_ENV_CARTPOLE: return 4; case ORC_ENV_PENDULUM: return 2;
                      case ORC_ENV_MOUNTAINCAR: return 2; case ORC_ENV_FROZENLAKE: return 1; case ORC_ENV_CURVE: return 5; }
    return 0;
}
static int obs_dim(int env_id) { return env_id == ORC_ENV_PENDULUM ? 3 : env_id == ORC_ENV_CURVE ? 4 : orc_state_dim(env_id); }
int orc_action_dim(int env_id) { return env_id == ORC_ENV_CURVE ? 2 : 1; }

static const char LAKE[16] = {'S','F','F','F','F','H','F','H','F','F','F','H','H','F','F','G'};

/* the observation the reference keys its children dicts with (float32 values for the gym envs, the float64
 * car.state for CurveEnv), held as doubles */
static void make_obs(int env_id, const double* s, double* obs) {
    switch (env_id) {
    case ORC_ENV_CARTPOLE: for (int i = 0; i < 4; ++i) obs[i] = (double)(float)s[i]; break;
    case ORC_ENV_PENDULUM: obs[0] = (double)(float)cos(s[0]); obs[1] = (double)(float)sin(s[0]); obs[2] = (double)(float)s[1]; break;
    case ORC_ENV_MOUNTAINCAR: obs[0] = (double)(float)s[0]; obs[1] = (double)(float)s[1]; break;
    case ORC_ENV_CURVE: for (int i = 0; i < 4; ++i) obs[i] = s[i]; break;
    default: obs[0] = (double)(float)s[0]; break;
    }
}

/* CurveEnv reward tables: islaMcts/environments/curve_env.py:61-72 */
static double curve_gradient(int second, int idx) {
    if (idx >= 27) return 50.0;
    if (!second) return 1.80 * idx;
    return 1.80 * (29 + (26 - idx));   /* idx 26 -> 1.8*29, 25 -> 1.8*30, ..., 20 -> 1.8*35 */
}

int orc_env_step(int env_id, const double* s, const double* act, double* ns, double* reward, int32_t* term,
                 double* obs) {
    const double action = act[0];
    switch (env_id) {
    case ORC_ENV_CARTPOLE: {
        const double gravity = 9.8, masscart = 1.0, masspole = 0.1, total_mass = masspole + masscart,
                     length = 0.5, polemass_length = masspole * length, force_mag = 10.0, tau = 0.02,
                     th_lim = 12 * 2 * M_PI / 360, x_lim = 2.4;
        double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
        double force = ((int)action == 1) ? force_mag : -force_mag;
        double costheta = cos(theta), sintheta = sin(theta);
        double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
        double thetaacc = (gravity * sintheta - costheta * temp) /
                          (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
        double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
        ns[0] = x + tau * x_dot; ns[1] = x_dot + tau * xacc;
        ns[2] = theta + tau * theta_dot; ns[3] = theta_dot + tau * thetaacc;
        *term = (ns[0] < -x_lim || ns[0] > x_lim || ns[2] < -th_lim || ns[2] > th_lim);
        *reward = 1.0;
        break;
    }
    case ORC_ENV_PENDULUM: {
        const double max_speed = 8.0, max_torque = 2.0, dt = 0.05, g = 10.0, m = 1.0, l = 1.0;
        double th = s[0], thdot = s[1];
        double u = fmin(fmax(action, -max_torque), max_torque);
        double ang = fmod(th + M_PI, 2 * M_PI);
        if (ang < 0) ang += 2 * M_PI;
        ang -= M_PI;
        double costs = ang * ang + 0.1 * (thdot * thdot) + 0.001 * (u * u);
        double newthdot = thdot + (3 * g / (2 * l) * sin(th) + 3.0 / (m * (l * l)) * u) * dt;
        newthdot = fmin(fmax(newthdot, -max_speed), max_speed);
        ns[0] = th + newthdot * dt; ns[1] = newthdot;
        *term = 0; *reward = -costs;
        break;
    }
    case ORC_ENV_MOUNTAINCAR: {
        const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.45,
                     goal_velocity = 0.0, power = 0.0015;
        double position = s[0], velocity = s[1];
        double force = fmin(fmax(action, -1.0), 1.0);
        velocity += force * power - 0.0025 * cos(3 * position);
        velocity = fmin(fmax(velocity, -max_speed), max_speed);
        position += velocity;
        position = fmin(fmax(position, min_position), max_position);
        if (position == min_position && velocity < 0) velocity = 0.0;
        *term = (position >= goal_position && velocity >= goal_velocity);
        *reward = (*term ? 100.0 : 0.0) - (action * action) * 0.1;
        ns[0] = (double)(float)position; ns[1] = (double)(float)velocity; /* state kept as float32 */
        break;
    }
    case ORC_ENV_FROZENLAKE: {
        int cell = (int)s[0], row = cell / 4, col = cell % 4, a = (int)action;
        if (LAKE[cell] == 'G' || LAKE[cell] == 'H') { ns[0] = s[0]; *reward = 0.0; *term = 1; break; }
        if (a == 0) col = col > 0 ? col - 1 : 0;
        else if (a == 1) row = row < 3 ? row + 1 : 3;
        else if (a == 2) col = col < 3 ? col + 1 : 3;
        else if (a == 3) row = row > 0 ? row - 1 : 0;
        cell = row * 4 + col;
        ns[0] = (double)cell; *reward = LAKE[cell] == 'G' ? 1.0 : 0.0; *term = (LAKE[cell] == 'G' || LAKE[cell] == 'H');
        break;
    }
    case ORC_ENV_CURVE: {
        /* islaMcts/environments/curve_env.py: Car.update :34-56, CurveEnv.step :120-174.  state = (x, y, velocity,
         * angle, n_step); the step counter is hidden state of the env object (it scales the rewards). */
        const double dt = 0.5;
        double x = s[0], y = s[1], velocity = s[2], angle = s[3];
        const double n_step = s[4] + 1.0;
        const double last_x = x, last_y = y;
        velocity += act[0] * dt;
        if (velocity > 20.0) velocity = 20.0;
        if (velocity < 0.0) velocity = 0.0;
        angle += act[1];
        angle = fmod(angle, 360.0);
        if (angle < 0) angle += 360.0;                       /* python float % */
        const double rad = angle * (M_PI / 180.0);           /* math.radians */
        const double vel_x = cos(rad) * velocity * dt, vel_y = sin(rad) * velocity * dt;
        x += vel_x; y += vel_y;
        /* 100 points of np.linspace(last, new) must lie inside the corridor (:136-143) */
        int inside = 1;
        const double step_x = (x - last_x) / 99.0, step_y = (y - last_y) / 99.0;
        for (int i = 0; i < 100 && inside; ++i) {
            const double px = i == 99 ? x : (double)i * step_x + last_x;
            const double py = i == 99 ? y : (double)i * step_y + last_y;
            const double lower = -(1.0 / 50) * (3 * (px * px)) + 27.3;
            const double higher = -(1.0 / 50) * (px * px) + 27.5;
            if (!(py >= lower) || !(py <= higher)) inside = 0;
        }
        if (!inside) { *term = 1; *reward = -1000.0; }
        else if (x >= 9 && 20 <= y && y <= 26) { *term = 1; *reward = 10000.0 / n_step; }
        else {
            *term = 0;
            const int second = !(x < 0);
            const int idx = y <= 21 ? 20 : (y >= 27 ? 27 : (int)y);
            *reward = curve_gradient(second, idx) / n_step;
        }
        ns[0] = x; ns[1] = y; ns[2] = velocity; ns[3] = angle; ns[4] = n_step;
        break;
    }
    default: return -1;
    }
    if (obs) make_obs(env_id, ns, obs);
    return 0;
}

void orc_env_reset(int env_id, uint64_t seed, uint64_t index, double* state) {
    uint32_t r[4];
    orc_philox4x32(0, 0, (uint32_t)index, STREAM_RESET, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    switch (env_id) {
    case ORC_ENV_CARTPOLE: for (int i = 0; i < 4; ++i) state[i] = -0.05 + 0.1 * u32_to_d01(r[i]); break;
    case ORC_ENV_PENDULUM: state[0] = -M_PI + 2 * M_PI * u32_to_d01(r[0]); state[1] = -1.0 + 2.0 * u32_to_d01(r[1]); break;
    case ORC_ENV_MOUNTAINCAR: state[0] = (double)(float)(-0.6 + 0.2 * u32_to_d01(r[0])); state[1] = 0.0; break;
    case ORC_ENV_CURVE: state[0] = -11.0; state[1] = 21.0; state[2] = 10.0; state[3] = 90.0; state[4] = 0.0; break;  /* Car(-11, 21), curve_env.py:181 */
    default: state[0] = 0.0; break;
    }
}

static void action_bounds(int env_id, float* low, float* high) {
    if (env_id == ORC_ENV_PENDULUM) { *low = -2.0f; *high = 2.0f; } else { *low = -1.0f; *high = 1.0f; }
}
static int env_n_actions(int env_id) { return env_id == ORC_ENV_CARTPOLE ? 2 : env_id == ORC_ENV_FROZENLAKE ? 4 : 0; }

/* ------------------------------------------------------------------ the tree */
typedef struct {
    int64_t ns; double total; int terminal; double terminal_reward;
    double obs[ORC_MAX_S]; /* node.data: the observation (key of the parent's children dict) */
    int first_act, last_act, n_act;   /* actions dict, insertion order */
    int next_sibling;      /* next entry of the parent's children dict */
    int blk;               /* instrumentation only: order in which nodes were first descended into */
} snode;

typedef struct {
    int64_t na; double total; double value[2]; /* node.data (1 or 2 action dimensions) */
    int64_t visits;        /* apw/dpw: parent's visit_actions[key] */
    int first_child, last_child, n_child;   /* children dict, insertion order */
    int next_act;          /* next entry of the parent's actions dict */
} anode;

struct orc_tree {
    orc_config cfg;
    double root_state[ORC_MAX_S];
    snode* S; anode* A; int nS, nA, capS, capA;
    /* the live environment of the current simulation (param.env) */
    double env[ORC_MAX_S];
    /* randomness */
    const uint8_t* tk; const double* tv; int64_t tlen, tpos; int fault;
    stream_t tree_stream, roll_stream;
    uint32_t sim_index; /* cumulative over run() calls */
    double* heur; int heur_states, heur_actions;   /* grid_policy prior knowledge [state][action], NULL = random rollouts */
    /* FrozenLake is_slippery=True: every simulation runs on a pickled clone of the env INCLUDING its generator state
     * (utils/mcts_utils.py:8-16), so inside one fit() the k-th env.step of every simulation draws the same number:
     * noise[k], k = env.step calls so far in the simulation.  NULL = deterministic env. */
    double* noise; int noise_n, sim_step;
    int64_t c_env_steps, c_roll_steps, c_levels, c_sims, c_expand_steps, c_seq; int next_blk, cur_parent_blk;
    /* instrumentation only: common prefix of consecutive simulation paths */
    int prev_path[4096], cur_path[4096], prev_len, cur_len; int64_t c_common, c_prevlen;
};

static int new_snode(orc_tree* t, const double* obs) {
    if (t->nS == t->capS) { t->capS = t->capS * 2 + 16; t->S = (snode*)realloc(t->S, sizeof(snode) * t->capS); }
    snode* s = &t->S[t->nS]; memset(s, 0, sizeof *s);
    s->first_act = s->last_act = -1; s->next_sibling = -1;
    memcpy(s->obs, obs, sizeof(double) * ORC_MAX_S);
    return t->nS++;
}
static int new_anode(orc_tree* t, const double* value) {
    if (t->nA == t->capA) { t->capA = t->capA * 2 + 16; t->A = (anode*)realloc(t->A, sizeof(anode) * t->capA); }
    anode* a = &t->A[t->nA]; memset(a, 0, sizeof *a);
    a->first_child = a->last_child = -1; a->next_act = -1; a->value[0] = value[0]; a->value[1] = value[1];
    return t->nA++;
}

/* ---- random source */
static double trace_next(orc_tree* t, int kind) {
    if (t->fault) return 0.0;
    if (t->tpos >= t->tlen) { t->fault = -1; return 0.0; }
    if (t->tk[t->tpos] != kind) { t->fault = -2; return 0.0; }
    return t->tv[t->tpos++];
}
static int scripted(const orc_tree* t) { return t->tlen >= 0; }

/* np.random.choice(candidates): returns the POSITION inside the candidate list */
static int rnd_choice(orc_tree* t, int n) {
    if (scripted(t)) {
        int p = (int)trace_next(t, ORC_K_CHOICE);
        if (p < 0 || p >= n) { if (!t->fault) t->fault = -3; return 0; }
        return p;
    }
    if (n <= 1) return 0; /* Philox mode: a single candidate consumes no draw */
    return (int)mulhi32(stream_u32(&t->tree_stream), (uint32_t)n);
}
/* number of discrete actions of the run (0 = continuous Box): DiscreteCurveEnv has grid0 x grid1 of them */
static int n_discrete(const orc_tree* t) {
    if (t->cfg.env_id == ORC_ENV_CURVE) return t->cfg.grid0 > 0 ? t->cfg.grid0 * t->cfg.grid1 : 0;
    return env_n_actions(t->cfg.env_id);
}
/* env.action_space.sample(): out[0] (and out[1] for the 2-D CurveEnv box) */
static void rnd_action(orc_tree* t, stream_t* st, double out[2]) {
    int nd = n_discrete(t);
    out[0] = 0; out[1] = 0;
    if (nd > 0) {
        if (scripted(t)) {
            int a = (int)trace_next(t, ORC_K_SAMPLE_D);
            if (a < 0 || a >= nd) { if (!t->fault) t->fault = -3; return; }
            out[0] = a; return;
        }
        /* Discrete(2) / Discrete(4): 1 / 2 bits of a stream word, most significant first */
        if (nd == 2) out[0] = (double)stream_bits(st, 1);
        else if (nd == 4) out[0] = (double)stream_bits(st, 2);
        else out[0] = (double)mulhi32(stream_u32(st), (uint32_t)nd);
        return;
    }
    if (t->cfg.env_id == ORC_ENV_CURVE) {
        /* Box(low=[-5,-30], high=[5,30], dtype=float64): np_random.uniform per dimension */
        if (scripted(t)) { out[0] = trace_next(t, ORC_K_SAMPLE_C); out[1] = trace_next(t, ORC_K_SAMPLE_C); return; }
        out[0] = -5.0 + 10.0 * u32_to_d01(stream_u32(st));
        out[1] = -30.0 + 60.0 * u32_to_d01(stream_u32(st));
        return;
    }
    if (scripted(t)) { out[0] = (double)(float)trace_next(t, ORC_K_SAMPLE_C); return; }
    float low, high; action_bounds(t->cfg.env_id, &low, &high);
    out[0] = (double)box_sample_f32(stream_u32(st), low, high);
}
/* random.choices(population, weights)[0] -> index */
static int rnd_choices(orc_tree* t, const double* w, int n) {
    if (scripted(t)) {
        int p = (int)trace_next(t, ORC_K_CHOICES);
        if (p < 0 || p >= n) { if (!t->fault) t->fault = -3; return 0; }
        return p;
    }
    double tot = 0; for (int i = 0; i < n; ++i) tot += w[i];
    double u = u32_to_d01(stream_u32(&t->tree_stream)) * tot, cum = 0;
    for (int i = 0; i < n; ++i) { cum += w[i]; if (cum > u) return i; }
    return n - 1;
}
static double rnd_u01(orc_tree* t) {
    if (scripted(t)) return trace_next(t, ORC_K_U01);
    return u32_to_d01(stream_u32(&t->tree_stream));
}
static double rnd_uniform(orc_tree* t, double low, double high) {
    if (scripted(t)) return trace_next(t, ORC_K_UNIFORM);
    return low + (high - low) * u32_to_d01(stream_u32(&t->tree_stream));
}

/* ---- env of the running simulation */
/* np.linspace(lo, hi, n)[i] */
static double linspace_at(double lo, double hi, int n, int i) {
    if (n <= 1) return lo;
    if (i == n - 1) return hi;
    return (double)i * ((hi - lo) / (double)(n - 1)) + lo;
}
static void env_step(orc_tree* t, const double* action, double* obs, double* r, int* term) {
    double ns[ORC_MAX_S], act[2] = {action[0], action[1]}; int32_t tm;
    if (t->cfg.env_id == ORC_ENV_CURVE && t->cfg.grid0 > 0) {
        /* DiscreteCurveEnv: itertools.product(linspace(acc), linspace(angle))[k]  (curve_env.py:187-201) */
        const int k = (int)action[0];
        act[0] = linspace_at(-5.0, 5.0, t->cfg.grid0, k / t->cfg.grid1);
        act[1] = linspace_at(-30.0, 30.0, t->cfg.grid1, k % t->cfg.grid1);
    }
    if (t->noise && t->cfg.env_id == ORC_ENV_FROZENLAKE) {
        if (t->sim_step >= t->noise_n) { if (!t->fault) t->fault = -3; }
        else {
            /* gym categorical_sample: argmax(cumsum([1/3, 1/3, 1/3]) > u); carried-out action = (a - 1 + i) mod 4 */
            const double u = t->noise[t->sim_step], c0 = 1.0 / 3.0, c1 = c0 + 1.0 / 3.0;
            const int i = c0 > u ? 0 : (c1 > u ? 1 : 2);
            act[0] = (double)(((int)action[0] - 1 + i + 4) % 4);
        }
        t->sim_step++;
    }
    orc_env_step(t->cfg.env_id, t->env, act, ns, r, &tm, obs);
    memcpy(t->env, ns, sizeof(double) * ORC_MAX_S);
    *term = tm; t->c_env_steps++;
}
/* SPW/DPW poke the node's observation into the env (mcts_state_...:24-25,127;
 * mcts_double_...:25-26,68-69,131-132): valid when obs == state (float32 values). */
static void env_set_from_obs(orc_tree* t, const double* obs) {
    int d = orc_state_dim(t->cfg.env_id);
    for (int i = 0; i < d; ++i) t->env[i] = obs[i];
}

/* abstract_mcts.py:72-91.  `budget` = number of steps the loop condition
 * `curr_depth + starting_depth < max_depth` admits.  R rollouts (build extension,
 * SURVEY 8a) share the leaf state; rollout j draws from stream (seed, tree, sim, j). */
/* grid_policy (utils/action_selection_functions.py:258-282): the rollout action is an arg-min of the prior knowledge of
 * the env's current cell, ties broken by np.random.choice over the tied actions (K_CHOICE; Philox: one tree-independent
 * rollout-stream word when there are several candidates). */
static double heuristic_action(orc_tree* t, stream_t* st) {
    const int cell = (int)t->env[0], A = t->heur_actions;
    const double* ks = t->heur + (size_t)cell * A;
    double m = INFINITY; for (int a = 0; a < A; ++a) if (ks[a] < m) m = ks[a];
    int ties = 0; for (int a = 0; a < A; ++a) ties += (ks[a] == m);
    int pos;
    if (scripted(t)) {
        pos = (int)trace_next(t, ORC_K_CHOICE);
        if (pos < 0 || pos >= ties) { if (!t->fault) t->fault = -3; pos = 0; }
    } else pos = ties > 1 ? (int)mulhi32(stream_u32(st), (uint32_t)ties) : 0;
    for (int a = 0; a < A; ++a) if (ks[a] == m) { if (pos-- == 0) return (double)a; }
    return 0.0;
}

static double do_rollout(orc_tree* t, int budget) {
    const int R = t->cfg.rollouts_per_leaf > 0 ? t->cfg.rollouts_per_leaf : 1;
    int P = 1; while (P < R) P <<= 1;
    double leaf[ORC_MAX_S]; memcpy(leaf, t->env, sizeof leaf);
    double* x = (double*)calloc((size_t)P, sizeof(double));
    const int leaf_step = t->sim_step;
    for (int j = 0; j < R; ++j) {
        memcpy(t->env, leaf, sizeof leaf);
        t->sim_step = leaf_step;   /* every rollout of a leaf batch continues the clone's noise from the leaf */
        if (!scripted(t)) stream_init(&t->roll_stream, t->cfg.seed, t->sim_index, (uint32_t)t->cfg.tree_id, (uint32_t)j);
        int done = 0, depth = 0; double reward = 0;
        while (!done && depth < budget) {
            double a[2] = {0.0, 0.0};
            if (t->heur) a[0] = heuristic_action(t, &t->roll_stream); else rnd_action(t, &t->roll_stream, a);
            double obs[ORC_MAX_S]; double r; int term;
            env_step(t, a, obs, &r, &term);
            reward += r * pow(t->cfg.gamma, depth);
            done = term; depth++; t->c_roll_steps++;
        }
        x[j] = reward;
    }
    /* R > 1 is a build extension with no reference: the batch mean is DEFINED as the fixed
     * power-of-two pairwise tree below (zero padded), divided by R -- the order the CUDA engine uses. */
    for (int stride = P >> 1; stride > 0; stride >>= 1)
        for (int i = 0; i < stride; ++i) x[i] = x[i] + x[i + stride];
    double sum = x[0];
    free(x);
    return R == 1 ? sum : sum / R;
}

static int budget_at(const orc_tree* t, int level_of_new_state) {
    if (t->cfg.budget_mode == ORC_BUDGET_ZERO) return 0;
    int b = t->cfg.max_depth - level_of_new_state;
    return b > 0 ? b : 0;
}

/* action_selection_functions.py:9-25; returns the child RANK (position in node.actions) */
static int ucb_pick(orc_tree* t, int sidx) {
    snode* s = &t->S[sidx];
    double lg = log((double)s->ns), best = -INFINITY;
    int n = s->n_act, ties = 0;
    double* score = (double*)malloc(sizeof(double) * n);
    int i = 0;
    for (int a = s->first_act; a >= 0; a = t->A[a].next_act, ++i) {
        anode* c = &t->A[a];
        double q = c->total / (double)c->na;
        double e = t->cfg.C * sqrt(lg / (double)c->na);
        score[i] = q + e;
        if (score[i] > best) best = score[i];
    }
    for (i = 0; i < n; ++i) ties += (score[i] == best);
    int pos = rnd_choice(t, ties), rank = 0;
    for (i = 0; i < n; ++i) if (score[i] == best) { if (pos-- == 0) { rank = i; break; } }
    free(score);
    return rank;
}
static int act_by_rank(const orc_tree* t, int sidx, int rank) {
    int a = t->S[sidx].first_act; while (rank-- > 0) a = t->A[a].next_act; return a;
}
static void append_action(orc_tree* t, int sidx, int aidx) {
    snode* s = &t->S[sidx];
    if (s->last_act < 0) s->first_act = aidx; else t->A[s->last_act].next_act = aidx;
    s->last_act = aidx; s->n_act++;
}
/* dict[key] = node where key exists: the entry keeps its position (python dict) */
static void replace_action(orc_tree* t, int sidx, int old, int neu) {
    snode* s = &t->S[sidx];
    t->A[neu].next_act = t->A[old].next_act;
    if (s->first_act == old) s->first_act = neu;
    else { int p = s->first_act; while (t->A[p].next_act != old) p = t->A[p].next_act; t->A[p].next_act = neu; }
    if (s->last_act == old) s->last_act = neu;
}
static int find_child(const orc_tree* t, int aidx, const double* obs) {
    int od = obs_dim(t->cfg.env_id);
    for (int c = t->A[aidx].first_child; c >= 0; c = t->S[c].next_sibling)
        if (memcmp(t->S[c].obs, obs, sizeof(double) * od) == 0) return c;
    return -1;
}
static void append_child(orc_tree* t, int aidx, int sidx) {
    anode* a = &t->A[aidx];
    if (a->last_child < 0) a->first_child = sidx; else t->S[a->last_child].next_sibling = sidx;
    a->last_child = sidx; a->n_child++;
}
static void replace_child(orc_tree* t, int aidx, int old, int neu) {
    anode* a = &t->A[aidx];
    t->S[neu].next_sibling = t->S[old].next_sibling;
    if (a->first_child == old) a->first_child = neu;
    else { int p = a->first_child; while (t->S[p].next_sibling != old) p = t->S[p].next_sibling; t->S[p].next_sibling = neu; }
    if (a->last_child == old) a->last_child = neu;
}
/* actions dict lookup by action.tobytes(): float32 samples are stored widened to double, so
 * comparing the double bits is the same test (voo yields float64 actions directly) */
static int find_action_bits(const orc_tree* t, int sidx, const double* value) {
    const int ad = orc_action_dim(t->cfg.env_id);
    for (int a = t->S[sidx].first_act; a >= 0; a = t->A[a].next_act)
        if (memcmp(value, t->A[a].value, sizeof(double) * ad) == 0) return a;
    return -1;
}

/* Box bounds of the env's action space widened to double, per dimension; gym boxes are float32, CurveEnv's is float64 */
static int box_bounds(const orc_tree* t, double lo[2], double hi[2]) {
    if (t->cfg.env_id == ORC_ENV_CURVE) { lo[0] = -5.0; hi[0] = 5.0; lo[1] = -30.0; hi[1] = 30.0; return 0; }
    float lowf, highf; action_bounds(t->cfg.env_id, &lowf, &highf);
    lo[0] = (double)lowf; hi[0] = (double)highf; lo[1] = hi[1] = 0.0;
    return 1; /* float32 box */
}
/* scipy cdist(..., 'euclidean') on float64 rows: sqrt(sum (u_i - v_i)^2); one dimension: sqrt(d*d) == |d| */
static double euclid(const double* u, const double* v, int D) {
    if (D == 1) return fabs(u[0] - v[0]);
    const double d0 = u[0] - v[0], d1 = u[1] - v[1];
    return sqrt(d0 * d0 + d1 * d1);
}

/* voo: utils/action_selection_functions.py:139-205 (one or two action dimensions) */
static void voo_expand(orc_tree* t, int sidx, double out[2]) {
    snode* s = &t->S[sidx];
    const int D = orc_action_dim(t->cfg.env_id);
    out[0] = out[1] = 0.0;
    double p = rnd_u01(t);
    if (p <= t->cfg.epsilon || s->n_act == 0) { rnd_action(t, &t->tree_stream, out); return; }
    int n = s->n_act;
    double* act = (double*)malloc(sizeof(double) * 2 * n);
    double* q = (double*)malloc(sizeof(double) * n);
    int i = 0; double qmax = -INFINITY;
    for (int a = s->first_act; a >= 0; a = t->A[a].next_act, ++i) {
        act[2 * i] = t->A[a].value[0]; act[2 * i + 1] = t->A[a].value[1];
        q[i] = t->A[a].total / (double)t->A[a].na; if (q[i] > qmax) qmax = q[i];
    }
    int ties = 0; for (i = 0; i < n; ++i) ties += (q[i] == qmax);
    int pos = rnd_choice(t, ties), best = 0;
    for (i = 0; i < n; ++i) if (q[i] == qmax) { if (pos-- == 0) { best = i; break; } }
    double lo[2], hi[2]; box_bounds(t, lo, hi);
    const int ns_ = t->cfg.voo_n_samples, max_try = t->cfg.voo_max_try;
    double* samp = (double*)malloc(sizeof(double) * 2 * (ns_ > 0 ? ns_ : 1));
    double* sdist = (double*)malloc(sizeof(double) * (ns_ > 0 ? ns_ : 1));
    double* tried = (double*)malloc(sizeof(double) * 2 * (max_try + 1));
    double* dtried = (double*)malloc(sizeof(double) * (max_try + 1));
    int n_acc = 0;
    for (int k = 0; k < ns_; ++k) {
        int n_try = 0;
        for (;;) {
            double x[2] = {0.0, 0.0};
            for (int d = 0; d < D; ++d) x[d] = rnd_uniform(t, lo[d], hi[d]);   /* np.random.uniform(low, high, shape) */
            double db = euclid(x, act + 2 * best, D);
            tried[2 * n_try] = x[0]; tried[2 * n_try + 1] = x[1]; dtried[n_try] = db;
            if (db == 0) break; /* :177-178 exits WITHOUT appending */
            int ok = 1; for (i = 0; i < n; ++i) if (i != best && !(db <= euclid(x, act + 2 * i, D))) { ok = 0; break; }
            if (ok) { samp[2 * n_acc] = x[0]; samp[2 * n_acc + 1] = x[1]; sdist[n_acc] = db; n_acc++; break; }
            n_try++;
            if (n_try >= max_try) {
                double dm = INFINITY; for (i = 0; i < max_try; ++i) if (dtried[i] < dm) dm = dtried[i];
                int tt = 0; for (i = 0; i < max_try; ++i) tt += (dtried[i] == dm);
                int pp = rnd_choice(t, tt), pick = 0;
                for (i = 0; i < max_try; ++i) if (dtried[i] == dm) { if (pp-- == 0) { pick = i; break; } }
                samp[2 * n_acc] = tried[2 * pick]; samp[2 * n_acc + 1] = tried[2 * pick + 1]; sdist[n_acc] = dtried[pick]; n_acc++;
                break;
            }
        }
        if (t->fault) break;
    }
    out[0] = act[2 * best]; out[1] = act[2 * best + 1];
    if (n_acc > 0) {
        double dm = INFINITY; for (i = 0; i < n_acc; ++i) if (sdist[i] < dm) dm = sdist[i];
        int tt = 0; for (i = 0; i < n_acc; ++i) tt += (sdist[i] == dm);
        int pp = rnd_choice(t, tt);
        for (i = 0; i < n_acc; ++i) if (sdist[i] == dm) { if (pp-- == 0) { out[0] = samp[2 * i]; out[1] = samp[2 * i + 1]; break; } }
    }
    free(act); free(q); free(samp); free(sdist); free(tried); free(dtried);
    /* NOTE: a float64 action; the reference keys it by its float64 bytes */
}

/* genetic_policy2: utils/action_selection_functions.py:104-136 (one or two action dimensions; float32 gym boxes,
 * float64 CurveEnv box).  p is drawn first on every call; 0/1/2 existing actions get the centre / a random bound /
 * the missing bound; later calls take default_policy with probability 1-epsilon, else the mean of the two LOWEST-q
 * actions (the list is sorted ascending and [0], [1] are used, :108-112). */
static void genetic2_expand(orc_tree* t, int sidx, double out[2]) {
    snode* s = &t->S[sidx];
    const int D = orc_action_dim(t->cfg.env_id);
    const double p = rnd_u01(t);
    double lo[2], hi[2];
    const int f32 = box_bounds(t, lo, hi);
    out[0] = out[1] = 0.0;
    if (s->n_act == 0) {                                                               /* np.mean([low, high], axis=0) */
        for (int d = 0; d < D; ++d) out[d] = f32 ? (double)(((float)lo[d] + (float)hi[d]) / 2.0f) : (lo[d] + hi[d]) / 2.0;
        return;
    }
    if (s->n_act == 1) {
        double w[2] = {1.0, 1.0};
        const int pick_high = rnd_choices(t, w, 2) == 0;                               /* random.choices([high, low]) */
        for (int d = 0; d < D; ++d) out[d] = pick_high ? hi[d] : lo[d];
        return;
    }
    if (s->n_act == 2) {
        const int has_high = find_action_bits(t, sidx, hi) >= 0;
        for (int d = 0; d < D; ++d) out[d] = has_high ? lo[d] : hi[d];
        return;
    }
    if (p < 1 - t->cfg.epsilon) { rnd_action(t, &t->tree_stream, out); return; }
    /* stable ascending sort by q: the first two entries */
    int b0 = -1, b1 = -1; double q0 = INFINITY, q1 = INFINITY;
    for (int a = s->first_act; a >= 0; a = t->A[a].next_act) {
        double q = t->A[a].total / (double)t->A[a].na;
        if (q < q0) { b1 = b0; q1 = q0; b0 = a; q0 = q; }
        else if (q < q1) { b1 = a; q1 = q; }
    }
    for (int d = 0; d < D; ++d) {                                                      /* np.mean([best, second], axis=0) */
        if (f32) { float x = (float)t->A[b0].value[d], y = (float)t->A[b1].value[d]; out[d] = (double)((x + y) / 2.0f); }
        else out[d] = (t->A[b0].value[d] + t->A[b1].value[d]) / 2.0;
    }
}

/* genetic_policy: utils/action_selection_functions.py:50-102 (one or two action dimensions).  0 / 1 / 2 existing
 * actions get the centre / a random bound / the missing bound (no draw of p); later calls draw p and, if p < epsilon,
 * sample n_samples points uniformly in the box spanned by the two LOWEST-q actions (sorted ascending, [0] and [1],
 * :56-59; which of them is the box's `low` is decided by their distance from action_space.low, :61-69) and return the
 * sample closest to the first of them. */
static void genetic_expand(orc_tree* t, int sidx, double out[2]) {
    snode* s = &t->S[sidx];
    const int D = orc_action_dim(t->cfg.env_id);
    double lo[2], hi[2];
    const int f32 = box_bounds(t, lo, hi);
    out[0] = out[1] = 0.0;
    if (s->n_act == 0) {     /* np.median([high[i], low[i]]) of two values = their mean, in the box's dtype */
        for (int d = 0; d < D; ++d) out[d] = f32 ? (double)(((float)hi[d] + (float)lo[d]) / 2.0f) : (hi[d] + lo[d]) / 2.0;
        return;
    }
    if (s->n_act == 1) {
        double w[2] = {1.0, 1.0};
        const int pick_high = rnd_choices(t, w, 2) == 0;
        for (int d = 0; d < D; ++d) out[d] = pick_high ? hi[d] : lo[d];
        return;
    }
    if (s->n_act == 2) {
        const int has_high = find_action_bits(t, sidx, hi) >= 0;
        for (int d = 0; d < D; ++d) out[d] = has_high ? lo[d] : hi[d];
        return;
    }
    const double p = rnd_u01(t);
    if (!(p < t->cfg.epsilon)) { rnd_action(t, &t->tree_stream, out); return; }
    int b0 = -1, b1 = -1; double q0 = INFINITY, q1 = INFINITY;
    for (int a = s->first_act; a >= 0; a = t->A[a].next_act) {
        double q = t->A[a].total / (double)t->A[a].na;
        if (q < q0) { b1 = b0; q1 = q0; b0 = a; q0 = q; }
        else if (q < q1) { b1 = a; q1 = q; }
    }
    const double* best = t->A[b0].value; const double* second = t->A[b1].value;
    /* scipy distance.euclidean(action_space.low, x) = norm(low - x): in the arrays' dtype */
    double db, ds;
    if (f32) {
        const float e0 = (float)lo[0] - (float)best[0], e1 = (float)lo[0] - (float)second[0];
        db = (double)fabsf(e0); ds = (double)fabsf(e1);
    } else { db = euclid(lo, best, D); ds = euclid(lo, second, D); }
    const double* bl; const double* bh;
    if (db >= ds) { bh = best; bl = second; } else { bh = second; bl = best; }
    int arg = -1; double dmin = INFINITY, pick[2] = {0.0, 0.0};
    for (int k = 0; k < t->cfg.voo_n_samples; ++k) {        /* np.random.uniform(low, high, [n_samples, D]) */
        double x[2] = {0.0, 0.0};
        for (int d = 0; d < D; ++d) x[d] = rnd_uniform(t, bl[d], bh[d]);
        const double dist = euclid(best, x, D);             /* cdist([best_action], samples) in float64 */
        if (dist < dmin) { dmin = dist; arg = k; pick[0] = x[0]; pick[1] = x[1]; }   /* np.argmin: first minimum */
    }
    (void)arg;
    out[0] = pick[0]; out[1] = pick[1];
}

static double state_visit(orc_tree* t, int sidx, int level);

/* build_tree_action.  `level` = tree depth of the parent state node (root = 0). */
static double action_visit(orc_tree* t, int aidx, int level, int* descended_without_backup) {
    const orc_config* c = &t->cfg;
    double obs[ORC_MAX_S] = {0}; double r; int term;
    env_step(t, t->A[aidx].value, obs, &r, &term);
    t->c_expand_steps++;
    *descended_without_backup = 0;

    if (c->agent == ORC_AGENT_VANILLA || c->agent == ORC_AGENT_APW) {
        /* mcts.py:85-120 = mcts_hash.py:96-131 = mcts_action_...:99-136 */
        int child = find_child(t, aidx, obs);
        if (term) {
            if (child < 0) { child = new_snode(t, obs); t->S[child].terminal = 1; append_child(t, aidx, child); }
            t->A[aidx].total += r; t->A[aidx].na += 1; t->S[child].ns += 1;
            return r;
        }
        if (child < 0) {
            child = new_snode(t, obs); append_child(t, aidx, child);
            double delayed = c->gamma * do_rollout(t, budget_at(t, level + 1));
            double g = r + delayed;
            t->A[aidx].na += 1; t->S[child].ns += 1;
            t->A[aidx].total += g; t->S[child].total += g;
            return g;
        }
        double delayed = c->gamma * state_visit(t, child, level + 1);
        double g = r + delayed;
        t->A[aidx].total += g; t->A[aidx].na += 1;
        return g;
    }

    const double kk = (c->agent == ORC_AGENT_SPW) ? c->k1 : c->k2;
    const double al = (c->agent == ORC_AGENT_SPW) ? c->alpha1 : c->alpha2;
    anode* A = &t->A[aidx];
    if (c->agent == ORC_AGENT_SPW && term) {
        /* mcts_state_...:81-95: always a FRESH terminal child, overwriting the same key */
        int old = find_child(t, aidx, obs);
        int child = new_snode(t, obs);
        t->S[child].terminal = 1; t->S[child].terminal_reward = r;
        if (old >= 0) replace_child(t, aidx, old, child); else append_child(t, aidx, child);
        A = &t->A[aidx];
        A->total += r; A->na += 1; t->S[child].ns += 1;
        return r;
    }
    int widen = (A->n_child == 0) || ((double)A->n_child <= kk * pow((double)A->na, al));
    if (widen) {
        /* mcts_state_...:97-113 ; mcts_double_...:87-112 : the new node REPLACES an equal key */
        int old = find_child(t, aidx, obs);
        int child = new_snode(t, obs);
        if (old >= 0) replace_child(t, aidx, old, child); else append_child(t, aidx, child);
        A = &t->A[aidx];
        if (c->agent == ORC_AGENT_DPW && term) {
            t->S[child].terminal = 1;
            A->total += r; A->na += 1; t->S[child].ns += 1;
            return r;
        }
        double delayed = c->gamma * do_rollout(t, budget_at(t, level + 1));
        double g = r + delayed;
        A->na += 1; t->S[child].ns += 1; A->total += g; t->S[child].total += g;
        return g;
    }
    /* resample an existing child with weights na / ns_c (mcts_state_...:114-130, mcts_double_...:113-135) */
    int n = 0; int* idx = (int*)malloc(sizeof(int) * A->n_child); double* w = (double*)malloc(sizeof(double) * A->n_child);
    for (int ch = A->first_child; ch >= 0; ch = t->S[ch].next_sibling) {
        if (c->agent == ORC_AGENT_DPW && t->S[ch].terminal) continue;
        idx[n] = ch; w[n] = (double)A->na / (double)t->S[ch].ns; n++;
    }
    int pick = idx[rnd_choices(t, w, n)];
    free(idx); free(w);
    if (c->agent == ORC_AGENT_SPW && t->S[pick].terminal) {
        double tr = t->S[pick].terminal_reward;
        A->total += tr; A->na += 1; t->S[pick].ns += 1;
        return tr;
    }
    env_set_from_obs(t, t->S[pick].obs);
    double delayed = c->gamma * state_visit(t, pick, level + 1);
    *descended_without_backup = 1; /* self.total / self.na untouched (:126-130 / :131-135) */
    return r + delayed;
}

/* build_tree_state.  `level` = tree depth of this node. */
static double state_visit(orc_tree* t, int sidx, int level) {
    const orc_config* c = &t->cfg;
    int aidx, disc = -1;
    t->c_levels++;
    /* instrumentation: would the device tree's block of this node directly follow its parent's? */
    if (t->S[sidx].blk == 0) t->S[sidx].blk = ++t->next_blk;
    if (level > 0 && t->S[sidx].blk == t->cur_parent_blk + 1) t->c_seq++;
    t->cur_parent_blk = t->S[sidx].blk;
    if (t->cur_len < 4096) t->cur_path[t->cur_len++] = sidx;
    if (c->agent == ORC_AGENT_VANILLA || c->agent == ORC_AGENT_SPW) {
        /* mcts.py:55-62 */
        /* visit_actions[a] == 0  <=>  action a has no node yet (the count is bumped right after the node's first visit) */
        unsigned char seen[ORC_MAX_DISCRETE] = {0};
        for (int a = t->S[sidx].first_act; a >= 0; a = t->A[a].next_act) seen[(int)t->A[a].value[0]] = 1;
        int cand[ORC_MAX_DISCRETE], nc = 0;
        for (int a = 0; a < c->n_actions; ++a) if (!seen[a]) cand[nc++] = a;
        if (nc > 0) {
            disc = cand[rnd_choice(t, nc)];
            { double dv[2] = {(double)disc, 0.0}; aidx = new_anode(t, dv); }
            /* self.actions[action] = child: replaces an entry with the same key (cannot exist
             * in vanilla; SPW keeps visit_actions in step as well) */
            int old = -1;
            for (int a = t->S[sidx].first_act; a >= 0; a = t->A[a].next_act) if ((int)t->A[a].value[0] == disc) old = a;
            if (old >= 0) replace_action(t, sidx, old, aidx); else append_action(t, sidx, aidx);
        } else {
            aidx = act_by_rank(t, sidx, ucb_pick(t, sidx));
            disc = (int)t->A[aidx].value[0];
        }
    } else {
        /* mcts_action_...:62-75 ; mcts_double_...:56-65 */
        const snode* s = &t->S[sidx];
        double cap = ceil(c->k1 * pow((double)s->ns, c->alpha1));
        if (s->n_act == 0 || (double)s->n_act <= cap) {
            double a[2] = {0.0, 0.0};
            if (c->agent == ORC_AGENT_APW && c->expansion == ORC_EXPAND_VOO) voo_expand(t, sidx, a);
            else if (c->agent == ORC_AGENT_APW && c->expansion == ORC_EXPAND_GENETIC2) genetic2_expand(t, sidx, a);
            else if (c->agent == ORC_AGENT_APW && c->expansion == ORC_EXPAND_GENETIC) genetic_expand(t, sidx, a);
            else rnd_action(t, &t->tree_stream, a); /* dpw ignores --ae: always action_space.sample() */
            int old = find_action_bits(t, sidx, a);
            if (c->agent == ORC_AGENT_APW) {
                if (old >= 0) aidx = old;
                else { aidx = new_anode(t, a); append_action(t, sidx, aidx); }
            } else {
                aidx = new_anode(t, a); /* always a new node; visit_actions[key] = 0 */
                if (old >= 0) replace_action(t, sidx, old, aidx); else append_action(t, sidx, aidx);
            }
        } else {
            aidx = act_by_rank(t, sidx, ucb_pick(t, sidx));
        }
        if (c->agent == ORC_AGENT_DPW) env_set_from_obs(t, t->S[sidx].obs); /* mcts_double_...:68-69 */
    }
    int skip = 0;
    double reward = action_visit(t, aidx, level, &skip);
    snode* s = &t->S[sidx];
    s->ns += 1;
    t->A[aidx].visits += 1;   /* visit_actions[action] += 1 */
    s->total += reward;
    (void)skip;
    return reward;
}

/* ------------------------------------------------------------------ public API */
orc_tree* orc_tree_create(const orc_config* cfg, const double* root_state) {
    orc_tree* t = (orc_tree*)calloc(1, sizeof *t);
    t->cfg = *cfg;
    int d = orc_state_dim(cfg->env_id);
    for (int i = 0; i < d; ++i) t->root_state[i] = root_state[i];
    double obs[ORC_MAX_S] = {0}; make_obs(cfg->env_id, t->root_state, obs);
    new_snode(t, obs);
    stream_init(&t->tree_stream, cfg->seed, 0, (uint32_t)cfg->tree_id, STREAM_TREE);
    t->tlen = -1;
    return t;
}
int orc_tree_set_noise(orc_tree* t, const double* u, int n) {
    free(t->noise); t->noise = NULL; t->noise_n = 0;
    if (!u) return 0;
    if (t->cfg.env_id != ORC_ENV_FROZENLAKE || n < 1) return -1;
    t->noise = (double*)malloc(sizeof(double) * n);
    memcpy(t->noise, u, sizeof(double) * n);
    t->noise_n = n;
    return 0;
}

int orc_tree_set_heuristic(orc_tree* t, const double* table, int n_states, int n_actions) {
    free(t->heur); t->heur = NULL;
    if (!table) return 0;
    if (t->cfg.env_id != ORC_ENV_FROZENLAKE || n_states != 16 || n_actions != 4) return -1;
    t->heur = (double*)malloc(sizeof(double) * n_states * n_actions);
    memcpy(t->heur, table, sizeof(double) * n_states * n_actions);
    t->heur_states = n_states; t->heur_actions = n_actions;
    return 0;
}

void orc_tree_destroy(orc_tree* t) { if (t) { free(t->S); free(t->A); free(t->heur); free(t->noise); free(t); } }

int orc_tree_run(orc_tree* t, int n_sim, const uint8_t* kinds, const double* values, int64_t trace_len) {
    t->tk = kinds; t->tv = values; t->tlen = trace_len; t->tpos = 0; t->fault = 0;
    for (int s = 0; s < n_sim && !t->fault; ++s) {
        /* fit(): param.env = my_deepcopy(initial_env, ...) -- mcts_hash.py:31 ; spw/dpw poke root_data */
        memcpy(t->env, t->root_state, sizeof t->env);
        /* Mcts / MctsHash / APW step a pickled CLONE of the env in every simulation (utils/mcts_utils.py:8-16): the clone's
         * generator restarts, so the k-th env.step of every simulation draws noise[k].  SPW / DPW step the LIVE env
         * (they only poke the state: mcts_state_...:24-25, mcts_double_...:25-26): its generator keeps running, so the
         * noise index runs on through the whole fit() -- an action node then really gets several children. */
        if (!(t->cfg.agent == ORC_AGENT_SPW || t->cfg.agent == ORC_AGENT_DPW)) t->sim_step = 0;
        if (t->cfg.agent == ORC_AGENT_SPW || t->cfg.agent == ORC_AGENT_DPW) env_set_from_obs(t, t->S[0].obs);
        t->cur_len = 0;
        state_visit(t, 0, 0);
        { int k = 0; while (k < t->cur_len && k < t->prev_len && t->cur_path[k] == t->prev_path[k]) ++k;
          t->c_common += k; t->c_prevlen += t->prev_len;
          memcpy(t->prev_path, t->cur_path, sizeof(int) * t->cur_len); t->prev_len = t->cur_len; }
        t->sim_index++; t->c_sims++;
    }
    return t->fault;
}
int64_t orc_tree_trace_pos(const orc_tree* t) { return t->tpos; }

/* byte-lexicographic order of the float32 little-endian keys (python bytes comparison) */
static int key_less_bytes(double x, double y) {
    float fx = (float)x, fy = (float)y; return memcmp(&fx, &fy, 4) < 0;
}

int orc_tree_best(orc_tree* t, int32_t* child_rank, double* action /* [2] */) {
    /* mcts_hash.py:42-50 (insertion order).  spw/dpw first REORDER root.actions by key
     * (mcts_state_...:29, mcts_double_...:30: OrderedDict(sorted(items))), which persists. */
    snode* r = &t->S[0];
    int n = r->n_act; if (n == 0) return -1;
    int* order = (int*)malloc(sizeof(int) * n);
    int i = 0;
    for (int a = r->first_act; a >= 0; a = t->A[a].next_act) order[i++] = a;
    if (t->cfg.agent == ORC_AGENT_SPW || t->cfg.agent == ORC_AGENT_DPW) {
        for (int x = 1; x < n; ++x) {               /* stable insertion sort, like sorted() */
            int cur = order[x], y = x;
            while (y > 0) {
                double kv = t->A[cur].value[0], pv = t->A[order[y - 1]].value[0];
                int less = (t->cfg.agent == ORC_AGENT_SPW) ? (kv < pv) : key_less_bytes(kv, pv);
                if (!less) break;
                order[y] = order[y - 1]; --y;
            }
            order[y] = cur;
        }
        r->first_act = order[0]; r->last_act = order[n - 1];
        for (i = 0; i < n; ++i) t->A[order[i]].next_act = (i + 1 < n) ? order[i + 1] : -1;
    }
    double qmax = -INFINITY;
    for (i = 0; i < n; ++i) { double q = t->A[order[i]].total / (double)t->A[order[i]].na; if (q > qmax) qmax = q; }
    int ties = 0;
    for (i = 0; i < n; ++i) ties += (t->A[order[i]].total / (double)t->A[order[i]].na == qmax);
    int pos = rnd_choice(t, ties);
    for (i = 0; i < n; ++i) if (t->A[order[i]].total / (double)t->A[order[i]].na == qmax) { if (pos-- == 0) break; }
    if (i >= n) i = n - 1;
    *child_rank = i; action[0] = t->A[order[i]].value[0]; action[1] = t->A[order[i]].value[1];
    free(order);
    return t->fault;
}

int orc_tree_root_children(const orc_tree* t, int cap, double* action, int64_t* na, double* total, double* action2) {
    int i = 0;
    for (int a = t->S[0].first_act; a >= 0 && i < cap; a = t->A[a].next_act, ++i) {
        action[i] = t->A[a].value[0]; if (action2) action2[i] = t->A[a].value[1]; na[i] = t->A[a].na; total[i] = t->A[a].total;
    }
    return t->S[0].n_act;
}

/* agents/abstract_mcts.py:120-142 (state node) and :190-209 (action node), iterative: the subtree below root action
 * node `a0` is walked with an explicit stack of (state index, level); the direct child state has level 1.
 *   get_depth_max(0)        = deepest level reached (0 when the action node has no child)
 *   get_depth_mean(0, True) = mean level of the leaves: state nodes without actions, or whose actions all lack children */
static void action_depth_stats(const orc_tree* t, int a0, int32_t* dmax, double* dmean) {
    int cap = 256, sp = 0;
    int* st = (int*)malloc(sizeof(int) * 2 * cap);
    for (int c = t->A[a0].first_child; c >= 0; c = t->S[c].next_sibling) { st[2 * sp] = c; st[2 * sp + 1] = 1; ++sp; }
    int best = 0; double sum = 0; int64_t leaves = 0;
    while (sp > 0) {
        --sp; const int s = st[2 * sp], lvl = st[2 * sp + 1];
        if (lvl > best) best = lvl;
        int pushed = 0;
        for (int a = t->S[s].first_act; a >= 0; a = t->A[a].next_act)
            for (int c = t->A[a].first_child; c >= 0; c = t->S[c].next_sibling) {
                if (sp + 1 >= cap) { cap *= 2; st = (int*)realloc(st, sizeof(int) * 2 * cap); }
                st[2 * sp] = c; st[2 * sp + 1] = lvl + 1; ++sp; ++pushed;
            }
        if (!pushed) { sum += lvl; ++leaves; }
    }
    free(st);
    *dmax = best;
    *dmean = leaves ? sum / (double)leaves : 0.0;
}
int orc_tree_depth_stats(const orc_tree* t, int cap, int32_t* depth_max, double* depth_mean) {
    int i = 0;
    for (int a = t->S[0].first_act; a >= 0 && i < cap; a = t->A[a].next_act, ++i) action_depth_stats(t, a, depth_max + i, depth_mean + i);
    return t->S[0].n_act;
}

int64_t orc_tree_serialise(const orc_tree* t, int64_t cap, uint8_t* kind, int32_t* depth, int64_t* count,
                           double* total, uint8_t* terminal, int32_t* fanout, double* action, double* action2) {
    /* explicit stack of (kind, index, depth); children pushed in reverse insertion order */
    int64_t n = 0, sp = 0, scap = 1024;
    int* st = (int*)malloc(sizeof(int) * 3 * scap);
    st[0] = 0; st[1] = 0; st[2] = 0; sp = 1;
    int* tmp = (int*)malloc(sizeof(int) * (t->nS + t->nA + 1));
    while (sp > 0) {
        --sp; int k = st[3 * sp], idx = st[3 * sp + 1], d = st[3 * sp + 2];
        if (n < cap) {
            kind[n] = (uint8_t)k; depth[n] = d;
            if (k == 0) { count[n] = t->S[idx].ns; total[n] = t->S[idx].total; terminal[n] = (uint8_t)t->S[idx].terminal;
                          fanout[n] = t->S[idx].n_act; action[n] = 0.0; if (action2) action2[n] = 0.0; }
            else { count[n] = t->A[idx].na; total[n] = t->A[idx].total; terminal[n] = 0; fanout[n] = t->A[idx].n_child;
                   action[n] = t->A[idx].value[0]; if (action2) action2[n] = t->A[idx].value[1]; }
        }
        ++n;
        int m = 0;
        if (k == 0) for (int a = t->S[idx].first_act; a >= 0; a = t->A[a].next_act) tmp[m++] = a;
        else for (int c = t->A[idx].first_child; c >= 0; c = t->S[c].next_sibling) tmp[m++] = c;
        if (sp + m >= scap) { while (sp + m >= scap) scap *= 2; st = (int*)realloc(st, sizeof(int) * 3 * scap); }
        for (int j = m - 1; j >= 0; --j) { st[3 * sp] = 1 - k; st[3 * sp + 1] = tmp[j]; st[3 * sp + 2] = (k == 0) ? d : d + 1; ++sp; }
    }
    free(st); free(tmp);
    return n;
}

void orc_tree_counters(const orc_tree* t, int64_t out[8]) {
    out[0] = t->c_env_steps; out[1] = t->c_roll_steps; out[2] = t->c_levels; out[3] = t->c_sims;
    out[4] = t->nS; out[5] = t->nA; out[6] = t->c_expand_steps; out[7] = t->c_seq;
    if (getenv("ORC_PATH_STATS")) { out[6] = t->c_common; out[7] = t->c_prevlen; }
}

int orc_rollout(int env_id, const double* state, int budget, double gamma, uint64_t seed, uint64_t tree,
                uint32_t sim, uint32_t j, double* ret) {
    return orc_rollout_grid(env_id, state, budget, gamma, seed, tree, sim, j, 0, 0, ret);
}
int orc_rollout_grid(int env_id, const double* state, int budget, double gamma, uint64_t seed, uint64_t tree,
                     uint32_t sim, uint32_t j, int grid0, int grid1, double* ret) {
    orc_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.grid0 = grid0; cfg.grid1 = grid1;
    cfg.env_id = env_id; cfg.gamma = gamma; cfg.seed = seed; cfg.tree_id = tree; cfg.rollouts_per_leaf = 1;
    orc_tree t; memset(&t, 0, sizeof t); t.cfg = cfg; t.tlen = -1;
    int d = orc_state_dim(env_id); for (int i = 0; i < d; ++i) t.env[i] = state[i];
    stream_init(&t.roll_stream, seed, sim, (uint32_t)tree, j);
    int done = 0, depth = 0; double reward = 0;
    while (!done && depth < budget) {
        double a[2]; rnd_action(&t, &t.roll_stream, a);
        double obs[ORC_MAX_S]; double r; int term;
        env_step(&t, a, obs, &r, &term);
        reward += r * pow(gamma, depth);
        done = term; depth++;
    }
    *ret = reward;
    return depth;
}

/* ------------------------------------------------------------------ many trees, host threads */
typedef struct {
    const orc_config* cfg; const double* roots; int64_t lo, hi; int n_sim;
    int64_t* na; double* total; int64_t counters[8];
} many_job;

static void* many_worker(void* p) {
    many_job* j = (many_job*)p;
    int d = orc_state_dim(j->cfg->env_id), A = j->cfg->n_actions;
    memset(j->counters, 0, sizeof j->counters);
    for (int64_t i = j->lo; i < j->hi; ++i) {
        orc_config c = *j->cfg; c.tree_id = j->cfg->tree_id + (uint64_t)i;
        orc_tree* t = orc_tree_create(&c, j->roots + i * d);
        orc_tree_run(t, j->n_sim, NULL, NULL, -1);
        for (int a = 0; a < A; ++a) { j->na[i * A + a] = 0; j->total[i * A + a] = 0; }
        for (int a = t->S[0].first_act; a >= 0; a = t->A[a].next_act) {
            int k = (int)t->A[a].value[0]; j->na[i * A + k] = t->A[a].na; j->total[i * A + k] = t->A[a].total;
        }
        int64_t cc[8]; orc_tree_counters(t, cc);
        for (int k = 0; k < 8; ++k) j->counters[k] += cc[k];
        orc_tree_destroy(t);
    }
    return NULL;
}

int orc_run_many(const orc_config* cfg, const double* roots, int64_t n_trees, int n_sim, int n_threads,
                 int64_t* na_out, double* total_out, int64_t* counters_out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_trees) n_threads = (int)(n_trees > 0 ? n_trees : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    many_job* jobs = (many_job*)malloc(sizeof(many_job) * n_threads);
    for (int i = 0; i < n_threads; ++i) {
        jobs[i].cfg = cfg; jobs[i].roots = roots; jobs[i].n_sim = n_sim;
        jobs[i].lo = n_trees * i / n_threads; jobs[i].hi = n_trees * (i + 1) / n_threads;
        jobs[i].na = na_out; jobs[i].total = total_out;
        pthread_create(&th[i], NULL, many_worker, &jobs[i]);
    }
    for (int k = 0; k < 8; ++k) counters_out[k] = 0;
    for (int i = 0; i < n_threads; ++i) {
        pthread_join(th[i], NULL);
        for (int k = 0; k < 8; ++k) counters_out[k] += jobs[i].counters[k];
    }
    free(th); free(jobs);
    return 0;
}
