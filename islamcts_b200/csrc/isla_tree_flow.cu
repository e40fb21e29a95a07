// Thread-per-tree vanilla MCTS, WORKER / OWNER variant of the search kernel for batches too small to fill the GPU with
// 28 trees per warp (the per-GPU shards of the strong-scaling C5 run: 8192 and 16 384 CartPole trees).
//
// Same algorithm, data layout and results as tpt_search_kernel (isla_tree_tpt.cu: node blocks, PV cache chunks, lazy
// backup through the return table, cooperative verification, exact float64 decisions) -- what changes is who runs what:
//   * tpt_search_kernel: a warp owns tpw trees and alternates between the cooperative phase (32 lanes on one tree's levels)
//     and the per-tree phases (one lane per tree, lock-step: the warp waits for its slowest descent and rollout).  With 4
//     trees per warp the per-tree phases, which cost the same for 1 or 32 busy lanes, are 40 % of a simulation's time.
//   * here a CTA = one OWNER warp + up to 7 WORKER warps.  Lane t of the owner warp owns tree t of the CTA (up to 32) and
//     runs its per-tree phases as a state machine -- wait / descend one level / one env step of the rollout -- so every
//     tree advances at its own pace and a lane that finishes a simulation starts the next one at once.  When a tree's
//     cached path is worth verifying the owner lane posts {path length, simulations done, return-table index} in shared
//     memory; the worker warp the tree is assigned to runs the cooperative phase on it (same cp.async ring, prefetching
//     across its ready trees) and posts the hand-over level back.  No barrier: hand-offs are sequence-numbered mailboxes.
// Reference lines as in isla_tree_tpt.cu: selection mcts_hash.py:59-83, ucb1 action_selection_functions.py:9-25,
// expansion mcts_hash.py:87-131, rollout abstract_mcts.py:72-91, backup mcts_hash.py:79-81,119-131.
#include "isla_tree_tpt.cuh"

namespace isla {

namespace {

constexpr int kFlowMaxWarps = 8;
constexpr int kSeqFinished = 0x7fffffff;
enum : int { F_START = 0, F_WAIT = 1, F_DESCEND = 2, F_ROLL = 3, F_FIN = 4 };

__device__ __forceinline__ int ld_volatile_s32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile_s32(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }

template <int ENV, typename T>
__global__ void __launch_bounds__(32 * kFlowMaxWarps, 2) tpt_flow_kernel(const TptParams p) {
    constexpr int A = EnvTraits<ENV>::A;
    constexpr int S = EnvTraits<ENV>::S;
    static_assert(A == 2, "worker / owner kernel: two-action envs (CartPole)");
    constexpr double kInterior = Rewards<ENV>::kInterior;
    using PC = PvChunk<A>;
    constexpr int kSlots = CoopRing<A>::kSlots, kAhead = CoopRing<A>::kAhead, kSlotBytes = CoopRing<A>::kSlotBytes;
    static_assert(PC::kBytes == CoopRing<A>::kChunkBytes && !ISLA_BULK, "ring slot = PV chunk + pending returns (cp.async producer)");

    // dynamic shared memory: [ln(n) as float] | mailboxes xin[32] (owner -> worker), xout[32] (worker -> owner) | worker rings
    extern __shared__ float lnf_smem[];
    const int ln_slots = p.ln_in_smem ? ((p.ln_count + 3) & ~3) : 0;
    int4* const xin = reinterpret_cast<int4*>(reinterpret_cast<char*>(lnf_smem) + ln_slots * 4);
    int* const xout = reinterpret_cast<int*>(xin + 32);
    char* const rings = reinterpret_cast<char*>(xout + 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_workers = (blockDim.x >> 5) - 1;
    if (p.ln_in_smem)
        for (int i = threadIdx.x; i < p.ln_count; i += blockDim.x) lnf_smem[i] = (float)p.ln_table[i];
    if (threadIdx.x < 32) { xin[threadIdx.x] = make_int4(0, 0, 0, 0); xout[threadIdx.x] = 0; }
    __syncthreads();
    const double* const ln_table = p.ln_table;
    const double gamma = p.gamma, Cexp = p.C;
    const float Cf = (float)p.C;
    const long long cta_tree0 = (long long)blockIdx.x * p.tpc;
    const int n_cta = (int)min((long long)p.tpc, p.n_trees - cta_tree0);   // trees of this CTA
    const int coop_min = (p.tuning & 2) ? 1 : kCoopMin;

    if (warp > 0) {
        // ============================================================================================ WORKER warp
        // verifies the CTA's trees t = (warp - 1) + i * n_workers; lane i keeps the mailbox sequence number of its i-th tree
        const int w0 = warp - 1;
        const int my_t = w0 + lane * n_workers;
        bool mine = my_t < n_cta;
        int expect = 1;
        char* const ring = rings + w0 * (kSlots * kSlotBytes);
        const unsigned ring_dst = (unsigned)__cvta_generic_to_shared(ring) + lane * 16;
        const unsigned ring_end = ring_dst + kSlots * kSlotBytes;
        const unsigned g_adj = PC::kBytes - lane * 8;
        const long long t_stride = (long long)n_workers * p.pv_stride;       // between two trees of this warp
        char* const pv_w = p.pv + (cta_tree0 + w0) * p.pv_stride;
        int plv[PC::kCopies];
#pragma unroll
        for (int q = 0; q < PC::kCopies; ++q) plv[q] = lane + 32 * q < PC::kPieces ? PC::piece_level(lane + 32 * q) : (1 << 30);
        unsigned idle_ns = 64;
        while (__any_sync(kFull, mine)) {
            // ---- which of my trees are waiting for their cooperative phase?
            int my_plen = 0, a_sims = 0, my_ridx = 0;
            if (mine) {
                const int sq = ld_volatile_s32(&xin[my_t].w);
                if (sq == kSeqFinished) mine = false;
                else if (sq == expect) {
                    __threadfence_block();
                    my_plen = ld_volatile_s32(&xin[my_t].x); a_sims = ld_volatile_s32(&xin[my_t].y); my_ridx = ld_volatile_s32(&xin[my_t].z);
                }
            }
            const unsigned work = __ballot_sync(kFull, my_plen > 0);
            if (!work) { __nanosleep(idle_ns); idle_ns = min(idle_ns * 2, 1024u); continue; }
            idle_ns = 64;
            // ---- producer cursor over the ready trees (tree by tree in lane order), kAhead rounds in front of the consumer
            unsigned pwork = work;
            int rem = 0;
            const char* p_src = nullptr;
            const double* p_ret = nullptr;
            auto next_tree = [&]() {
                const int pj = __ffs(pwork) - 1;
                pwork &= pwork - 1;
                rem = __shfl_sync(kFull, my_plen, pj);
                p_src = pv_w + pj * t_stride + lane * 16;
                p_ret = p.ret_table + (__shfl_sync(kFull, my_ridx, pj) - lane);
            };
            unsigned pdst = ring_dst;
            auto issue_one = [&]() {
#pragma unroll
                for (int q = 0; q < PC::kCopies; ++q)
                    if (plv[q] < rem) cp_async16(pdst + q * 512, p_src + q * 512);
                if (lane < rem) cp_async8(pdst + g_adj, p_ret);
                rem -= 32; p_src += PC::kBytes; p_ret -= 32;
                cp_async_commit();
                pdst = pdst + kSlotBytes == ring_end ? ring_dst : pdst + kSlotBytes;
                if (rem <= 0 && pwork) next_tree();
            };
            next_tree();
#pragma unroll
            for (int i = 0; i < kAhead; ++i) issue_one();
            unsigned cslot = 0;
            for (unsigned todo = work; todo; todo &= todo - 1) {
                const int j = __ffs(todo) - 1;
                const int Lj = __shfl_sync(kFull, my_plen, j);
                const int sims_j = __shfl_sync(kFull, a_sims, j);
                char* wr_tot = pv_w + j * t_stride + PC::kTot + lane * 8;
                int event = -1;
                constexpr int NR = 2;
                for (int l0 = 0; l0 < Lj;) {
                    const bool decide = event < 0;
                    const int nr = min(NR, (Lj - l0 + 31) >> 5);
                    cp_async_wait<kAhead - NR>();
                    __syncwarp();
                    const char* slot[NR];
                    bool act[NR], ev[NR];
                    double tot_[NR];
                    int na_[NR];
                    {
                        unsigned cs = cslot;
                        const int lim = Lj - l0 - lane;
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            slot[r] = ring + cs;
                            cs = cs + kSlotBytes == kSlots * kSlotBytes ? 0u : cs + kSlotBytes;
                            act[r] = lim > 32 * r;
                            ev[r] = false;
                        }
                    }
                    int4 sv[NR];
                    double ov[NR], gv[NR];
#pragma unroll
                    for (int r = 0; r < NR; ++r) sv[r] = *reinterpret_cast<const int4*>(slot[r] + lane * 16);
#pragma unroll
                    for (int r = 0; r < NR; ++r) ov[r] = *reinterpret_cast<const double*>(slot[r] + PC::kTot + lane * 8);
#pragma unroll
                    for (int r = 0; r < NR; ++r) gv[r] = *reinterpret_cast<const double*>(slot[r] + PC::kBytes + lane * 8);
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        tot_[r] = __dadd_rn(ov[r], gv[r]);   // lazy backup: the return of the previous simulation
                        na_[r] = sv[r].y + sims_j;           // own_base + simulations done
                    }
                    if (decide) {
                        int sna[NR], ns[NR], ver[NR];
                        double st[NR];
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            sna[r] = sv[r].x;
                            st[r] = __hiloint2double(sv[r].w, sv[r].z);
                            ns[r] = na_[r] + sna[r] + ((r > 0 || l0 + lane > 0) ? 1 : 0);
                            ver[r] = 0;
                        }
                        {
                            float lg[NR], ro[NR], rs[NR], qo[NR], qs[NR], eo[NR], es[NR];
#pragma unroll
                            for (int r = 0; r < NR; ++r) {
                                const int i = act[r] ? ns[r] : 0;
                                lg[r] = p.ln_in_smem ? lnf_smem[i] : (float)__ldg(ln_table + i);
                            }
#pragma unroll
                            for (int r = 0; r < NR; ++r) { ro[r] = rcp_approx((float)na_[r]); rs[r] = rcp_approx((float)sna[r]); }
#pragma unroll
                            for (int r = 0; r < NR; ++r) { qo[r] = (float)tot_[r] * ro[r]; qs[r] = (float)st[r] * rs[r]; }
#pragma unroll
                            for (int r = 0; r < NR; ++r) { eo[r] = Cf * sqrt_approx(lg[r] * ro[r]); es[r] = Cf * sqrt_approx(lg[r] * rs[r]); }
#pragma unroll
                            for (int r = 0; r < NR; ++r) {
                                const float d = (qo[r] + eo[r]) - (qs[r] + es[r]);
                                const float t = 1e-5f * fmaxf(fabsf(qo[r]) + fabsf(eo[r]), fabsf(qs[r]) + fabsf(es[r])) + 1e-30f;
                                ver[r] = d > t ? 1 : (d < -t ? -1 : 0);
                            }
                        }
                        bool undecided = false;
#pragma unroll
                        for (int r = 0; r < NR; ++r) undecided |= act[r] && ver[r] == 0;
                        if (undecided) {
#pragma unroll
                            for (int r = 0; r < NR; ++r)
                                if (act[r] && ver[r] == 0) {
                                    const double lg = ns[r] < p.ln_count ? ln_table[ns[r]] : ::log((double)ns[r]);
                                    const double so = __dadd_rn(__ddiv_rn(tot_[r], (double)na_[r]), __dmul_rn(Cexp, __dsqrt_rn(__ddiv_rn(lg, (double)na_[r]))));
                                    const double ss = __dadd_rn(__ddiv_rn(st[r], (double)sna[r]), __dmul_rn(Cexp, __dsqrt_rn(__ddiv_rn(lg, (double)sna[r]))));
                                    ver[r] = so > ss ? 1 : -1;   // a tie needs a draw: the owner decides it
                                }
                        }
#pragma unroll
                        for (int r = 0; r < NR; ++r) ev[r] = act[r] && ver[r] < 0;
                        bool any = false;
#pragma unroll
                        for (int r = 0; r < NR; ++r) any |= ev[r];
                        if (__any_sync(kFull, any)) {
#pragma unroll
                            for (int r = NR - 1; r >= 0; --r) {
                                const unsigned e = __ballot_sync(kFull, ev[r]);
                                if (e) event = l0 + 32 * r + __ffs(e) - 1;
                            }
                        }
                    }
                    if (event < 0 && l0 + 32 * nr < Lj) {
#pragma unroll
                        for (int r = 0; r < NR; ++r) if (act[r]) __stcg(reinterpret_cast<double*>(wr_tot + r * PC::kBytes), tot_[r]);
                    } else {
#pragma unroll
                        for (int r = 0; r < NR; ++r) if (act[r]) {
                            const int l = l0 + 32 * r + lane;
                            const bool leaves = (event >= 0 && l >= event) || (event < 0 && l == Lj - 1);
                            if (!leaves) __stcg(reinterpret_cast<double*>(wr_tot + r * PC::kBytes), tot_[r]);
                            else {
                                const uint32_t e = *reinterpret_cast<const uint32_t*>(slot[r] + PC::kEdge + lane * 4);
                                char* const bj = p.blocks + (cta_tree0 + w0 + (long long)j * n_workers) * p.slots_per_tree * 16;
                                __stcg(total_ptr<A>(bj, (int)(e >> 2), (int)(e & 3u)), tot_[r]);
                                __stcg(na_ptr<A>(bj, (int)(e >> 2), (int)(e & 3u)), na_[r]);
                            }
                        }
                    }
                    wr_tot += nr * PC::kBytes;
                    cslot += (unsigned)nr * kSlotBytes;
                    if (cslot >= (unsigned)(kSlots * kSlotBytes)) cslot -= (unsigned)(kSlots * kSlotBytes);
                    l0 += 32 * nr;
                    __syncwarp();
                    issue_one();
                    if (nr > 1) issue_one();
                }
                // hand the tree back: every lane's global stores first, then the mailbox word (sequence << 16 | level)
                __threadfence_block();
                __syncwarp();
                if (lane == j) {
                    st_volatile_s32(&xout[my_t], ((expect & 0x7fff) << 16) | (event >= 0 ? event : Lj - 1));
                    ++expect;
                }
            }
            cp_async_wait<0>();
        }
        return;
    }

    // ================================================================================================ OWNER warp
    const long long tree = lane < n_cta ? cta_tree0 + lane : p.n_trees;
    const bool valid = tree < p.n_trees;
    const long long tix = valid ? tree : 0;
    char* const blocks = p.blocks + tix * p.slots_per_tree * 16;
    T* const states = reinterpret_cast<T*>(p.states) + tix * p.slots_per_tree * S;
    double* const gfirst = p.gfirst + tix * p.slots_per_tree;
    int32_t* const meta = p.meta + tix * META_WORDS;
    char* const pv = p.pv + tix * p.pv_stride;

    int alloc = valid ? meta[META_ALLOC] : 0;
    uint32_t sims_done = valid ? (uint32_t)meta[META_SIMS] : 0u;
    int fault = valid ? meta[META_FAULT] : 0;
    int plen = 0;
    bool pend = false;
    int ret_j = 0;
    const uint32_t tree_id = (uint32_t)(p.tree_id_base + (unsigned long long)tix);
    uint32_t rng_used = valid ? (uint32_t)meta[META_RNG] : 0u;
    unsigned int c_env = 0, c_roll = 0, c_levels = 0, c_nodes = 0, c_flushed = 0;
    int s_done = 0, seq = 0;

    auto choice = [&](int n) -> int {
        if (n <= 1) return 0;
        Stream ts;
        ts.resume(p.seed, 0u, tree_id, kStreamTree, rng_used);
        const uint32_t w = ts.next();
        ++rng_used;
        return (int)mulhi_u32(w, (uint32_t)n);
    };
    // UCB1 with fp32 screening and exact float64 near-ties (see isla_tree_tpt.cu)
    auto ucb = [&](const NodeBlock<A>& c, bool nonroot, int& a_sel, double (&sc)[A], double& best) -> int {
        int ns = nonroot ? 1 : 0;
#pragma unroll
        for (int a = 0; a < A; ++a) ns += c.na[a];
        {
            const float lgf = ns < p.ln_count ? (p.ln_in_smem ? lnf_smem[ns] : (float)ln_table[ns]) : (float)::log((double)ns);
            float sf[A], m1 = -INFINITY, scale = 0.f;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const float rn = rcp_approx((float)c.na[a]);
                const float qf = (float)c.total[a] * rn;
                const float ef = Cf * sqrt_approx(lgf * rn);
                sf[a] = qf + ef;
                m1 = fmaxf(m1, sf[a]);
                scale = fmaxf(scale, fabsf(qf) + fabsf(ef));
            }
            const float thr = m1 - 1e-5f * scale - 1e-30f;
            int ncand = 0;
#pragma unroll
            for (int a = 0; a < A; ++a) { if (sf[a] >= thr) { ++ncand; a_sel = a; } }
            if (ncand == 1) return 1;
        }
        const double lg = ns < p.ln_count ? ln_table[ns] : ::log((double)ns);
        best = -INFINITY;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const double na = (double)c.na[a];
            sc[a] = __dadd_rn(__ddiv_rn(c.total[a], na), __dmul_rn(Cexp, __dsqrt_rn(__ddiv_rn(lg, na))));
            best = fmax(best, sc[a]);
        }
        int ties = 0, first_rank = A;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            if (sc[a] == best) {
                ++ties;
                const int rk = (int)(c.link[a] >> kLinkRankShift);
                if (rk < first_rank) { first_rank = rk; a_sel = a; }
            }
        }
        return ties;
    };
    auto ret_value = [&](int j, int k) -> double {
        const double* row = p.ret_table + (long long)j * p.ret_K;
        if (k < p.ret_K) return __ldg(row + k);
        double g = __ldg(row + p.ret_K - 1);
        for (int i = p.ret_K - 1; i < k; ++i) g = __dadd_rn(kInterior, __dmul_rn(gamma, g));
        return g;
    };
    auto write_back = [&](int from) {
        const int top = plen;
        for (int l = from; l < plen; ++l) {
            const double t_ = __ldcg(PC::tot(pv, l));
            const uint32_t e = __ldcg(PC::edge(pv, l));
            const double g = pend ? ret_value(ret_j, top - l) : 0.0;
            __stcg(total_ptr<A>(blocks, (int)(e >> 2), (int)(e & 3u)), pend ? __dadd_rn(t_, g) : t_);
            __stcg(na_ptr<A>(blocks, (int)(e >> 2), (int)(e & 3u)), PC::own_base(pv, l) + (int)sims_done);
            ++c_flushed;
        }
        plen = from;
        if (from == 0) pend = false;
    };

    // per-lane state of the simulation in flight
    int mode = valid ? F_START : F_FIN;
    int sims_left = p.n_sim;
    int level = 0, node_slot = 0, blk = 1, exp_a = 0, exp_rank = 0;
    bool fresh = false;
    double G = 0.0;
    T st[S];
#pragma unroll
    for (int i = 0; i < S; ++i) st[i] = T(0);
    T r_exp = T(0), last_r = T(0);
    int t_roll = 0, budget = 0;
    Stream rs;
    rs.init(p.seed, 0u, tree_id, 0u);

    // a simulation ends: the new edge (if one was expanded) is written, the return stays pending (return table row ret_j)
    auto end_simulation = [&](bool expanded, bool rolling) {
        if (expanded) {
            uint32_t lk = (uint32_t)exp_rank << kLinkRankShift;
            if (!rolling) lk |= kLinkTerminal | (G == 1.0 ? kLinkRewardOne : 0u);
            __stcg(total_ptr<A>(blocks, blk, exp_a), G);
            __stcg(na_ptr<A>(blocks, blk, exp_a), 1);
            __stcg(link_ptr<A>(blocks, blk, exp_a), lk);
            __stcg(gfirst + (long long)blk * A + exp_a, rolling ? G : 0.0);
            ++c_nodes;
        }
        c_levels += (unsigned)level + 1u;
        pend = level > 0;
        ++sims_done; ++s_done; --sims_left;
        mode = F_START;
    };

    while (__any_sync(kFull, mode != F_FIN)) {
        if (mode == F_START) {
            if (sims_left <= 0 || fault) {
                if (plen > 0) write_back(0);
                __threadfence_block();
                st_volatile_s32(&xin[lane].w, kSeqFinished);
                mode = F_FIN;
            } else {
                if (pend && plen >= p.ret_K) write_back(0);   // deeper than the return table (rare)
                if (n_workers > 0 && pend && plen >= coop_min) {
                    // the cached path goes to this tree's worker warp
                    ++seq;
                    xin[lane].x = plen; xin[lane].y = (int)sims_done; xin[lane].z = ret_j * p.ret_K + plen;
                    __threadfence_block();
                    st_volatile_s32(&xin[lane].w, seq);
                    mode = F_WAIT;
                } else {
                    if (plen > 0) write_back(0);
                    level = 0; node_slot = 0; blk = 1; fresh = false;
                    mode = F_DESCEND;
                }
            }
        }
        if (mode == F_WAIT) {
            const int w = ld_volatile_s32(&xout[lane]);
            if ((w >> 16) == (seq & 0x7fff)) {
                __threadfence_block();
                level = w & 0xffff; plen = level; pend = false;
                node_slot = 0; blk = 1; fresh = false;
                if (level > 0) {
                    const uint32_t prev = __ldcg(PC::edge(pv, level - 1)), here = __ldcg(PC::edge(pv, level));
                    node_slot = (int)(prev >> 2) * A + (int)(prev & 3u);
                    blk = (int)(here >> 2);
                }
                mode = F_DESCEND;
            }
        }
        if (mode == F_DESCEND) {
            // one level of the serial continuation on the node blocks
            NodeBlock<A> c;
            if (fresh) {
#pragma unroll
                for (int a = 0; a < A; ++a) { c.total[a] = 0.0; c.na[a] = 0; c.link[a] = 0u; }
            } else {
                c = load_block<A>(blocks, blk);
            }
            int nun = 0;
#pragma unroll
            for (int a = 0; a < A; ++a) nun += (c.na[a] == 0);
            int a_sel = 0;
            if (nun > 0) {
                // uniformly random unvisited action (mcts_hash.py:70-74), expanded at once
                const int pos = choice(nun);
                int seen = 0;
#pragma unroll
                for (int a = 0; a < A; ++a) { if (c.na[a] == 0) { if (seen == pos) exp_a = a; ++seen; } }
                exp_rank = A - nun;
                if constexpr (S == 4 && sizeof(T) == 4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(states + (long long)node_slot * S));
                    st[0] = v.x; st[1] = v.y; st[2] = v.z; st[3] = v.w;
                } else {
#pragma unroll
                    for (int i = 0; i < S; ++i) st[i] = __ldcg(states + (long long)node_slot * S + i);
                }
                const bool term = Env<ENV, T>::step(st, T(exp_a), r_exp);
                ++c_env;
                if (term) {
                    G = (double)r_exp;   // terminal child: total += r, na += 1 (mcts_hash.py:96-107)
                    if (G != 0.0 && G != 1.0) fault = ISLA_ERR_INVALID;
                    ret_j = Rewards<ENV>::leaf_row_terminal(G != 0.0, p.max_depth);
                    end_simulation(true, false);
                } else {
                    if ((double)r_exp != kInterior) fault = ISLA_ERR_INVALID;
                    const int child = blk * A + exp_a;
                    if constexpr (S == 4 && sizeof(T) == 4) {
                        __stcg(reinterpret_cast<float4*>(states + (long long)child * S), make_float4(st[0], st[1], st[2], st[3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < S; ++i) __stcg(states + (long long)child * S + i, st[i]);
                    }
                    budget = p.budget_mode == ISLA_BUDGET_ZERO ? 0 : min(p.max_depth - (level + 1), p.disc_count);   // mcts_hash.py:116
                    t_roll = 0; last_r = T(0);
                    rs.init(p.seed, sims_done, tree_id, 0u);
                    if (budget > 0) mode = F_ROLL;
                    else {
                        G = __dadd_rn((double)r_exp, __dmul_rn(gamma, Rewards<ENV>::rollout_return(0, 0.0, p.disc_table, p.cum_table)));
                        ret_j = Rewards<ENV>::leaf_row_rollout(0, 0.0);
                        end_simulation(true, true);
                    }
                }
            } else {
                double sc[A], best;
                const int ties = ucb(c, node_slot != 0, a_sel, sc, best);
                if (ties > 1) {
                    int pos = choice(ties);
#pragma unroll
                    for (int rk = 0; rk < A; ++rk) {
#pragma unroll
                        for (int a = 0; a < A; ++a) {
                            if ((int)(c.link[a] >> kLinkRankShift) == rk && sc[a] == best) { if (pos == 0) a_sel = a; --pos; }
                        }
                    }
                }
                uint32_t lk = c.link[0];
#pragma unroll
                for (int a = 1; a < A; ++a) if (a == a_sel) lk = c.link[a];
                if (lk & kLinkTerminal) {
                    // terminal child re-visited: total += r, na += 1 (mcts_hash.py:96-107)
                    G = (lk & kLinkRewardOne) ? 1.0 : 0.0;
                    ret_j = Rewards<ENV>::leaf_row_terminal(G != 0.0, p.max_depth);
#pragma unroll
                    for (int a = 0; a < A; ++a) if (a == a_sel) {
                        __stcg(total_ptr<A>(blocks, blk, a), __dadd_rn(c.total[a], G));
                        __stcg(na_ptr<A>(blocks, blk, a), c.na[a] + 1);
                    }
                    end_simulation(false, false);
                } else {
                    // this level joins the PV cache: own statistics + the sibling's, as they are before the backup
                    PvSta<A> sb;
                    double ot = c.total[0]; int on_ = c.na[0];
#pragma unroll
                    for (int a = 0; a < A; ++a) {
                        if (a == a_sel) { ot = c.total[a]; on_ = c.na[a]; }
                        else { sb.sib_total[0] = c.total[a]; sb.sib_na[0] = c.na[a]; }
                    }
                    sb.own_base = on_ - (int)sims_done;
                    store_sta<A>(PC::sta(pv, level), sb);
                    __stcg(PC::tot(pv, level), ot);
                    __stcg(PC::edge(pv, level), ((uint32_t)blk << 2) | (uint32_t)a_sel);
                    ++level;
                    plen = level;
                    node_slot = blk * A + a_sel;
                    int nb = (int)(lk & kLinkBlockMask);
                    fresh = false;
                    if (nb == 0) {
                        if (alloc >= p.block_cap) { fault = ISLA_ERR_CAPACITY; mode = F_START; }
                        else { nb = alloc++; __stcg(link_ptr<A>(blocks, blk, a_sel), lk | (uint32_t)nb); fresh = true; }
                    }
                    blk = nb;
                }
            }
            if (fault && mode != F_FIN) mode = F_START;   // F_START retires a faulted tree
        } else if (mode == F_ROLL) {
            // one env step of the random rollout (abstract_mcts.py:72-91)
            const T act = sample_action<ENV, T>(rs);
            const bool done = Env<ENV, T>::step(st, act, last_r);
            ++t_roll;
            if (done || t_roll >= budget) {
                c_roll += t_roll; c_env += t_roll;
                const double R = Rewards<ENV>::rollout_return(t_roll, (double)last_r, p.disc_table, p.cum_table);
                G = __dadd_rn((double)r_exp, __dmul_rn(gamma, R));
                ret_j = Rewards<ENV>::leaf_row_rollout(t_roll, (double)last_r);
                end_simulation(true, true);
            }
        }
    }

    if (valid) {
        meta[META_ALLOC] = alloc;
        meta[META_SIMS] = (int32_t)sims_done;
        meta[META_RNG] = (int32_t)rng_used;
        meta[META_FAULT] = fault;
        meta[META_PENDING] = 0;
        atomicAdd(p.counters + CNT_ENV_STEPS, (unsigned long long)c_env);
        atomicAdd(p.counters + CNT_ROLLOUT_STEPS, (unsigned long long)c_roll);
        atomicAdd(p.counters + CNT_LEVELS, (unsigned long long)c_levels);
        atomicAdd(p.counters + CNT_SIMS, (unsigned long long)s_done);
        atomicAdd(p.counters + CNT_NODES, (unsigned long long)c_nodes);
        if (fault) atomicAdd(p.counters + CNT_FAULTS, 1ull);
        atomicAdd(p.counters + 6, (unsigned long long)c_flushed);
    }
}

template <typename T>
int launch_flow(TptParams& p, int sm_count, int tuning, cudaStream_t st) {
    constexpr int ENV = ENV_CARTPOLE;
    constexpr int A = EnvTraits<ENV>::A;
    constexpr int kMailBytes = 32 * 16 + 32 * 4;
    const int ln_fits = p.ln_count * 4 <= 8 * 1024;
    auto smem_for = [&](int workers) {
        return (ln_fits ? (size_t)((p.ln_count + 3) & ~3) * 4 : 0) + kMailBytes + (size_t)workers * CoopRing<A>::kWarpBytes;
    };
    static bool attr_set = false;
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(tpt_flow_kernel<ENV, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(CoopRing<A>::kWarpBytes * (kFlowMaxWarps - 1) + 8192 + kMailBytes + 4096));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tpt_flow_kernel)");
        attr_set = true;
    }
    // CTA shape: `workers` worker warps + the owner warp; all CTAs resident at once; as few trees per CTA as that allows.
    // tuning bits 8..13 pin the trees per CTA, bits 16..19 the worker warps (A/B timing, tests).
    int workers = (tuning >> 16) & 15;
    if (workers < 1 || workers > kFlowMaxWarps - 1) workers = kFlowMaxWarps - 1;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tpt_flow_kernel<ENV, T>, 32 * (workers + 1), smem_for(workers)) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    const long long resident = (long long)(sm_count > 0 ? sm_count : 148) * occ;
    const long long waves = (p.n_trees + 32 * resident - 1) / (32 * resident);
    long long tpc = (p.n_trees + waves * resident - 1) / (waves * resident);
    const int tpc_forced = (tuning >> 8) & 63;
    if (tpc_forced >= 1 && tpc_forced <= 32) tpc = tpc_forced;
    p.tpc = (int)(tpc < 1 ? 1 : (tpc > 32 ? 32 : tpc));
    p.ln_in_smem = ln_fits;
    const unsigned grid = (unsigned)((p.n_trees + p.tpc - 1) / p.tpc);
    tpt_flow_kernel<ENV, T><<<grid, 32 * (workers + 1), smem_for(workers), st>>>(p);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : cuda_fail(err, "tpt_flow_kernel launch");
}

}  // namespace

// Entry used by tpt_run (isla_tree_tpt.cu) for free-running CartPole searches.
int tpt_flow_launch(TptParams& p, bool f64, int sm_count, int tuning, cudaStream_t st) {
    return f64 ? launch_flow<double>(p, sm_count, tuning, st) : launch_flow<float>(p, sm_count, tuning, st);
}

}  // namespace isla
