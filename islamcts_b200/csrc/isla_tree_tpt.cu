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
//   * Device tree = node blocks in HBM (isla_internal.h): the A edges of one state node are ONE 32-byte sector
//     (A=2) / 64 bytes (A=4); an edge's link word carries the block of the child.  States and creation returns
//     are cold side arrays.
//   * PV cache + lazy backup.  Consecutive simulations share 99.7 % of their path (measured, CartPole C5).  The
//     statistics of the path's edges therefore live in per-tree chunks of 32 consecutive levels (per level: sibling
//     statistics 16 B, own total 8 B, pending return 8 B, edge id 4 B; the own visit count is implicit); the returns
//     G_l of simulation s are added by simulation s+1 when it re-evaluates the level, and the node blocks are only
//     touched where the path changes.  Every level sees exactly the statistics an eager backup would have produced,
//     so the decisions are those of the sequential algorithm.
//   * Cooperative verification.  A warp owns up to 32 trees; for each tree in turn its 32 lanes re-evaluate UCB1 on 32
//     consecutive LEVELS of the cached path and ballot whether the decisions stand -- a dependent pointer chase of
//     ~180 levels becomes ~6 coalesced rounds.  Only the tail (first changed decision, tie-break draws, expansion)
//     is walked serially on the node blocks, lock-step across the warp's trees.
//   * Visit counts and totals are exact: int32 counts, float64 totals, UCB decisions are the float64 ones (fp32
//     screening settles only unambiguous arg-maxes), ln(n) and gamma^t come from host libm tables.  Env state and
//     rollouts live in registers; Philox4x32-10 replaces numpy/gym RNGs.
// tuning bits: 2 = cooperative verification at any path depth (tests); -DISLA_NO_SCREEN builds without the fp32 screening (A/B timing),
// bits 8..13 = trees per warp (A/B timing).
#include <cmath>
#include <cstring>

#include "isla_tree_tpt.cuh"

namespace isla {

namespace {
// One simulation of the 32 trees of a warp = four phases.
//   1a  cooperative verification   for each tree in turn, the 32 lanes take 32 consecutive LEVELS of its cached
//                                   path: add the pending return to the level's own statistics, re-evaluate UCB1
//                                   against the cached sibling statistics, and vote whether the old decision still
//                                   stands (it does on 99.7 % of the levels).  All reads and writes are consecutive
//                                   16-byte entries of one tree's rows; node blocks are not touched.  The first level
//                                   that needs a draw or decides differently is handed to phase 1b; the levels from
//                                   there on leave the cache (their statistics are written back to the node blocks).
//   1b  serial continuation        every lane descends its own tree from the hand-over level on the node blocks
//                                   (lock-step) and caches the levels it passes
//   2   expansion + rollout        one env step per iteration for every lane still playing
//   3   return pass                G_l = r + gamma G_{l+1} written to the tree's pending-return row
// The decisions are those of the sequential algorithm: a level's UCB sees exactly the statistics an eager backup
// would have produced; draws happen only in 1b, in path order.
// COOP = false compiles the kernel WITHOUT the PV cache and the cooperative phase: every simulation descends from the root
// on the node blocks and backs its return up eagerly (the same additions in the same order, so the same trees).  Used for
// shallow trees that never reach the cooperative threshold (FrozenLake: paths of <= ~10 levels) and for scripted replays:
// no ring in shared memory, no per-level cache traffic, fewer registers.
template <int ENV, typename T, bool SCRIPTED, bool COOP>
__global__ void __launch_bounds__(kTptBlock, ISLA_TPT_MINBLOCKS) tpt_search_kernel(const TptParams p) {
    static_assert(!(SCRIPTED && COOP), "scripted replays descend from the root");
    constexpr int A = EnvTraits<ENV>::A;
    constexpr int S = EnvTraits<ENV>::S;
    constexpr double kInterior = Rewards<ENV>::kInterior;
    // dynamic shared memory: [ln(n) rounded to float, for the fp32 screening pass (when it fits)] | cooperative ring.
    // The float64 table (host libm values) stays in global memory: only near-ties read it.
    extern __shared__ float lnf_smem[];
    const int ln_slots = p.ln_in_smem ? ((p.ln_count + 3) & ~3) : 0;
    if (p.ln_in_smem) {
        for (int i = threadIdx.x; i < p.ln_count; i += blockDim.x) lnf_smem[i] = (float)p.ln_table[i];
        __syncthreads();
    }
    const double* const ln_table = p.ln_table;
    const int lane = threadIdx.x & 31;
    constexpr int kSlots = CoopRing<A>::kSlots, kAhead = CoopRing<A>::kAhead, kSlotBytes = CoopRing<A>::kSlotBytes;
    char* const coop_ring = reinterpret_cast<char*>(lnf_smem) + ln_slots * 4;
#if ISLA_BULK
    // mbarriers of this warp's ring slots (behind all rings); par_bits = phase parity the consumer waits for, per slot
    const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(coop_ring + (kTptBlock / 32) * (kSlots * kSlotBytes)) + (threadIdx.x >> 5) * (kSlots * 8);
    unsigned par_bits = 0u;
    if (lane == 0) {
        for (int i = 0; i < kSlots; ++i) mbar_init(mbar0 + i * 8, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#endif

    // scripted (parity) mode drives ONE tree from lane 0 of a single warp; the other lanes are dummies
    const long long warp_tree0 = ((long long)blockIdx.x * (kTptBlock / 32) + (threadIdx.x >> 5)) * p.tpw;  // tree of lane 0 (free-running)
    const long long tree = SCRIPTED ? (threadIdx.x == 0 ? p.scripted_tree : p.n_trees)
                                    : (lane < p.tpw ? warp_tree0 + lane : p.n_trees);
    const bool valid = tree < p.n_trees;
    const long long tix = valid ? tree : 0;
    constexpr long long kStaBytes = sizeof(PvSta<A>);
    using PC = PvChunk<A>;
    static_assert(PC::kBytes == CoopRing<A>::kChunkBytes, "a ring slot starts with a copy of a PV chunk");

    char* const blocks = p.blocks + tix * p.slots_per_tree * 16;
    T* const states = reinterpret_cast<T*>(p.states) + tix * p.slots_per_tree * S;
    double* const gfirst = p.gfirst + tix * p.slots_per_tree;
    int32_t* const meta = p.meta + tix * META_WORDS;
    char* const pv = p.pv + tix * p.pv_stride;  // this tree's PV chunks

    int alloc = valid ? meta[META_ALLOC] : 0;
    uint32_t sims_done = valid ? (uint32_t)meta[META_SIMS] : 0u;
    int fault = valid ? meta[META_FAULT] : 0;
    int plen = 0;        // cached levels [0, plen): pv_tot authoritative, node blocks stale for those edges
    bool pend = false;   // the returns G_l of those levels are still to be added:
    int ret_j = 0;       //   G_l = ret_table[ret_j][plen - l] (row = the leaf value of the last simulation)
    const uint32_t tree_id = (uint32_t)(p.tree_id_base + (unsigned long long)tix);
    // the tree stream is rarely drawn from (expansion picks, exact UCB ties): only its position is kept in a register and
    // the Philox state is rebuilt per draw
    uint32_t rng_used = valid ? (uint32_t)meta[META_RNG] : 0u;
    TraceCursor tc{p.trace_kinds, p.trace_values, p.trace_len, 0, 0};

    unsigned int c_env = 0, c_roll = 0, c_levels = 0, c_nodes = 0, c_flushed = 0;   // per thread and launch: 32 bits suffice
    const double gamma = p.gamma, Cexp = p.C;
    const float Cf = (float)p.C;

    auto choice = [&](int n) -> int {  // np.random.choice(candidates) -> position
        if constexpr (SCRIPTED) {
            const int v = (int)tc.take(ISLA_K_CHOICE);
            if (v < 0 || v >= n) { if (!tc.fault) tc.fault = ISLA_ERR_TRACE; return 0; }
            return v;
        } else {
            if (n <= 1) return 0;
            Stream ts;
            ts.resume(p.seed, 0u, tree_id, kStreamTree, rng_used);
            const uint32_t w = ts.next();
            ++rng_used;
            return (int)mulhi_u32(w, (uint32_t)n);
        }
    };
    // UCB1 over the A visited children (action_selection_functions.py:9-25).  The reference evaluates the scores
    // in float64 and breaks exact ties at random.  An fp32 screening pass (error < 1e-6 relative, tolerance 1e-5)
    // settles the arg-max whenever it is unambiguous; only near-ties take the exact float64 path, so the DECISIONS
    // are the float64 ones.  Returns the number of exactly tied maxima (1 = unique); a_sel = the first of them
    // in insertion order; sc/best are filled when ties > 1 is possible.
    auto ucb = [&](const NodeBlock<A>& c, bool nonroot, int& a_sel, double (&sc)[A], double& best) -> int {
        int ns = nonroot ? 1 : 0;
#pragma unroll
        for (int a = 0; a < A; ++a) ns += c.na[a];
#ifndef ISLA_NO_SCREEN
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
#endif
        const double lg = ns < p.ln_count ? ln_table[ns] : ::log((double)ns);
        best = -INFINITY;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const double na = (double)c.na[a];
            const double q = __ddiv_rn(c.total[a], na);
            const double ex = __dmul_rn(Cexp, __dsqrt_rn(__ddiv_rn(lg, na)));
            sc[a] = __dadd_rn(q, ex);
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
    // Return table (host libm arithmetic, isla_capi.cu): row j = a leaf value G_leaf, column k = the return k levels
    // above it, G = r + gamma * G' applied k times (mcts_hash.py:126-131, 79-81) with this kernel's constant interior
    // reward.  The backup of a simulation is therefore two integers per tree (leaf level, row); nothing is written per
    // level.  Columns beyond the table (paths deeper than ret_K levels) continue the recurrence from its last column.
    auto ret_value = [&](int j, int k) -> double {
        const double* row = p.ret_table + (long long)j * p.ret_K;
        if (k < p.ret_K) return __ldg(row + k);
        double g = __ldg(row + p.ret_K - 1);
        for (int i = p.ret_K - 1; i < k; ++i) g = __dadd_rn(kInterior, __dmul_rn(gamma, g));
        return g;
    };
    // write the cached statistics of the own tree's levels [from, plen) back to their node blocks (adding the
    // pending return if there is one): those levels leave the cache
    auto write_back = [&](int from) {
        const int top = plen;   // the pending returns count from the leaf level of the last simulation (= plen then)
        for (int l = from; l < plen; ++l) {
            const double t_ = __ldcg(PC::tot(pv, l));
            const uint32_t e = __ldcg(PC::edge(pv, l));
            const double g = pend ? ret_value(ret_j, top - l) : 0.0;
            __stcg(total_ptr<A>(blocks, (int)(e >> 2), (int)(e & 3u)), pend ? __dadd_rn(t_, g) : t_);
            // with a pending return the stored count is one behind: own_base + sims_done either way
            __stcg(na_ptr<A>(blocks, (int)(e >> 2), (int)(e & 3u)), PC::own_base(pv, l) + (int)sims_done);
            ++c_flushed;
        }
        plen = from;
        if (from == 0) pend = false;
    };

    int s_done = 0;
    // -DISLA_PHASE_TIMING: cycles per phase of a few sample warps, printed at the end (development aid)
#ifdef ISLA_PHASE_TIMING
    long long t_1a = 0, t_1b = 0, t_2 = 0, t_3 = 0, t_mark;
    long long u_issue = 0, u_wait = 0, u_eval = 0, u_settle = 0, u_tree = 0, u_mark = 0;
#define SUB_MARK() u_mark = clock64()
#define SUB_ADD(acc) do { const long long now_ = clock64(); acc += now_ - u_mark; u_mark = now_; } while (0)
#define PHASE_MARK() t_mark = clock64()
#define PHASE_ADD(acc) do { const long long now_ = clock64(); acc += now_ - t_mark; t_mark = now_; } while (0)
#else
#define PHASE_MARK()
#define PHASE_ADD(acc)
#define SUB_MARK()
#define SUB_ADD(acc)
#endif
    for (int sim = 0; sim < p.n_sim; ++sim) {
        PHASE_MARK();
        const bool live = valid && !fault;
        int mode = live ? MODE_DESCEND : MODE_IDLE;
        int level = 0;  // hand-over level of phase 1a = first level phase 1b decides

        // ------------------------------------------------------------------ phase 1a: cooperative verification
        // (shallow trees -- FrozenLake, the first simulations -- have nothing to verify in parallel: their cache is
        //  written back at once and the descent starts at the root)
        if constexpr (COOP) {
        if (live && pend && plen >= p.ret_K) write_back(0);   // deeper than the return table: no cooperative pass (rare)
        int warp_max_plen = (live && pend) ? plen : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) warp_max_plen = max(warp_max_plen, __shfl_xor_sync(kFull, warp_max_plen, o));
        if (warp_max_plen < ((p.tuning & 2) ? 1 : kCoopMin)) {   // tuning bit 2: verify cooperatively at any depth (tests)
            if (live && plen > 0) write_back(0);
        } else {
            // The rounds of all 32 trees form ONE sequence of tasks (tree j, levels l0..l0+31).  A producer cursor runs
            // kAhead tasks in front of the consumer and copies the task's PV chunk (one contiguous piece of HBM, 16 bytes
            // per lane and cp.async) into a per-warp shared-memory ring: the L2/HBM latency of a round is overlapped with
            // the evaluation of the kAhead rounds before it, across tree boundaries.  Lanes read entries other lanes
            // copied: cp.async.wait_group + __syncwarp publish a slot, a second __syncwarp frees it for the producer.
            const int my_plen = (live && pend) ? plen : 0;
            const unsigned work = __ballot_sync(kFull, my_plen > 0);
            char* const ring = coop_ring + (threadIdx.x >> 5) * (kSlots * kSlotBytes);
            // this lane's 16-byte pieces of a chunk: piece q = lane + 32 q, at the same offset in the ring slot
            const unsigned ring_dst = (unsigned)__cvta_generic_to_shared(ring) + lane * 16;
            const unsigned g_adj = PC::kBytes - lane * 8;   // from this lane's first piece of a slot to its pending return
            const char* const pv_warp = p.pv + warp_tree0 * p.pv_stride + lane * 16;
            // the own tree's pending returns: G_l = ret_table[my_ridx - l] (row ret_j, column plen - l)
            const int my_ridx = ret_j * p.ret_K + my_plen;
            int plv[PC::kCopies];               // first level (within the chunk) piece q holds data of
#pragma unroll
            for (int q = 0; q < PC::kCopies; ++q) plv[q] = lane + 32 * q < PC::kPieces ? PC::piece_level(lane + 32 * q) : (1 << 30);
            // producer cursor: tree by tree in lane order, `rem` = levels of the current tree not yet issued
            unsigned pwork = work;
            int rem = 0;
            const char* p_src = nullptr;     // this lane's first piece of the next chunk
#if ISLA_BULK
            int p_ridx = 0;                  // ret_table index of the LAST level (lowest index) of the next round
#else
            const double* p_ret = nullptr;   // this lane's pending return of the next round
#endif
            auto next_tree = [&]() {
                const int pj = __ffs(pwork) - 1;
                pwork &= pwork - 1;
                rem = __shfl_sync(kFull, my_plen, pj);
                p_src = pv_warp + (long long)pj * p.pv_stride;
#if ISLA_BULK
                p_ridx = __shfl_sync(kFull, my_ridx, pj) - 31;
#else
                p_ret = p.ret_table + (__shfl_sync(kFull, my_ridx, pj) - lane);
#endif
            };
            unsigned pdst = ring_dst;           // this lane's first piece in the producer's slot
            const unsigned ring_end = ring_dst + kSlots * kSlotBytes;
            // one task = one round of the current tree: its PV chunk (pieces that only hold levels beyond the path are
            // not copied) + the 32 pending returns, 256 contiguous bytes of one row of the return table
            auto issue_one = [&]() {
#if ISLA_BULK
                // one lane: a bulk copy of the whole chunk and one of the (16-byte aligned) window of the return table
                // that holds the round's 32 entries, both completing on the slot's mbarrier
                if (rem > 0 && lane == 0) {
                    const unsigned slot0 = pdst;    // lane 0: pdst = start of the slot
                    const unsigned mb = mbar0 + ((pdst - ring_dst) / kSlotBytes) * 8;
                    mbar_expect_tx(mb, PC::kBytes + kRetWin);
                    bulk_g2s(slot0, p_src, PC::kBytes, mb);
                    bulk_g2s(slot0 + PC::kBytes, p.ret_table + (p_ridx & ~1), kRetWin, mb);
                }
                rem -= 32; p_src += PC::kBytes; p_ridx -= 32;
#else
#pragma unroll
                for (int q = 0; q < PC::kCopies; ++q)
                    if (plv[q] < rem) cp_async16(pdst + q * 512, p_src + q * 512);
                if (lane < rem) cp_async8(pdst + g_adj, p_ret);
                rem -= 32; p_src += PC::kBytes; p_ret -= 32;
                cp_async_commit();   // one group per task (an empty one once the sequence is exhausted)
#endif
                pdst = pdst + kSlotBytes == ring_end ? ring_dst : pdst + kSlotBytes;
                if (rem <= 0 && pwork) next_tree();   // warp-uniform
            };
            next_tree();
#pragma unroll
            for (int i = 0; i < kAhead; ++i) issue_one();
            unsigned cslot = 0;                 // byte offset of the consumer's slot
            SUB_MARK();
            for (unsigned todo = work; todo; todo &= todo - 1) {
                const int j = __ffs(todo) - 1;
                const int Lj = __shfl_sync(kFull, my_plen, j);
                const int sims_j = (int)__shfl_sync(kFull, sims_done, j);   // the path edge's na (return added) = own_base + sims_j
                // this lane's total in the chunk of the iteration's first round (advances with the iterations)
                char* wr_tot = p.pv + (warp_tree0 + j) * p.pv_stride + PC::kTot + lane * 8;
#if ISLA_BULK
                // the slot holds the aligned window of the return table in table order: level l0 + i is entry 31 - i (+ parity)
                const int g_off = PC::kBytes + (31 - lane + ((__shfl_sync(kFull, my_ridx, j) + 1) & 1)) * 8;
#else
                const int g_off = PC::kBytes + lane * 8;
#endif
                int event = -1;
                constexpr int NR = A == 2 ? ISLA_COOP_NR : 2;   // rounds (levels per lane) evaluated together
                static_assert(NR >= 1 && NR <= kAhead, "rounds per iteration");
                // one level of the cached path, from this lane's entries in ring slot `slot`: adds the pending return
                // (lazy backup) and decides whether the old decision stands
                struct Lev { double total; int na; bool event; };
                auto eval_level = [&](const char* slot, int l, bool decide) -> Lev {
                    Lev out;
                    const PvSta<A> sb = decode_sta<A>(reinterpret_cast<const int4*>(slot + lane * kStaBytes));
                    const double own_t = *reinterpret_cast<const double*>(slot + PC::kTot + lane * 8);
                    const double pg = *reinterpret_cast<const double*>(slot + g_off);
                    const double new_total = __dadd_rn(own_t, pg);   // lazy backup: the return of the previous simulation
                    const int new_na = sb.own_base + sims_j;
                    out.total = new_total; out.na = new_na;
                    out.event = false;
                    if (!decide) return out;
                    const int pa = (int)(*reinterpret_cast<const uint32_t*>(slot + PC::kEdge + lane * 4) & 3u);
                    NodeBlock<A> c;
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < A; ++a) {
                        c.link[a] = 0u;
                        if (a == pa) { c.total[a] = new_total; c.na[a] = new_na; }
                        else {
                            // siblings are stored in ascending action order
                            double st_ = sb.sib_total[0]; int sn_ = sb.sib_na[0];
#pragma unroll
                            for (int q = 1; q < A - 1; ++q) if (q == k) { st_ = sb.sib_total[q]; sn_ = sb.sib_na[q]; }
                            c.total[a] = st_; c.na[a] = sn_; ++k;
                        }
                    }
                    int nun = 0;
#pragma unroll
                    for (int a = 0; a < A; ++a) nun += (c.na[a] == 0);
                    int a_sel = -1;
                    double sc[A], best;
                    const int ties = nun > 0 ? 2 : ucb(c, l > 0, a_sel, sc, best);
                    out.event = (ties != 1) || (a_sel != pa);
                    return out;
                };
                // levels before the event stay cached (with the return added); the event level and everything deeper
                // leave the cache: their statistics go back to the node blocks.  Without an event the LAST level is
                // handed over too (phase 1b re-decides it on its block).
                auto settle_level = [&](const Lev& v, int l, const char* slot, char* wr) {
                    const bool leaves = (event >= 0 && l >= event) || (event < 0 && l == Lj - 1);
                    // 8 bytes: the visit count is implicit (own_base + sims_done)
                    if (!leaves) __stcg(reinterpret_cast<double*>(wr), v.total);
                    else {
                        const uint32_t e = *reinterpret_cast<const uint32_t*>(slot + PC::kEdge + lane * 4);
                        char* const bj = p.blocks + (warp_tree0 + j) * p.slots_per_tree * 16;
                        __stcg(total_ptr<A>(bj, (int)(e >> 2), (int)(e & 3u)), v.total);
                        __stcg(na_ptr<A>(bj, (int)(e >> 2), (int)(e & 3u)), v.na);
                    }
                };
                // two rounds (64 levels) per iteration.  For two actions both levels of a lane are evaluated in ONE
                // straight-line block (loads, lazy backup, fp32 screening), so that the two dependent chains overlap in
                // the in-order issue; only undecided near-ties branch to the exact float64 comparison.
                // After an event the remaining rounds only flush: the levels leave the cache from the ring entries the
                // producer has already copied (no evaluation, no second read).
                for (int l0 = 0; l0 < Lj;) {
                    const bool decide = event < 0;   // warp-uniform
                    const int nr = min(NR, (Lj - l0 + 31) >> 5);   // rounds of this iteration
                    SUB_ADD(u_tree);
#if ISLA_BULK
                    {
                        unsigned cs = cslot;
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            if (r < nr) { mbar_wait(mbar0 + (cs / kSlotBytes) * 8, (par_bits >> (cs / kSlotBytes)) & 1u); par_bits ^= 1u << (cs / kSlotBytes); }
                            cs = cs + kSlotBytes == kSlots * kSlotBytes ? 0u : cs + kSlotBytes;
                        }
                    }
#else
                    cp_async_wait<kAhead - NR>();
                    __syncwarp();
#endif
                    SUB_ADD(u_wait);
                    const char* slot[NR];
                    bool act[NR];
                    Lev v[NR];
                    {
                        unsigned cs = cslot;
                        const int lim = Lj - l0 - lane;   // levels of the path from this lane's first one on
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            slot[r] = ring + cs;
                            cs = cs + kSlotBytes == kSlots * kSlotBytes ? 0u : cs + kSlotBytes;
                            act[r] = lim > 32 * r;        // (rounds beyond nr start past the path)
                            v[r].event = false;
                        }
                    }
                    if constexpr (A == 2) {
                        // slot = {static 32 x 16 B | totals 32 x 8 B | edge ids 32 x 4 B | pending returns 32 x 8 B}
                        // (slots beyond nr hold other trees' rounds or copies in flight: read, never used)
                        int4 sv[NR];
                        double ov[NR], gv[NR];
#pragma unroll
                        for (int r = 0; r < NR; ++r) sv[r] = *reinterpret_cast<const int4*>(slot[r] + lane * 16);
#pragma unroll
                        for (int r = 0; r < NR; ++r) ov[r] = *reinterpret_cast<const double*>(slot[r] + 512 + lane * 8);
#pragma unroll
                        for (int r = 0; r < NR; ++r) gv[r] = *reinterpret_cast<const double*>(slot[r] + g_off);
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            v[r].total = __dadd_rn(ov[r], gv[r]);   // lazy backup: the return of the previous simulation
                            v[r].na = sv[r].y + sims_j;             // own_base + simulations done
                        }
                        if (decide) {
                            // a cached level was decided by UCB1, so its sibling has been visited: sna >= 1
                            int sna[NR], ns[NR], ver[NR];
                            double st[NR];
#pragma unroll
                            for (int r = 0; r < NR; ++r) {
                                sna[r] = sv[r].x;
                                st[r] = __hiloint2double(sv[r].w, sv[r].z);
                                ns[r] = v[r].na + sna[r] + ((r > 0 || l0 + lane > 0) ? 1 : 0);
                                ver[r] = 0;   // +1 the old decision stands, -1 it changes, 0 undecided
                            }
#ifndef ISLA_NO_SCREEN
                            {
                                float lg[NR], ro[NR], rs[NR], qo[NR], qs[NR], eo[NR], es[NR];
#pragma unroll
                                for (int r = 0; r < NR; ++r) {
                                    // ns <= simulations done + 1 < ln_count on a live level; idle lanes read entry 0
                                    const int i = act[r] ? ns[r] : 0;
                                    lg[r] = p.ln_in_smem ? lnf_smem[i] : (float)__ldg(ln_table + i);
                                }
#pragma unroll
                                for (int r = 0; r < NR; ++r) { ro[r] = rcp_approx((float)v[r].na); rs[r] = rcp_approx((float)sna[r]); }
#pragma unroll
                                for (int r = 0; r < NR; ++r) { qo[r] = (float)v[r].total * ro[r]; qs[r] = (float)st[r] * rs[r]; }
#pragma unroll
                                for (int r = 0; r < NR; ++r) { eo[r] = Cf * sqrt_approx(lg[r] * ro[r]); es[r] = Cf * sqrt_approx(lg[r] * rs[r]); }
#pragma unroll
                                for (int r = 0; r < NR; ++r) {
                                    const float d = (qo[r] + eo[r]) - (qs[r] + es[r]);
                                    const float t = 1e-5f * fmaxf(fabsf(qo[r]) + fabsf(eo[r]), fabsf(qs[r]) + fabsf(es[r])) + 1e-30f;
                                    ver[r] = d > t ? 1 : (d < -t ? -1 : 0);
                                }
                            }
#endif
                            bool undecided = false;
#pragma unroll
                            for (int r = 0; r < NR; ++r) undecided |= act[r] && ver[r] == 0;
                            if (undecided) {
                                auto exact = [&](double ot, int ona, double st_, int sna_, int ns_) -> int {
                                    const double lg = ns_ < p.ln_count ? ln_table[ns_] : ::log((double)ns_);
                                    const double so = __dadd_rn(__ddiv_rn(ot, (double)ona), __dmul_rn(Cexp, __dsqrt_rn(__ddiv_rn(lg, (double)ona))));
                                    const double ss = __dadd_rn(__ddiv_rn(st_, (double)sna_), __dmul_rn(Cexp, __dsqrt_rn(__ddiv_rn(lg, (double)sna_))));
                                    return so > ss ? 1 : -1;   // a tie needs a draw: phase 1b
                                };
#pragma unroll
                                for (int r = 0; r < NR; ++r)
                                    if (act[r] && ver[r] == 0) ver[r] = exact(v[r].total, v[r].na, st[r], sna[r], ns[r]);
                            }
#pragma unroll
                            for (int r = 0; r < NR; ++r) v[r].event = act[r] && ver[r] < 0;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < NR; ++r) if (act[r]) v[r] = eval_level(slot[r], l0 + 32 * r + lane, decide);
                    }
                    SUB_ADD(u_eval);
                    if (decide) {
                        bool any = false;
#pragma unroll
                        for (int r = 0; r < NR; ++r) any |= v[r].event;
                        if (__any_sync(kFull, any)) {
#pragma unroll
                            for (int r = NR - 1; r >= 0; --r) {   // the first event from the top wins
                                const unsigned e = __ballot_sync(kFull, v[r].event);
                                if (e) event = l0 + 32 * r + __ffs(e) - 1;
                            }
                        }
                    }
                    if (event < 0 && l0 + 32 * nr < Lj) {
                        // the common case: every level of the iteration stays cached with the return added
#pragma unroll
                        for (int r = 0; r < NR; ++r) if (act[r]) __stcg(reinterpret_cast<double*>(wr_tot + r * PC::kBytes), v[r].total);
                    } else {
#pragma unroll
                        for (int r = 0; r < NR; ++r) if (act[r]) settle_level(v[r], l0 + 32 * r + lane, slot[r], wr_tot + r * PC::kBytes);
                    }
                    wr_tot += nr * PC::kBytes;
                    cslot += (unsigned)nr * kSlotBytes;
                    if (cslot >= (unsigned)(kSlots * kSlotBytes)) cslot -= (unsigned)(kSlots * kSlotBytes);
                    l0 += 32 * nr;
                    __syncwarp();   // the slots just read may be overwritten by the next issue
                    SUB_ADD(u_settle);
                    issue_one();
                    if (nr > 1) issue_one();
                    if constexpr (NR > 2) { for (int i = 2; i < nr; ++i) issue_one(); }
                    SUB_ADD(u_issue);
                }
                const int r = event >= 0 ? event : Lj - 1;
                __syncwarp();
                if (lane == j) { level = r; plen = r; pend = false; }
            }
#if !ISLA_BULK
            cp_async_wait<0>();
#endif
        }
        }   // COOP

        PHASE_ADD(t_1a);
        // ------------------------------------------------------------------ phase 1b: serial continuation
        int node_slot = 0, blk = 1, exp_a = 0, exp_rank = 0;
        bool fresh = false;  // block just allocated: all children unvisited, nothing to load
        double G = 0.0;
        if (mode == MODE_DESCEND && level > 0) {
            const uint32_t prev = __ldcg(PC::edge(pv, level - 1)), here = __ldcg(PC::edge(pv, level));
            node_slot = (int)(prev >> 2) * A + (int)(prev & 3u);
            blk = (int)(here >> 2);
        }
        while (__any_sync(kFull, mode == MODE_DESCEND)) {
            if (mode == MODE_DESCEND) {
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
                    // uniformly random unvisited action (mcts_hash.py:70-74); expanded in phase 2
                    const int pos = choice(nun);
                    int seen = 0;
#pragma unroll
                    for (int a = 0; a < A; ++a) { if (c.na[a] == 0) { if (seen == pos) exp_a = a; ++seen; } }
                    exp_rank = A - nun;
                    mode = MODE_EXPAND;
                } else {
                    double sc[A], best;
                    const int ties = ucb(c, node_slot != 0, a_sel, sc, best);
                    if (SCRIPTED || ties > 1) {
                        int pos = choice(ties);
                        if (ties > 1) {
                            // the tied set is ordered like node.actions (insertion order = rank)
#pragma unroll
                            for (int rk = 0; rk < A; ++rk) {
#pragma unroll
                                for (int a = 0; a < A; ++a) {
                                    if ((int)(c.link[a] >> kLinkRankShift) == rk && sc[a] == best) { if (pos == 0) a_sel = a; --pos; }
                                }
                            }
                        }
                    }
                    uint32_t lk = c.link[0];
#pragma unroll
                    for (int a = 1; a < A; ++a) if (a == a_sel) lk = c.link[a];
                    if (lk & kLinkTerminal) {
                        // terminal child re-visited: total += r, na += 1 (mcts_hash.py:96-107); the block is in hand
                        G = (lk & kLinkRewardOne) ? 1.0 : 0.0;
                        ret_j = Rewards<ENV>::leaf_row_terminal(G != 0.0, p.max_depth);
#pragma unroll
                        for (int a = 0; a < A; ++a) if (a == a_sel) {
                            __stcg(total_ptr<A>(blocks, blk, a), __dadd_rn(c.total[a], G));
                            __stcg(na_ptr<A>(blocks, blk, a), c.na[a] + 1);
                        }
                        mode = MODE_RETURN;
                    } else {
                        // this level joins the PV cache: own statistics + the siblings', as they are before the backup
                        {
                            PvSta<A> sb;
                            double ot = c.total[0]; int on_ = c.na[0];
                            int k = 0;
#pragma unroll
                            for (int a = 0; a < A; ++a) {
                                if (a == a_sel) { ot = c.total[a]; on_ = c.na[a]; }
                                else {
#pragma unroll
                                    for (int q = 0; q < A - 1; ++q) if (q == k) { sb.sib_total[q] = c.total[a]; sb.sib_na[q] = c.na[a]; }
                                    ++k;
                                }
                            }
                            sb.own_base = on_ - (int)sims_done;   // na with the pending return added = own_base + sims_done
                            if constexpr (COOP) {
                                store_sta<A>(PC::sta(pv, level), sb);
                                __stcg(PC::tot(pv, level), ot);
                            }
                            __stcg(PC::edge(pv, level), ((uint32_t)blk << 2) | (uint32_t)a_sel);
                        }
                        ++level;
                        plen = level;
                        node_slot = blk * A + a_sel;
                        int nb = (int)(lk & kLinkBlockMask);
                        fresh = false;
                        if (nb == 0) {
                            if (alloc >= p.block_cap) { fault = ISLA_ERR_CAPACITY; mode = MODE_IDLE; }
                            else { nb = alloc++; __stcg(link_ptr<A>(blocks, blk, a_sel), lk | (uint32_t)nb); fresh = true; }
                        }
                        blk = nb;
                    }
                }
            }
        }

        PHASE_ADD(t_1b);
        // ------------------------------------------------------------------ phase 2: expansion + rollout
        T st[S];
        T r_exp = T(0);
        bool rolling = false;
        int budget = 0;
        if (mode == MODE_EXPAND) {
            if constexpr (S == 4 && sizeof(T) == 4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(states + (long long)node_slot * S));
                st[0] = v.x; st[1] = v.y; st[2] = v.z; st[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < S; ++i) st[i] = __ldcg(states + (long long)node_slot * S + i);
            }
            int step_a = exp_a;
            if constexpr (ENV == ENV_FROZENLAKE) {
                if (p.noise) {   // is_slippery=True: step `level` of the simulation
                    if (level >= p.noise_n) fault = ISLA_ERR_INVALID; else step_a = lake_slip(exp_a, __ldg(p.noise + level));
                }
            }
            const bool term = Env<ENV, T>::step(st, T(step_a), r_exp);
            ++c_env;
            if (term) {
                G = (double)r_exp;  // terminal child: total += r, na += 1, child.ns += 1 (mcts_hash.py:96-107)
                if (G != 0.0 && G != 1.0) fault = ISLA_ERR_INVALID;
                ret_j = Rewards<ENV>::leaf_row_terminal(G != 0.0, p.max_depth);
            } else {
                if ((double)r_exp != kInterior) fault = ISLA_ERR_INVALID;  // reward model of this kernel (isla_internal.h)
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
            Stream rs;
            rs.init(p.seed, sims_done, tree_id, 0u);
            int t = 0;
            T last_r = T(0);
            bool go = rolling && budget > 0;
#ifndef ISLA_TPT_ROLL_BLOCK
#define ISLA_TPT_ROLL_BLOCK 2
#endif
#ifndef ISLA_TPT_LAKE_BLOCK
#define ISLA_TPT_LAKE_BLOCK 4
#endif
#ifndef ISLA_TPT_ROLL_BLOCK_TPW
#define ISLA_TPT_ROLL_BLOCK_TPW 16   // launches with at most this many trees per warp use the blocks (measured below)
#endif
            if (ENV == ENV_CARTPOLE && !SCRIPTED && ISLA_TPT_ROLL_BLOCK > 1 && p.tpw <= ISLA_TPT_ROLL_BLOCK_TPW) {
                // Blocks of ISLA_TPT_ROLL_BLOCK predicated steps (computed on a copy, committed by selects): straight-line
                // code in which the tail of one step (cart acceleration, position, thresholds) overlaps the head of the
                // next (the pole-angle chain) -- a lone rollout is a dependent chain, and with a few trees per warp
                // this phase is a third of the launch (blocks of 2: one tree 7.3 -> 6.3 ms, 8192 trees 398 -> 423 M sims/s,
                // 16 384 trees 539 -> 567 M; the full batch, issue- and DRAM-bound, loses 1 % and keeps the plain loop;
                // blocks of 4 / 8 waste more predicated steps than they overlap: 412 / 329 M at 8192 trees).  Step t of a rollout draws bit t of its stream (most significant
                // bit of word t / 32 first, like next_bits(1)); blocks start at multiples of the block size, so a block
                // never straddles a word.
                constexpr int NB = ISLA_TPT_ROLL_BLOCK;
                static_assert(32 % NB == 0, "a block of steps must not straddle a stream word");
                uint32_t word = 0u;
                while (__any_sync(kFull, go)) {
                    if (go && (t & 31) == 0) word = rs.next();
                    uint32_t bits = word << (t & 31);
#pragma unroll
                    for (int u = 0; u < NB; ++u) {
                        T ns[S];
#pragma unroll
                        for (int i = 0; i < S; ++i) ns[i] = st[i];
                        T r1;
                        const bool done = Env<ENV, T>::step(ns, T(bits >> 31), r1);
                        bits <<= 1;
                        if (go) {
#pragma unroll
                            for (int i = 0; i < S; ++i) st[i] = ns[i];
                            last_r = r1;
                            ++t;
                            go = !done && t < budget;
                        }
                    }
                }
            } else if (ENV == ENV_FROZENLAKE && !SCRIPTED && ISLA_TPT_LAKE_BLOCK > 1 && !p.heur && !p.noise) {
                // FrozenLake under the plain random policy: a step is a dozen instructions, the vote and branch around it
                // nearly as many -- blocks of predicated steps, two stream bits per step (16 steps per word)
                constexpr int NB = ISLA_TPT_LAKE_BLOCK;
                static_assert(16 % NB == 0, "a block of steps must not straddle a stream word");
                uint32_t word = 0u;
                while (__any_sync(kFull, go)) {
                    if (go && (t & 15) == 0) word = rs.next();
                    uint32_t bits = word << (2 * (t & 15));
#pragma unroll
                    for (int u = 0; u < NB; ++u) {
                        T ns[S];
#pragma unroll
                        for (int i = 0; i < S; ++i) ns[i] = st[i];
                        T r1;
                        const bool done = Env<ENV, T>::step(ns, T(bits >> 30), r1);
                        bits <<= 2;
                        if (go) {
#pragma unroll
                            for (int i = 0; i < S; ++i) st[i] = ns[i];
                            last_r = r1;
                            ++t;
                            go = !done && t < budget;
                        }
                    }
                }
            } else
            while (__any_sync(kFull, go)) {
                if (go) {
                    T act;
                    bool heuristic = false;
                    if constexpr (ENV == ENV_FROZENLAKE) heuristic = p.heur != nullptr;
                    if (heuristic) {   // rollout_selection_fn = grid_policy (utils/action_selection_functions.py:258-282)
                        act = T(lake_heuristic_action(p.heur, (int)st[0], [&](int ties) -> int {
                            if constexpr (SCRIPTED) {
                                const int v = (int)tc.take(ISLA_K_CHOICE);
                                if (v < 0 || v >= ties) { if (!tc.fault) tc.fault = ISLA_ERR_TRACE; return 0; }
                                return v;
                            } else {
                                return ties > 1 ? (int)mulhi_u32(rs.next(), (uint32_t)ties) : 0;
                            }
                        }));
                    } else if constexpr (SCRIPTED) {
                        const int v = (int)tc.take(ISLA_K_SAMPLE_D);
                        if (v < 0 || v >= A) { if (!tc.fault) tc.fault = ISLA_ERR_TRACE; }
                        act = T(tc.fault ? 0 : v);
                    } else {
                        act = sample_action<ENV, T>(rs);
                    }
                    if constexpr (ENV == ENV_FROZENLAKE) {
                        if (p.noise) {   // rollout step t is step level + 1 + t of the simulation
                            const int k = level + 1 + t;
                            if (k >= p.noise_n) fault = ISLA_ERR_INVALID; else act = T(lake_slip((int)act, __ldg(p.noise + k)));
                        }
                    }
                    const bool done = Env<ENV, T>::step(st, act, last_r);
                    ++t;
                    go = !done && t < budget;
                    if constexpr (SCRIPTED) { if (tc.fault) go = false; }
                }
            }
            c_roll += t; c_env += t;
            if (rolling) {
                const double R = Rewards<ENV>::rollout_return(t, (double)last_r, p.disc_table, p.cum_table);
                G = __dadd_rn((double)r_exp, __dmul_rn(gamma, R));
                ret_j = Rewards<ENV>::leaf_row_rollout(t, (double)last_r);
            }
        }
        if (mode == MODE_EXPAND) {
            // the new edge: na = 1, total = G; the child's cold record keeps the creation return
            uint32_t lk = (uint32_t)exp_rank << kLinkRankShift;
            if (!rolling) lk |= kLinkTerminal | (G == 1.0 ? kLinkRewardOne : 0u);
            __stcg(total_ptr<A>(blocks, blk, exp_a), G);
            __stcg(na_ptr<A>(blocks, blk, exp_a), 1);
            __stcg(link_ptr<A>(blocks, blk, exp_a), lk);
            __stcg(gfirst + (long long)blk * A + exp_a, rolling ? G : 0.0);
            ++c_nodes;
            mode = MODE_RETURN;
        }
        if constexpr (SCRIPTED) { if (tc.fault && !fault) fault = tc.fault; }
        if (fault) mode = MODE_IDLE;

        PHASE_ADD(t_2);
        // ------------------------------------------------------------------ phase 3: backup
        // G_l = r + gamma * G_{l+1} up the recorded path (mcts_hash.py:126-131, 79-81) is a function of (leaf value,
        // levels above the leaf) only, because the interior reward is constant: the returns are NOT computed or stored
        // here -- the next simulation's phase 1a reads them from the return table (row ret_j, column plen - l) when it
        // adds them to the cached statistics (lazy backup).
        if (mode == MODE_RETURN) {
            c_levels += (unsigned)level + 1u;
            if constexpr (COOP) {
                pend = level > 0;
            } else {
                // eager backup along the recorded edges: total += G_l, na += 1 (the leaf edge itself was updated above)
                for (int l = level - 1; l >= 0; --l) {
                    const uint32_t e = __ldcg(PC::edge(pv, l));
                    double* const tp = total_ptr<A>(blocks, (int)(e >> 2), (int)(e & 3u));
                    int32_t* const np = na_ptr<A>(blocks, (int)(e >> 2), (int)(e & 3u));
                    __stcg(tp, __dadd_rn(__ldcg(tp), ret_value(ret_j, level - l)));
                    __stcg(np, __ldcg(np) + 1);
                }
                plen = 0; pend = false;
            }
            ++sims_done; ++s_done;
        }
        __syncwarp();
        PHASE_ADD(t_3);
    }
#ifdef ISLA_PHASE_TIMING
    if ((blockIdx.x % 197) == 0 && threadIdx.x == 0)
        printf("block %d cycles 1a %lld 1b %lld 2 %lld 3 %lld | 1a: tree %lld issue %lld wait %lld eval %lld settle %lld\n", blockIdx.x, t_1a, t_1b, t_2, t_3,
               u_tree, u_issue, u_wait, u_eval, u_settle);
#endif

    if (valid) {
        // leave the tree consistent for the host / the next launch: the cached levels go back to their blocks
        if (plen > 0) write_back(0);
        meta[META_ALLOC] = alloc;
        meta[META_SIMS] = (int32_t)sims_done;
        meta[META_RNG] = (int32_t)rng_used;
        meta[META_FAULT] = fault;
        meta[META_PENDING] = 0;
        if constexpr (SCRIPTED) meta[META_TRACE_POS] = (int32_t)tc.pos;
        // counters (once per thread at kernel end)
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
__global__ void tpt_set_roots_kernel(const double* roots, T* states, int32_t* meta, long long n_trees,
                                     long long slots_per_tree, int S) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    for (int i = 0; i < S; ++i) states[tree * slots_per_tree * S + i] = (T)roots[tree * S + i];
    int32_t* m = meta + tree * META_WORDS;
    m[META_ALLOC] = 2;  // block 0 is reserved (link 0 = none), block 1 holds the root's edges
    m[META_SIMS] = 0; m[META_RNG] = 0; m[META_FAULT] = 0; m[4] = 0; m[5] = 0; m[6] = 0; m[7] = 0;
}

__global__ void tpt_root_stats_kernel(const char* blocks, long long n_trees, long long slots_per_tree, int A,
                                      int64_t* na_out, double* total_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_trees * A) return;
    const long long tree = i / A;
    const int a = (int)(i % A);
    const char* b = blocks + tree * slots_per_tree * 16 + (long long)A * 16;  // block 1
    na_out[i] = reinterpret_cast<const int32_t*>(b + A * 8)[a];
    total_out[i] = reinterpret_cast<const double*>(b)[a];
}

// end of fit(): q = total/na over the root's children, uniformly random arg-max (mcts_hash.py:42-50)
__global__ void tpt_best_kernel(const char* blocks, int32_t* meta, long long n_trees, long long slots_per_tree, int A,
                                unsigned long long seed, unsigned long long tree_id_base, int scripted_pos,
                                int32_t* best_out, double* best_action_out) {
    const long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    const char* b = blocks + tree * slots_per_tree * 16 + (long long)A * 16;  // block 1
    double q[16];
    int rank[16];
    double best = -INFINITY;
    int ties = 0;
    for (int a = 0; a < A; ++a) {
        const int na = reinterpret_cast<const int32_t*>(b + A * 8)[a];
        const double tot = reinterpret_cast<const double*>(b)[a];
        const uint32_t lk = reinterpret_cast<const uint32_t*>(b + A * 12)[a];
        rank[a] = na > 0 ? (int)(lk >> kLinkRankShift) : -1;
        q[a] = na > 0 ? __ddiv_rn(tot, (double)na) : -INFINITY;
        if (na > 0) best = fmax(best, q[a]);
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

// Post-fit depth analytics of one tree (agents/abstract_mcts.py:120-142,190-209; sweep.py:199-200) on the node blocks,
// level-synchronous like the CTA engine's.  A visited edge whose child has no block yet, or is terminal, is a leaf state
// one level below its block.  Scratch: the tree's edge-id row of the PV cache (free between launches), as {level, owner}.
__global__ void __launch_bounds__(256) tpt_depth_stats_kernel(const char* blocks_all, const int32_t* meta_all, char* scratch_all,
                                                              long long slots_per_tree, long long scratch_stride, int A, long long tree,
                                                              int cap, int32_t* dmax, double* dmean, int32_t* count) {
    const char* blocks = blocks_all + tree * slots_per_tree * 16;
    const int alloc = meta_all[tree * META_WORDS + META_ALLOC];
    int2* const tag = reinterpret_cast<int2*>(scratch_all + tree * scratch_stride);   // {level, root-child rank} per block
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int changed, n_root;
    __shared__ int leaves[4];
    for (int b = tid; b < alloc; b += nt) tag[b] = make_int2(-1, -1);
    if (tid < 4) leaves[tid] = 0;
    if (tid == 0) {
        int n = 0;
        for (int a = 0; a < A; ++a) n += reinterpret_cast<const int32_t*>(blocks + (long long)A * 16 + A * 8)[a] > 0;
        n_root = n; *count = n;
        for (int i = 0; i < n && i < cap; ++i) { dmax[i] = 0; dmean[i] = 0.0; }
    }
    __syncthreads();
    if (tid == 0 && alloc > 1) tag[1] = make_int2(0, -1);
    __syncthreads();
    for (int cur = 0;; ++cur) {
        if (tid == 0) changed = 0;
        __syncthreads();
        for (int b = 1 + tid; b < alloc; b += nt) {
            if (tag[b].x != cur) continue;
            const char* blk = blocks + (long long)b * (A * 16);
            for (int a = 0; a < A; ++a) {
                if (reinterpret_cast<const int32_t*>(blk + A * 8)[a] <= 0) continue;
                const uint32_t lk = reinterpret_cast<const uint32_t*>(blk + A * 12)[a];
                const int o = b == 1 ? (int)(lk >> kLinkRankShift) : tag[b].y;
                const int nb = (int)(lk & kLinkBlockMask);
                if (o < cap) atomicMax(dmax + o, cur + 1);
                if (nb != 0 && !(lk & kLinkTerminal)) { tag[nb] = make_int2(cur + 1, o); changed = 1; }
                else if (o < cap) { atomicAdd(leaves + o, 1); atomicAdd(dmean + o, (double)(cur + 1)); }
            }
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    if (tid < n_root && tid < cap) dmean[tid] = leaves[tid] ? __ddiv_rn(dmean[tid], (double)leaves[tid]) : 0.0;
}

// Tree reuse between real steps (build extension; the reference builds a new agent per step,
// sweep.py:186): the subtree below the root edge `action[tree]` becomes the tree.  One CTA per tree copies the subtree
// breadth first into a scratch arena -- blocks renumbered in visiting order, the chosen child's block becomes block 1 and
// its state the root state (slot 0) -- and then back over the tree's own arena, zeroing what is left of the old blocks
// (fresh blocks are assumed zero).  Statistics are copied bit for bit; the re-rooted node is searched as a root from
// now on (its ns = the sum of its edges' visits).  Scratch per tree: the PV area (the queue; free between launches) and
// `tmp` (one tree's node blocks, states and creation returns).
template <typename T>
__global__ void __launch_bounds__(128) tpt_reroot_kernel(char* blocks_all, T* states_all, double* gfirst_all, int32_t* meta_all,
                                                         char* pv_all, long long pv_stride, long long slots_per_tree, int A, int S,
                                                         const int32_t* __restrict__ action, const uint8_t* __restrict__ alive,
                                                         char* tmp_all, long long tmp_stride, long long n_trees, int max_blocks,
                                                         int max_visits) {
    const long long tree = blockIdx.x;
    if (tree >= n_trees || (alive && !alive[tree])) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    char* const blocks = blocks_all + tree * slots_per_tree * 16;
    T* const states = states_all + tree * slots_per_tree * S;
    double* const gfirst = gfirst_all + tree * slots_per_tree;
    int32_t* const meta = meta_all + tree * META_WORDS;
    const int alloc = meta[META_ALLOC];
    const int blk_bytes = A * 16;
    int* const queue = reinterpret_cast<int*>(pv_all + tree * pv_stride);   // new block - 1 -> old block
    char* const tblocks = tmp_all + tree * tmp_stride;
    T* const tstates = reinterpret_cast<T*>(tblocks + slots_per_tree * 16);
    double* const tg = reinterpret_cast<double*>(reinterpret_cast<char*>(tstates) + slots_per_tree * S * (long long)sizeof(T));
    __shared__ int head, tail, lvl_end;
    const int a0 = min(max(action[tree], 0), A - 1);
    const uint32_t lk0 = *link_ptr_rt(blocks, 1, a0, A);
    const int na0 = *na_ptr_rt(blocks, 1, a0, A);
    int cb0 = (na0 > 0 && !(lk0 & kLinkTerminal)) ? (int)(lk0 & kLinkBlockMask) : 0;
    // a subtree that would not leave room for the coming simulations (blocks, or visit counts beyond the ln(n) table) is
    // dropped: the search restarts from the bare child state, as the per-step rebuild would
    if (cb0 && na0 - 1 > max_visits) cb0 = 0;
    // the new root's state: the chosen child's (slot 1 * A + a0 of the old tree)
    if (tid < S) tstates[tid] = states[(long long)(1 * A + a0) * S + tid];
    if (tid == 0) { head = 0; tail = 0; if (cb0) { queue[0] = cb0; tail = 1; } }
    __syncthreads();
    for (;;) {
        if (tid == 0) lvl_end = tail;
        __syncthreads();
        const int lo = head, hi = lvl_end;
        if (lo >= hi) break;
        for (int qi = lo + tid; qi < hi; qi += nt) {
            const int ob = queue[qi], nb = qi + 1;
            for (int a = 0; a < A; ++a) {
                const double tot = *total_ptr_rt(blocks, ob, a, A);
                const int na = *na_ptr_rt(blocks, ob, a, A);
                uint32_t lk = *link_ptr_rt(blocks, ob, a, A);
                const int cb = (int)(lk & kLinkBlockMask);
                if (cb != 0) {   // (a block has exactly one parent edge: its new index is assigned here)
                    const int pos = atomicAdd(&tail, 1);
                    queue[pos] = cb;
                    lk = (lk & ~kLinkBlockMask) | (uint32_t)(pos + 1);
                }
                *total_ptr_rt(tblocks, nb, a, A) = tot;
                *na_ptr_rt(tblocks, nb, a, A) = na;
                *link_ptr_rt(tblocks, nb, a, A) = lk;
                for (int k = 0; k < S; ++k) tstates[(long long)(nb * A + a) * S + k] = states[(long long)(ob * A + a) * S + k];
                tg[(long long)nb * A + a] = gfirst[(long long)ob * A + a];
            }
        }
        __syncthreads();
        if (tid == 0) head = hi;
        __syncthreads();
    }
    int n_new = tail;                  // blocks of the new tree: 1 .. n_new (none: the root has no edges yet)
    if (n_new > max_blocks) n_new = 0; // (see above: too large to keep)
    __syncthreads();
    // back over the tree's own arena
    const int new_alloc = (n_new > 0 ? n_new : 1) + 1;
    for (int i = tid; i < (new_alloc * blk_bytes) / 16; i += nt) {
        int4 v = make_int4(0, 0, 0, 0);
        if (i >= blk_bytes / 16 && i < ((n_new + 1) * blk_bytes) / 16) v = reinterpret_cast<const int4*>(tblocks)[i];
        reinterpret_cast<int4*>(blocks)[i] = v;
    }
    for (int i = (new_alloc * blk_bytes) / 16 + tid; i < (alloc * blk_bytes) / 16; i += nt) reinterpret_cast<int4*>(blocks)[i] = make_int4(0, 0, 0, 0);
    for (long long i = tid; i < (long long)(n_new + 1) * A * S; i += nt) if (i < S || i >= (long long)A * S) states[i] = tstates[i];
    for (long long i = (long long)A + tid; i < (long long)(n_new + 1) * A; i += nt) gfirst[i] = tg[i];
    if (tid == 0) meta[META_ALLOC] = new_alloc;
}

template <int ENV, typename T>
int launch_search(TptParams& p, bool scripted, int sm_count, int tpw_forced, cudaStream_t st) {
    constexpr int A = EnvTraits<ENV>::A;
    // FrozenLake's trees are too shallow for the cooperative phase (tuning bit 2 forces it: tests)
    const bool coop = !scripted && (ENV != ENV_FROZENLAKE || (p.tuning & 2));
    const size_t smem = (p.ln_in_smem ? (size_t)((p.ln_count + 3) & ~3) * 4 : 0) + (coop ? (size_t)(kTptBlock / 32) * CoopRing<A>::kWarpBytes : 0);
    if (scripted) {
        p.tpw = 1;
        tpt_search_kernel<ENV, T, true, false><<<1, 32, smem, st>>>(p);
    } else {
        // Trees per warp.  A launch is bound by the serial instruction stream of a warp, and the phases that run one
        // lane per tree (serial continuation, expansion, rollout) cost the same for 1 or 32 busy lanes: so as many
        // warps as the GPU keeps resident at once, each with as few trees as that allows -- n_trees / resident warps,
        // rounded up (8192 trees: 4, 16 384: 7, 65 536: 28 on a B200 at 8 CTAs per SM); batches beyond 32 trees per
        // resident warp run in equal waves.  Measured: profiles/README.md (round 2).
        int occ = 0;
        const cudaError_t oe = coop ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tpt_search_kernel<ENV, T, false, true>, kTptBlock, smem)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tpt_search_kernel<ENV, T, false, false>, kTptBlock, smem);
        if (oe != cudaSuccess || occ < 1) {
            cudaGetLastError();
            occ = ISLA_TPT_MINBLOCKS;
        }
        const long long resident = (long long)(sm_count > 0 ? sm_count : 148) * occ * (kTptBlock / 32);
        const long long waves = (p.n_trees + 32 * resident - 1) / (32 * resident);
        long long tpw = (p.n_trees + waves * resident - 1) / (waves * resident);
        if (kTreesPerWarpFixed > 0) tpw = kTreesPerWarpFixed;
        if (tpw_forced >= 1 && tpw_forced <= 32) tpw = tpw_forced;   // tuning bits 8..13 (A/B timing, tests)
        p.tpw = (int)(tpw < 1 ? 1 : (tpw > 32 ? 32 : tpw));
        const int kTreesPerBlock = (kTptBlock / 32) * p.tpw;
        const unsigned grid = (unsigned)((p.n_trees + kTreesPerBlock - 1) / kTreesPerBlock);
        if (coop) tpt_search_kernel<ENV, T, false, true><<<grid, kTptBlock, smem, st>>>(p);
        else tpt_search_kernel<ENV, T, false, false><<<grid, kTptBlock, smem, st>>>(p);
    }
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : cuda_fail(err, "tpt_search_kernel launch");
}

}  // namespace

// bytes of one tree's PV chunks
static long long tpt_pv_stride(long long blocks, int A) {
    return (blocks + 31) / 32 * (A == 2 ? PvChunk<2>::kBytes : PvChunk<4>::kBytes);
}

int tpt_fill_layout(const isla_config& cfg, isla_layout& lay) {
    const int A = env_n_actions(cfg.env_id), S = env_state_dim(cfg.env_id);
    if (A == 0 || cfg.agent != ISLA_AGENT_VANILLA) { set_error("thread-per-tree kernel: vanilla agent on a discrete env only"); return ISLA_ERR_INVALID; }
    if (cfg.n_actions != A) { set_error("n_actions=%d does not match env (%d)", cfg.n_actions, A); return ISLA_ERR_INVALID; }
    if (cfg.rollouts_per_leaf > 1) { set_error("thread-per-tree kernel: rollouts_per_leaf must be 1"); return ISLA_ERR_INVALID; }
    if (cfg.tree_parallel > 1) { set_error("thread-per-tree kernel: tree_parallel needs the CTA-per-tree kernel"); return ISLA_ERR_INVALID; }
    const int64_t blocks = (int64_t)cfg.sim_capacity + 2;
    if (blocks >= (int64_t)kLinkBlockMask) { set_error("sim_capacity too large for the 28-bit block index"); return ISLA_ERR_INVALID; }
    const int rb = cfg.precision == ISLA_PRECISION_F64 ? 8 : 4;
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    lay.kernel = ISLA_KERNEL_THREAD_PER_TREE;
    lay.state_dim = S; lay.real_bytes = rb; lay.n_actions = A; lay.action_dim = 1;
    lay.slots_per_tree = blocks * A;
    int64_t off = 0;
    lay.rec_offset = off; lay.rec_stride = lay.slots_per_tree * 16; off = up(off + cfg.n_trees * lay.rec_stride);
    lay.state_offset = off; lay.state_stride = lay.slots_per_tree * S * rb; off = up(off + cfg.n_trees * lay.state_stride);
    lay.gfirst_offset = off; lay.gfirst_stride = lay.slots_per_tree * 8; off = up(off + cfg.n_trees * lay.gfirst_stride);
    lay.meta_offset = off; off = up(off + cfg.n_trees * META_WORDS * 4);
    // PV cache: per tree ceil(blocks / 32) chunks of 32 levels (static entries, totals, pending returns, edge ids)
    lay.path_offset = off; off = up(off + cfg.n_trees * tpt_pv_stride(blocks, A));
    lay.pathg_offset = lay.pvown_offset = lay.path_offset;   // pending returns and own totals live inside the chunks
    lay.counters_offset = off; off = up(off + 8 * 8);
    add_table_layout(cfg, lay, off);
    lay.s_cap = lay.a_cap = 0; lay.pool_offset = lay.pool_stride = 0;
    lay.total_bytes = off;
    return 0;
}

int tpt_set_roots(isla_engine* e, const double* roots_dev, cudaStream_t st) {
    const isla_layout& L = e->lay;
    // node blocks + counters cleared (fresh blocks are assumed zero); states / scratch need no clearing
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
    p.blocks = e->ws + L.rec_offset;
    p.states = e->ws + L.state_offset;
    p.gfirst = (double*)(e->ws + L.gfirst_offset);
    p.meta = (int32_t*)(e->ws + L.meta_offset);
    p.pv = e->ws + L.path_offset;
    p.pv_stride = tpt_pv_stride((long long)e->cfg.sim_capacity + 2, L.n_actions);
    p.counters = (unsigned long long*)(e->ws + L.counters_offset);
    p.trace_kinds = kinds; p.trace_values = values; p.trace_len = trace_len;
    p.n_trees = e->cfg.n_trees; p.slots_per_tree = L.slots_per_tree; p.scripted_tree = scripted_tree;
    p.seed = e->cfg.seed; p.tree_id_base = e->cfg.tree_id_base;
    p.C = e->cfg.C; p.gamma = e->cfg.gamma;
    p.n_sim = n_sim; p.max_depth = e->cfg.max_depth; p.budget_mode = e->cfg.budget_mode;
    p.block_cap = e->cfg.sim_capacity + 2;
    p.ln_table = (const double*)(e->ws + L.ln_offset);
    p.disc_table = (const double*)(e->ws + L.disc_offset);
    p.cum_table = (const double*)(e->ws + L.cumdisc_offset);
    p.ln_count = (int)L.ln_count; p.disc_count = (int)L.disc_count;
    p.ret_table = e->ret_dev + kRetTablePad; p.ret_K = e->ret_K;
    p.ln_in_smem = (L.ln_count * 4 <= 8 * 1024) ? 1 : 0;
    p.heur = e->heur_dev;
    p.noise = e->noise_dev; p.noise_n = e->noise_n;
    const int tpw_forced = (e->cfg.tuning >> 8) & 63;   // tuning bits 8..13: trees per warp (A/B timing, tests)
    p.tpw = 1;
    p.tuning = e->cfg.tuning;
    const bool scripted = scripted_tree >= 0;
    const bool f64 = e->cfg.precision == ISLA_PRECISION_F64;
    switch (e->cfg.env_id) {
    case ISLA_ENV_CARTPOLE:
        return f64 ? launch_search<ENV_CARTPOLE, double>(p, scripted, e->sm_count, tpw_forced, st)
                   : launch_search<ENV_CARTPOLE, float>(p, scripted, e->sm_count, tpw_forced, st);
    case ISLA_ENV_FROZENLAKE:
        return f64 ? launch_search<ENV_FROZENLAKE, double>(p, scripted, e->sm_count, tpw_forced, st)
                   : launch_search<ENV_FROZENLAKE, float>(p, scripted, e->sm_count, tpw_forced, st);
    default:
        set_error("thread-per-tree kernel: env %d unsupported", e->cfg.env_id);
        return ISLA_ERR_INVALID;
    }
}

int tpt_reroot(isla_engine* e, const int32_t* action_dev, const uint8_t* alive_dev, int n_sim_next, cudaStream_t st) {
    const isla_layout& L = e->lay;
    const long long rb = L.real_bytes;
    const long long tmp_stride = (L.slots_per_tree * (16 + L.state_dim * rb + 8) + 255) / 256 * 256;
    const size_t need = (size_t)(e->cfg.n_trees * tmp_stride);
    if (e->reroot_bytes < need) {
        ISLA_CUDA_OK(cudaStreamSynchronize(st));
        if (e->reroot_tmp) cudaFree(e->reroot_tmp);
        e->reroot_tmp = nullptr; e->reroot_bytes = 0;
        ISLA_CUDA_OK(cudaMalloc((void**)&e->reroot_tmp, need));
        e->reroot_bytes = need;
    }
    const long long pvs = tpt_pv_stride((long long)e->cfg.sim_capacity + 2, L.n_actions);
    if (pvs < (long long)(e->cfg.sim_capacity + 2) * 4) { set_error("reroot: PV scratch too small"); return ISLA_ERR_INVALID; }
    const unsigned grid = (unsigned)e->cfg.n_trees;
    const int max_blocks = e->cfg.sim_capacity - n_sim_next - 2, max_visits = e->cfg.sim_capacity - n_sim_next - 2;
    if (max_blocks < 0) { set_error("reroot: sim_capacity %d leaves no room for %d more simulations", e->cfg.sim_capacity, n_sim_next); return ISLA_ERR_INVALID; }
    if (rb == 4)
        tpt_reroot_kernel<float><<<grid, 128, 0, st>>>(e->ws + L.rec_offset, (float*)(e->ws + L.state_offset), (double*)(e->ws + L.gfirst_offset),
            (int32_t*)(e->ws + L.meta_offset), e->ws + L.path_offset, pvs, L.slots_per_tree, L.n_actions, L.state_dim, action_dev, alive_dev,
            e->reroot_tmp, tmp_stride, e->cfg.n_trees, max_blocks, max_visits);
    else
        tpt_reroot_kernel<double><<<grid, 128, 0, st>>>(e->ws + L.rec_offset, (double*)(e->ws + L.state_offset), (double*)(e->ws + L.gfirst_offset),
            (int32_t*)(e->ws + L.meta_offset), e->ws + L.path_offset, pvs, L.slots_per_tree, L.n_actions, L.state_dim, action_dev, alive_dev,
            e->reroot_tmp, tmp_stride, e->cfg.n_trees, max_blocks, max_visits);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int tpt_root_stats(isla_engine* e, int64_t* na_dev, double* total_dev, cudaStream_t st) {
    const isla_layout& L = e->lay;
    const long long n = e->cfg.n_trees * L.n_actions;
    tpt_root_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->ws + L.rec_offset, e->cfg.n_trees,
                                                                       L.slots_per_tree, L.n_actions, na_dev, total_dev);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int tpt_depth_stats(isla_engine* e, long long tree, int cap, int32_t* dmax, double* dmean, int32_t* count, cudaStream_t st) {
    const isla_layout& L = e->lay;
    tpt_depth_stats_kernel<<<1, 256, 0, st>>>(e->ws + L.rec_offset, (const int32_t*)(e->ws + L.meta_offset), e->ws + L.path_offset,
                                            L.slots_per_tree, tpt_pv_stride(L.slots_per_tree / L.n_actions, L.n_actions), L.n_actions, tree, cap, dmax, dmean, count);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

int tpt_best(isla_engine* e, int scripted_pos, int32_t* best_dev, double* best_action_dev, cudaStream_t st) {
    const isla_layout& L = e->lay;
    tpt_best_kernel<<<(unsigned)((e->cfg.n_trees + 255) / 256), 256, 0, st>>>(
        e->ws + L.rec_offset, (int32_t*)(e->ws + L.meta_offset), e->cfg.n_trees, L.slots_per_tree,
        L.n_actions, e->cfg.seed, e->cfg.tree_id_base, scripted_pos, best_dev, best_action_dev);
    ISLA_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace isla
