// Data layout and small device helpers of the thread-per-tree search kernel (isla_tree_tpt.cu; kept in a header since the
// worker / owner A/B variant of round 2 shared them): node blocks, the PV cache chunks, the cooperative ring, cp.async.
#pragma once
#include <cmath>
#include <cstring>

#include "isla_envs.cuh"
#include "isla_internal.h"
#include "isla_philox.cuh"

namespace isla {
namespace {


#ifndef ISLA_TPT_BLOCK
#define ISLA_TPT_BLOCK 64
#endif
constexpr int kTptBlock = ISLA_TPT_BLOCK;   // threads per CTA (32 / 128 with matching ISLA_TPT_MINBLOCKS: no faster, profiles/README.md)
// Trees owned by one warp (its first tpw lanes), chosen per launch.  A launch of this kernel is bound by the serial
// instruction stream of a warp, most of it the cooperative phase, whose length is proportional to the warp's trees.  A
// batch that does not fill the GPU's resident warps (148 SMs x 14) with 32 trees each is therefore spread over more
// warps with fewer trees each: the phases that run one lane per tree then use part of the warp, which costs nothing
// while latency-bound (4096 CartPole trees: 44.6 -> 206 M sims/s).  ISLA_TPW pins it (experiments).
#ifndef ISLA_TPW
#define ISLA_TPW 0
#endif
#ifndef ISLA_TPT_MINBLOCKS
#define ISLA_TPT_MINBLOCKS 7
#endif
constexpr int kTreesPerWarpFixed = ISLA_TPW;
constexpr unsigned kFull = 0xffffffffu;

struct TraceCursor {
    const uint8_t* k;
    const double* v;
    long long len, pos;
    int fault;
    __device__ __forceinline__ double take(int kind) {
        if (fault) return 0.0;
        if (pos >= len) { fault = ISLA_ERR_TRACE; return 0.0; }
        if (k[pos] != kind) { fault = ISLA_ERR_TRACE; return 0.0; }
        return v[pos++];
    }
};

template <int A> struct NodeBlock {
    double total[A];
    int32_t na[A];
    uint32_t link[A];
};
static_assert(sizeof(NodeBlock<2>) == 32 && sizeof(NodeBlock<4>) == 64, "node block is one or two sectors");

template <int A>
__device__ __forceinline__ NodeBlock<A> load_block(const char* base, int blk) {
    // L2-only loads (ld.global.cg): blocks are also updated by RED atomics at L2
    NodeBlock<A> b;
    const int4* p = reinterpret_cast<const int4*>(base + (long long)blk * (A * 16));
    if constexpr (A == 2) {
        const int4 lo = __ldcg(p), hi = __ldcg(p + 1);
        b.total[0] = __hiloint2double(lo.y, lo.x); b.total[1] = __hiloint2double(lo.w, lo.z);
        b.na[0] = hi.x; b.na[1] = hi.y; b.link[0] = (uint32_t)hi.z; b.link[1] = (uint32_t)hi.w;
    } else {
        const int4 t0 = __ldcg(p), t1 = __ldcg(p + 1), n = __ldcg(p + 2), l = __ldcg(p + 3);
        b.total[0] = __hiloint2double(t0.y, t0.x); b.total[1] = __hiloint2double(t0.w, t0.z);
        b.total[2] = __hiloint2double(t1.y, t1.x); b.total[3] = __hiloint2double(t1.w, t1.z);
        b.na[0] = n.x; b.na[1] = n.y; b.na[2] = n.z; b.na[3] = n.w;
        b.link[0] = (uint32_t)l.x; b.link[1] = (uint32_t)l.y; b.link[2] = (uint32_t)l.z; b.link[3] = (uint32_t)l.w;
    }
    return b;
}
template <int A> __device__ __forceinline__ double* total_ptr(char* base, int blk, int a) {
    return reinterpret_cast<double*>(base + (long long)blk * (A * 16)) + a;
}
template <int A> __device__ __forceinline__ int32_t* na_ptr(char* base, int blk, int a) {
    return reinterpret_cast<int32_t*>(base + (long long)blk * (A * 16) + A * 8) + a;
}
template <int A> __device__ __forceinline__ uint32_t* link_ptr(char* base, int blk, int a) {
    return reinterpret_cast<uint32_t*>(base + (long long)blk * (A * 16) + A * 12) + a;
}

// the same accessors with the action count as a run-time value (epilogue / maintenance kernels)
__device__ __forceinline__ double* total_ptr_rt(char* base, int blk, int a, int A) { return reinterpret_cast<double*>(base + (long long)blk * (A * 16)) + a; }
__device__ __forceinline__ int32_t* na_ptr_rt(char* base, int blk, int a, int A) { return reinterpret_cast<int32_t*>(base + (long long)blk * (A * 16) + A * 8) + a; }
__device__ __forceinline__ uint32_t* link_ptr_rt(char* base, int blk, int a, int A) { return reinterpret_cast<uint32_t*>(base + (long long)blk * (A * 16) + A * 12) + a; }

// cp.async with the destination given as a 32-bit shared-window address
__device__ __forceinline__ void cp_async16(unsigned smem, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem), "l"(g) : "memory");
}
// MUFU approximations (relative error ~2^-22) for the fp32 screening pass of UCB1; exact float64 decides near-ties
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-memory ring of the cooperative phase, per warp: kSlots slots of {static entries 32 x sizeof(PvSta) | totals
// 32 x 8 B | pending returns 32 x 8 B}; the producer runs kAhead = kSlots - 1 tasks in front of the consumer
#ifndef ISLA_COOP_SLOTS
#define ISLA_COOP_SLOTS 8
#endif
#ifndef ISLA_COOP_NR
#define ISLA_COOP_NR 2
#endif
// -DISLA_BULK=1: the producer of the cooperative phase is ONE lane issuing cp.async.bulk copies that complete on an
// mbarrier per ring slot, instead of 32 lanes x 3 cp.async (A/B: profiles/README.md)
#ifndef ISLA_BULK
#define ISLA_BULK 0
#endif
constexpr int kRetWin = ISLA_BULK ? 272 : 256;   // bytes of pending returns per slot (bulk: 16-byte aligned window of 34)
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(mbar), "r"(parity) : "memory");
}
template <int A> struct CoopRing {
    static constexpr int kSlots = A == 2 ? ISLA_COOP_SLOTS : 3;
    static constexpr int kAhead = kSlots - 1;
    static constexpr int kChunkBytes = 32 * (A == 2 ? 16 : 48) + 32 * 8 + 32 * 4;   // = one PV chunk (HBM)
    static constexpr int kSlotBytes = kChunkBytes + kRetWin;                         // + the 32 pending returns (from the return table)
    static constexpr int kWarpBytes = kSlots * kSlotBytes + (ISLA_BULK ? kSlots * 8 : 0);   // + one mbarrier per slot
};

enum : int { MODE_DESCEND = 0, MODE_EXPAND = 1, MODE_RETURN = 2, MODE_IDLE = 3 };
constexpr int kCoopMin = 16;  // deepest cached path in the warp below which phase 1a is skipped

// PV cache (per tree, per level of the last simulation's path; tree-major rows in HBM):
//   tot    8 B  double total of the edge the path takes -- the AUTHORITATIVE copy while the level is cached (changes
//               every simulation: the pending return of the previous simulation is added when the level is re-evaluated)
//   sta   16 B (A=2) / 48 B (A=4)  {int32 sib_na[A-1]; int32 own_base; double sib_total[A-1]}   static while cached:
//               the sibling edges' statistics, and own_base = (path edge's na) - (tree's completed simulations) at
//               caching time.  Every simulation passes through a cached level exactly once, so the path edge's visit
//               count is own_base + simulations_completed: it is never stored again.
//   g      8 B  the return backed up through the edge, pending until the next simulation
//   edge   4 B  block << 2 | action of the path edge: cold, read only where a level leaves the cache
// While a level is cached its node block in HBM is stale for that one edge; it is written back when the level leaves
// the path (or at the end of the launch).  Per level and simulation: 32 bytes read, 8 + 8 bytes written.
template <int A> struct __align__(16) PvSta { int32_t sib_na[A - 1]; int32_t own_base; double sib_total[A - 1]; };
static_assert(sizeof(PvSta<2>) == 16 && sizeof(PvSta<4>) == 48, "PV cache entry sizes");

template <int A> __device__ __forceinline__ PvSta<A> decode_sta(const int4* p) {
    PvSta<A> s;
    if constexpr (A == 2) {
        const int4 v = p[0];
        s.sib_na[0] = v.x; s.own_base = v.y; s.sib_total[0] = __hiloint2double(v.w, v.z);
    } else {
        const int4 v0 = p[0], v1 = p[1], v2 = p[2];
        s.sib_na[0] = v0.x; s.sib_na[1] = v0.y; s.sib_na[2] = v0.z; s.own_base = v0.w;
        s.sib_total[0] = __hiloint2double(v1.y, v1.x); s.sib_total[1] = __hiloint2double(v1.w, v1.z);
        s.sib_total[2] = __hiloint2double(v2.y, v2.x);
    }
    return s;
}
template <int A> __device__ __forceinline__ void store_sta(char* entry, const PvSta<A>& s) {
    int4* p = reinterpret_cast<int4*>(entry);
    if constexpr (A == 2) {
        __stcg(p, make_int4(s.sib_na[0], s.own_base, __double2loint(s.sib_total[0]), __double2hiint(s.sib_total[0])));
    } else {
        __stcg(p, make_int4(s.sib_na[0], s.sib_na[1], s.sib_na[2], s.own_base));
        __stcg(p + 1, make_int4(__double2loint(s.sib_total[0]), __double2hiint(s.sib_total[0]), __double2loint(s.sib_total[1]), __double2hiint(s.sib_total[1])));
        __stcg(p + 2, make_int4(__double2loint(s.sib_total[2]), __double2hiint(s.sib_total[2]), 0, 0));
    }
}
// HBM layout of the PV cache: per tree, CHUNKS of 32 consecutive levels, each chunk laid out exactly like a ring slot
//   {static entries 32 x sizeof(PvSta) | totals 32 x 8 B | pending returns 32 x 8 B | edge ids 32 x 4 B}
// so that a round of the cooperative phase is ONE contiguous copy (1152 B for A = 2) through a single pointer.
template <int A> struct PvChunk {
    static constexpr int kSta = (int)sizeof(PvSta<A>);
    static constexpr int kTot = 32 * kSta, kEdge = kTot + 256, kBytes = kEdge + 128;
    static constexpr int kPieces = kBytes / 16;              // 16-byte pieces of a chunk (56 / 120)
    static constexpr int kCopies = (kPieces + 31) / 32;      // cp.async16 per lane and chunk (2 / 4)
    // first level (within the chunk) a 16-byte piece holds data of
    __host__ __device__ static constexpr int piece_level(int piece) {
        return piece * 16 < kTot ? piece * 16 / kSta : (piece * 16 < kEdge ? (piece * 16 - kTot) / 8 : (piece * 16 - kEdge) / 4);
    }
    __device__ __forceinline__ static char* chunk(char* pv, int l) { return pv + (long long)(l >> 5) * kBytes; }
    __device__ __forceinline__ static char* sta(char* pv, int l) { return chunk(pv, l) + (l & 31) * kSta; }
    __device__ __forceinline__ static double* tot(char* pv, int l) { return reinterpret_cast<double*>(chunk(pv, l) + kTot) + (l & 31); }
    __device__ __forceinline__ static uint32_t* edge(char* pv, int l) { return reinterpret_cast<uint32_t*>(chunk(pv, l) + kEdge) + (l & 31); }
    __device__ __forceinline__ static int own_base(char* pv, int l) { return __ldcg(reinterpret_cast<const int32_t*>(sta(pv, l)) + (A - 1)); }
};
static_assert(PvChunk<2>::kBytes == 896 && PvChunk<4>::kBytes == 1920, "PV chunk sizes");

}  // namespace
}  // namespace isla
