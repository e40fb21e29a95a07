// Philox4x32-10 counter-based RNG and the stream conventions of the engine.
//
// Replaces the reference's host RNGs on the hot path: numpy's global generator behind
// np.random.choice (islaMcts/agents/mcts.py:57, utils/action_selection_functions.py:24) and the
// gym action_space generator behind env.action_space.sample()
// (utils/action_selection_functions.py:47).  Counter layout (c0,c1,c2,c3), key = 64-bit seed:
//   tree stream    (blk, 0,   tree, 0x80000000)  selection / expansion / tie-break draws of one tree
//   rollout stream (blk, sim, tree, j)           rollout j of simulation `sim` (cumulative index)
//   reset stream   (0,   0,   idx,  0xC0000000)  synthetic root states
// One 32-bit word per draw; discrete draws use the multiply-high map u*n>>32, except env.action_space.sample()
// on a Discrete(2^k) space, which takes k bits (most significant first) from a 32-bit word of the stream.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace isla {

constexpr uint32_t kStreamTree = 0x80000000u;
constexpr uint32_t kStreamReset = 0xC0000000u;

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Buffered sequential stream: 4 words per Philox call.
struct Stream {
    uint32_t k0, k1, c1, c2, c3, blk;
    uint32_t buf[4];
    int idx;
    uint32_t bitbuf;  // reservoir of next_bits()
    int bits_left;
    __device__ __forceinline__ void init(uint64_t seed, uint32_t c1_, uint32_t c2_, uint32_t c3_) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); c1 = c1_; c2 = c2_; c3 = c3_; blk = 0; idx = 4;
        bitbuf = 0u; bits_left = 0;
    }
    // resume a stream that has already consumed `consumed` words
    __device__ __forceinline__ void resume(uint64_t seed, uint32_t c1_, uint32_t c2_, uint32_t c3_, uint32_t consumed) {
        init(seed, c1_, c2_, c3_);
        blk = consumed >> 2;
        const int r = consumed & 3;
        if (r) { philox4x32_10(blk, c1, c2, c3, k0, k1, buf); ++blk; idx = r; }
    }
    __device__ __forceinline__ uint32_t consumed() const { return blk * 4u - (uint32_t)(4 - idx); }
    __device__ __forceinline__ uint32_t next() {
        if (idx == 4) { philox4x32_10(blk++, c1, c2, c3, k0, k1, buf); idx = 0; }
        // idx is 0..3: select without dynamic register indexing
        const uint32_t v = idx == 0 ? buf[0] : idx == 1 ? buf[1] : idx == 2 ? buf[2] : buf[3];
        ++idx;
        return v;
    }
    // k bits (k divides 32), most significant first; a fresh word is drawn when the reservoir is empty
    __device__ __forceinline__ uint32_t next_bits(int k) {
        if (bits_left == 0) { bitbuf = next(); bits_left = 32; }
        const uint32_t v = bitbuf >> (32 - k);
        bitbuf <<= k;
        bits_left -= k;
        return v;
    }
};

__host__ __device__ __forceinline__ uint32_t mulhi_u32(uint32_t u, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __umulhi(u, n);
#else
    return (uint32_t)(((uint64_t)u * n) >> 32);
#endif
}
__host__ __device__ __forceinline__ float u32_to_f01(uint32_t u) { return (float)(u >> 8) * 5.9604644775390625e-08f; }
__host__ __device__ __forceinline__ double u32_to_d01(uint32_t u) { return (double)u * 2.3283064365386963e-10; }

// gym Box.sample() for a bounded float32 box: low + u*(high-low), every step rounded to float32
__device__ __forceinline__ float box_sample_f32(uint32_t u, float low, float high) {
    return __fadd_rn(low, __fmul_rn(u32_to_f01(u), __fsub_rn(high, low)));
}

}  // namespace isla
