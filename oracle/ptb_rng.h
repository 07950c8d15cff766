/* ptb_rng.h -- the counter-based per-sample random stream, CPU statement.
 *
 * TEST INFRASTRUCTURE (oracle side).  The product (CUDA) re-states the same
 * generator in cpu-path-tracing_b200/csrc/ptb_rng.cuh; the two must agree bit
 * for bit, which tests/test_rng.py checks through the C ABI.
 *
 * Replaces (by design) the reference's per-row
 * std::mt19937 + uniform_real_distribution<double> stream
 * (/root/reference/src/random_state.hpp:12-22, random_state.cpp:3-17), which is
 * seeded from std::random_device and therefore not reproducible, and which is
 * sequential per image row and therefore not shardable over GPU threads.
 *
 * Stream definition
 *   key     = (seed:u64, slot:u32, sample:u32)
 *             slot   = ((y*W + x)*nsub + sy)*nsub + sx   (reference loop
 *                      coordinates of src/main.cpp:226-232, NOT the flipped row)
 *             sample = absolute sample index inside the sub-pixel
 *                      (the `s` of src/main.cpp:184)
 *   h       = mix64( mix64(seed + 0x9E3779B97F4A7C15) ^ ((u64)slot << 32 | sample) )
 *             mix64 = SplitMix64 finaliser
 *   state   = low 32 bits of h,  inc = (high 32 bits of h) | 1
 *   draw    : PCG-RXS-M-XS-32 on a 32-bit LCG with per-stream increment
 *               old   = state;  state = old * 747796405 + inc
 *               word  = ((old >> ((old >> 28) + 4)) ^ old) * 277803737
 *               r     = (word >> 22) ^ word
 *   uniform = (r >> 9) * 2^-23      in [0, 1), 23 bits: exactly representable
 *             in binary32 AND binary64, so the FP32 render, the FP64 parity
 *             render and this oracle all see the identical value.
 * Draws are consumed in the order of SURVEY.md section 3 "RNG draw order".
 */
#ifndef PTB_RNG_H
#define PTB_RNG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptb_rng
{
    uint32_t state;
    uint32_t inc;
    uint64_t draws; /* statistics only */
} ptb_rng;

static inline uint64_t ptb_mix64(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

static inline void ptb_rng_key(ptb_rng* g, uint64_t seed, uint32_t slot, uint32_t sample)
{
    uint64_t const k = ptb_mix64(seed + 0x9E3779B97F4A7C15ull);
    uint64_t const h = ptb_mix64(k ^ (((uint64_t)slot << 32) | (uint64_t)sample));
    g->state = (uint32_t)h;
    g->inc = (uint32_t)(h >> 32) | 1u;
}

static inline uint32_t ptb_rng_next32(ptb_rng* g)
{
    uint32_t const old = g->state;
    g->state = old * 747796405u + g->inc;
    uint32_t const word = ((old >> ((old >> 28) + 4u)) ^ old) * 277803737u;
    g->draws++;
    return (word >> 22) ^ word;
}

static inline double ptb_rng_uniform(ptb_rng* g)
{
    return (double)(ptb_rng_next32(g) >> 9) * (1.0 / 8388608.0);
}

#ifdef __cplusplus
}
#endif

#endif /* PTB_RNG_H */
