// ptb_rng.cuh -- device statement of the counter-based per-sample random stream.
//
// Replaces pt::rand_state (/root/reference/src/random_state.hpp:12-22,
// random_state.cpp:3-17: mt19937 seeded from random_device, one sequential stream
// per image row) with a stream keyed by (seed, slot, sample) so that any thread of
// any GPU can open the stream of any camera sample.  Definition: see oracle/ptb_rng.h
// (the CPU checker states the same generator; tests/test_rng.py compares the two
// bit for bit through ptb_rng_draws).
//
// One draw = 1 IMAD + ~7 ALU ops; the uniform is built by mantissa injection
// (0x3f800000 | r>>9) - 1, a 23-bit value exact in binary32 and binary64 alike.
#pragma once

#include "ptb_types.h"

namespace ptb {

struct Rng
{
    uint32_t state;
    uint32_t inc;
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

// Host side folds the seed once per launch; the device then needs one mix64 per sample.
__host__ __device__ __forceinline__ uint64_t seed_key(uint64_t seed)
{
    return mix64(seed + 0x9E3779B97F4A7C15ull);
}

__device__ __forceinline__ Rng rng_open(uint64_t key, uint32_t slot, uint32_t sample)
{
    uint64_t const h = mix64(key ^ ((static_cast<uint64_t>(slot) << 32) | static_cast<uint64_t>(sample)));
    Rng g;
    g.state = static_cast<uint32_t>(h);
    g.inc = static_cast<uint32_t>(h >> 32) | 1u;
    return g;
}

__device__ __forceinline__ uint32_t rng_next32(Rng& g)
{
    uint32_t const old = g.state;
    g.state = old * 747796405u + g.inc;
    uint32_t const word = ((old >> ((old >> 28) + 4u)) ^ old) * 277803737u;
    return (word >> 22) ^ word;
}

// [0,1) with 23 random bits
__device__ __forceinline__ float rng_uniform_f32(Rng& g)
{
    return __uint_as_float(0x3f800000u | (rng_next32(g) >> 9)) - 1.0f;
}

__device__ __forceinline__ double rng_uniform_f64(Rng& g)
{
    return static_cast<double>(rng_next32(g) >> 9) * (1.0 / 8388608.0);
}

} // namespace ptb
