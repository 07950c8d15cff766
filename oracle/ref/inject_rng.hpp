// inject_rng.hpp -- force-included (g++ -include) ahead of the reference's own
// sources when building oracle/_ref/libptref_ctr.so.
//
// TEST INFRASTRUCTURE.  /root/reference/src/random_state.hpp is guarded by
// `#ifndef PT_RAND_STATE` (random_state.hpp:1-2).  Defining that macro here and
// supplying a `pt::rand_state` with the same three members the reference calls
// (default_with_seed / generate / generate_between, random_state.hpp:19-21)
// lets camera.cpp and main.cpp compile UNMODIFIED against the counter-based
// stream of oracle/ptb_rng.h, which is the stream the CUDA kernels reproduce.
// random_state.cpp is left out of that link.
#ifndef PT_RAND_STATE
#define PT_RAND_STATE

// What random_state.hpp:5-8 provided transitively (camera.cpp:5 uses std::tan
// with no <cmath> of its own).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <random>

#include "../ptb_rng.h"

namespace pt {

struct rand_state
{
    ptb_rng g{};

    [[nodiscard]] static auto default_with_seed(unsigned short seed) -> rand_state
    {
        rand_state r{};
        ptb_rng_key(&r.g, seed, 0u, 0u);
        return r;
    }

    // not part of the reference interface: position the stream on one sample
    auto key(std::uint64_t seed, std::uint32_t slot, std::uint32_t sample) -> void
    {
        ptb_rng_key(&g, seed, slot, sample);
    }

    [[nodiscard]] auto generate() -> double
    {
        return ptb_rng_uniform(&g);
    }

    // same expression as random_state.cpp:14-17
    [[nodiscard]] auto generate_between(double const min, double const max) -> double
    {
        return min + (max - min) * this->generate();
    }
};

} // namespace pt

#endif // !PT_RAND_STATE
