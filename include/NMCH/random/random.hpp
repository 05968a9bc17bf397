// NMCH/random/random.hpp -- generator tags of the method templates.
//
// The reference's classes are templates over a cuRAND state type (include/NMCH/methods/NMCH.hpp:28,
// explicit instantiations src/NMCH/methods/NMCH.cu:30-32).  In this engine the tag only SELECTS the stream
// mode of the fused kernels (no per-thread cuRAND state object exists for the native mode), so the tags are
// forward declarations: user code that names curandStateXORWOW_t etc. compiles with plain g++, and the
// declarations coexist with <curand_kernel.h> when the user includes that too.
//
//   curandStateXORWOW_t        -> cuRAND-XORWOW-compatible stream (draw-for-draw validation mode)
//   curandStatePhilox4_32_10_t -> native counter-based Philox4x32-10 (the product path)
//   curandStateMRG32k3a_t      -> cuRAND-MRG32k3a-compatible stream (validation mode, like XORWOW)
#ifndef NMCH_RANDOM_HPP
#define NMCH_RANDOM_HPP

struct curandStateXORWOW;
struct curandStateMRG32k3a;
struct curandStatePhilox4_32_10;
typedef struct curandStateXORWOW curandStateXORWOW_t;
typedef struct curandStateMRG32k3a curandStateMRG32k3a_t;
typedef struct curandStatePhilox4_32_10 curandStatePhilox4_32_10_t;

namespace nmch::random {

enum class stream_mode { native_philox, xorwow_compat, philox_compat, mrg32k3a_compat };

template <typename rnd_state> struct tag_traits;
template <> struct tag_traits<curandStateXORWOW_t> { static constexpr stream_mode mode = stream_mode::xorwow_compat; static constexpr bool alias = false; };
template <> struct tag_traits<curandStatePhilox4_32_10_t> { static constexpr stream_mode mode = stream_mode::native_philox; static constexpr bool alias = false; };
template <> struct tag_traits<curandStateMRG32k3a_t> { static constexpr stream_mode mode = stream_mode::mrg32k3a_compat; static constexpr bool alias = false; };

}  // namespace nmch::random

#endif  // NMCH_RANDOM_HPP
