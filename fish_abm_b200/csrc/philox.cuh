// Counter-based random integers for the device path (and the host-side fishabm_draw introspection call).
//
// The reference draws from six std::uniform_int_distribution<int> objects fed directly by a
// std::random_device (fish_mod.cpp:27-33).  Here every draw is a pure function of
// (seed, replicate, step, fish id, stream, k): Philox4x32-10 produces the 32-bit engine word and the
// mapping to [a,b] is the one libstdc++ applies to a full-range 32-bit engine (Lemire's multiply-shift
// with rejection of the low `2^32 mod range` residues), so a harness that serves the same words to the
// reference's own distributions gets the same integers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FISH_HD __host__ __device__ __forceinline__
#else
#define FISH_HD inline
#endif

namespace fishrng {

FISH_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// first output word of Philox4x32-10 for counter (c0,c1,c2,c3), key (k0,k1)
FISH_HD uint32_t philox4x32_10_word0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

FISH_HD uint32_t word(uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id, uint32_t stream, uint32_t k) {
    // counter words: fish id | step | stream (8 bits) + the low 24 bits of k | replicate ^ the high 8 bits of k (k is a
    // neighbour's id for tie draws: ids reach 2^28; identical to the plain packing for k < 2^24)
    return philox4x32_10_word0(id, step, (stream & 0xFFu) | (k << 8), replicate ^ (k & 0xFF000000u), (uint32_t)seed, (uint32_t)(seed >> 32));
}

// integer in [a,b]; consumes words k0, k0+1, ... until one passes the rejection test
FISH_HD int draw(uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id, uint32_t stream, uint32_t k0, int a, int b) {
    const uint32_t range = (uint32_t)(b - a) + 1u;
    const uint32_t reject_below = (0u - range) % range;
    for (uint32_t k = k0;; ++k) {
        const uint32_t u = word(seed, replicate, step, id, stream, k);
        const uint32_t lo = u * range;
        if (lo >= reject_below) return a + (int)mulhi32(u, range);
    }
}

// the same draw for a range whose rejection threshold 2^32 mod range is already known (the per-fish update: the model's
// error range is a launch constant; the modulo is an integer division on the device)
FISH_HD int draw_known(uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id, uint32_t stream, uint32_t k0, int a, uint32_t range,
                       uint32_t reject_below) {
    for (uint32_t k = k0;; ++k) {
        const uint32_t u = word(seed, replicate, step, id, stream, k);
        const uint32_t lo = u * range;
        if (lo >= reject_below) return a + (int)mulhi32(u, range);
    }
}

}  // namespace fishrng
