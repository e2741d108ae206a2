// oracle/lemire_check.cpp -- TEST INFRASTRUCTURE.  Feeds the oracle's Philox words to the REAL
// std::uniform_int_distribution<int> of this toolchain's libstdc++ (the mapping the reference relies on,
// fish_mod.cpp:27-33) and prints the integers, so tests can pin orc_rng.h's restated mapping against it.
// usage: lemire_check seed count   -> lines "stream step id a b value words_consumed"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "orc_rng.h"

struct WordEngine {
    typedef uint32_t result_type;
    uint64_t seed; uint32_t step, id, stream, k;
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xFFFFFFFFu; }
    result_type operator()() { return orc_word(seed, 0, step, id, stream, k++); }
};

int main(int argc, char** argv) {
    uint64_t seed = argc > 1 ? strtoull(argv[1], 0, 0) : 1;
    int count = argc > 2 ? atoi(argv[2]) : 1000;
    const int ranges[6][3] = {{0, 0, 1}, {1, -10, 10}, {2, -45, 45}, {3, 0, 10}, {5, 0, 35}, {6, 0, 200}};
    for (int t = 0; t < count; ++t)
        for (int r = 0; r < 6; ++r) {
            WordEngine e{seed, (uint32_t)(t * 7 + 1), (uint32_t)(t * 13 + r), (uint32_t)ranges[r][0], 0};
            std::uniform_int_distribution<int> d(ranges[r][1], ranges[r][2]);
            int v = d(e);
            printf("%d %u %u %d %d %d %u\n", ranges[r][0], e.step, e.id, ranges[r][1], ranges[r][2], v, e.k);
        }
    // crafted first words around the rejection threshold (2^32 mod range): "99 word 0 a b value words_consumed"
    struct Crafted {
        typedef uint32_t result_type;
        uint32_t first; uint32_t k;
        static constexpr result_type min() { return 0u; }
        static constexpr result_type max() { return 0xFFFFFFFFu; }
        result_type operator()() { return k++ == 0 ? first : 0x9E3779B9u * k; }
    };
    for (int r = 0; r < 6; ++r) {
        const uint32_t range = (uint32_t)(ranges[r][2] - ranges[r][1]) + 1u;
        for (uint32_t w = 0; w < 400; ++w) {
            const uint32_t firsts[3] = {w, 0xFFFFFFFFu - w, (uint32_t)(((uint64_t)w << 32) / range) + (w & 1)};
            for (int f = 0; f < 3; ++f) {
                Crafted e{firsts[f], 0};
                std::uniform_int_distribution<int> d(ranges[r][1], ranges[r][2]);
                int v = d(e);
                printf("99 %u 0 %d %d %d %u\n", firsts[f], ranges[r][1], ranges[r][2], v, e.k);
            }
        }
    }
    return 0;
}
