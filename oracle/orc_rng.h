/* oracle/orc_rng.h -- TEST INFRASTRUCTURE (parity oracle), not product code.
 *
 * Counter-based random integers for the oracle side:
 *   - Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11),
 *     restated from the published round function; pinned by the Random123 known-answer
 *     vectors in tests/test_oracle_rng.py.
 *   - the integer mapping of libstdc++'s uniform_int_distribution<int>(a,b) when the URBG
 *     yields full 32-bit words (Lemire's nearly-divisionless method,
 *     /usr/include/c++/13/bits/uniform_int_dist.h:257-279,323-328).  This is what the
 *     reference's six distributions (fish_mod.cpp:28-33) do with each word of `rd`.
 *
 * Every draw is addressed by (seed, replicate, step, fish id, stream, k): k counts the words
 * consumed by one logical draw (k>0 only after a Lemire rejection) or, for the tie-break
 * stream, identifies the neighbour the tie is with.
 */
#ifndef ORC_RNG_H
#define ORC_RNG_H
#include <stdint.h>

enum {
    ORC_STREAM_CHOICE = 0,     /* distrib_choice(0,1)       fish_mod.cpp:28 */
    ORC_STREAM_ERROR = 1,      /* distrib_error(-10,10)     fish_mod.cpp:29 */
    ORC_STREAM_RANDOMWALK = 2, /* distrib_randomwalk(-45,45) fish_mod.cpp:30 */
    ORC_STREAM_QUALITY = 3,    /* distrib_quality(0,10)     fish_mod.cpp:31 */
    ORC_STREAM_SEED = 4,       /* distrib_seed(0,0)         fish_mod.cpp:32 */
    ORC_STREAM_SAMPLE = 5,     /* distrib_sample(0,35)      fish_mod.cpp:33 */
    ORC_STREAM_INIT = 6        /* initial positions (rand()%201 in fish_mod.cpp:199,202) */
};

static inline void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* one 32-bit word for (seed, replicate, step, id, stream, k) */
static inline uint32_t orc_word(uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id,
                                uint32_t stream, uint32_t k) {
    /* low 24 bits of k beside the stream, its high 8 bits folded into the replicate word (k = neighbour id for ties) */
    uint32_t ctr[4] = {id, step, (stream & 0xFFu) | (k << 8), replicate ^ (k & 0xFF000000u)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[0];
}

/* Lemire acceptance test for one word: returns 1 and the mapped value, or 0 (= libstdc++ redraws). */
static inline int orc_lemire_accept(uint32_t word, int a, int b, int* value) {
    uint32_t range = (uint32_t)(b - a) + 1u; /* all reference ranges are tiny; range==0 not needed */
    uint64_t product = (uint64_t)word * range;
    uint32_t low = (uint32_t)product;
    uint32_t threshold = (uint32_t)(0u - range) % range; /* 2^32 mod range */
    if (low < threshold) return 0;
    *value = a + (int)(product >> 32);
    return 1;
}

/* the logical draw: words k0, k0+1, ... until one is accepted */
static inline int orc_draw_from(uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id,
                                uint32_t stream, uint32_t k0, int a, int b) {
    int v = a;
    for (uint32_t k = k0;; ++k)
        if (orc_lemire_accept(orc_word(seed, replicate, step, id, stream, k), a, b, &v)) return v;
}

static inline int orc_draw(uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id,
                           uint32_t stream, int a, int b) {
    return orc_draw_from(seed, replicate, step, id, stream, 0, a, b);
}

#endif
