// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// Counter-based uniform streams.  The reference draws from SFMT-19937 seeded from /dev/urandom
// (src/libcore/random.cpp:473-489, nextFloat :632-641) and is therefore not reproducible; the
// parity contract (BASELINE.json north_star) is "identical uniform streams => identical
// decisions".  Both the oracle and the CUDA path address uniforms by KEY instead of by draw
// order, using Philox4x32-10 (Salmon et al. 2011, public algorithm):
//
//   word(stream, a, b, j) = philox4x32_10(ctr = {lo32(a), hi32(a), b, (stream << 24) | (j >> 2)},
//                                          key = {lo32(seed), hi32(seed)})[j & 3]
//   uniform = (word >> 8) * 2^-24          (24 random mantissa bits, in [0,1))
//
// Streams (DESIGN.md "uniform address space"):
//   1 BOOT     a = bootstrap sample index, b = sampler (0 sensor, 1 emitter, 2 direct), j = coordinate
//   2 RESAMPLE a = chain id, b = 0, j = 0           (seedPDF.sample(next1D()), pathsampler.cpp:951-954)
//   3 COIN     a = chain id, b = mutation index, j = 0 large-step, 1 accept-1, 2 accept-2, 3 mixture
//   4+s STAGE1 a = chain id, b = mutation index, j = 2*coordinate + draw, s = sampler
//   7+s STAGE2 same for the second-stage proposal
#pragma once
#include <cstdint>

namespace orc {

struct Philox {
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t) 0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t) 0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c[1] ^ k[0];
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c[3] ^ k[1];
        uint32_t n3 = (uint32_t) p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static inline void gen(uint32_t c[4], uint64_t seed) {
        uint32_t k[2] = { (uint32_t) seed, (uint32_t) (seed >> 32) };
        for (int i = 0; i < 10; ++i) {
            round(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
    }
};

enum Stream { S_BOOT = 1, S_RESAMPLE = 2, S_COIN = 3, S_STAGE1 = 4, S_STAGE2 = 7, S_DIRECT = 10 };

inline float keyedUniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b, uint32_t j) {
    uint32_t c[4] = { (uint32_t) a, (uint32_t) (a >> 32), b, (stream << 24) | (j >> 2) };
    Philox::gen(c, seed);
    return (float) (c[j & 3] >> 8) * (1.0f / 16777216.0f);
}

} // namespace orc
