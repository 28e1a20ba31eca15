// Counter-based dropout RNG contract, device side.
// Philox4x32-10; key = (seed_lo, seed_hi); counter = (site, t, b_global, chunk).
// p = 0.5 "bit sites": channel c -> chunk c/128, word (c%128)/32, bit c%32; keep iff bit == 1; scale 2.
// p = 0.1 "word sites": channel c -> chunk c/4, word c%4; drop iff word < floor(p * 2^32).
// The oracle's integer restatement of the same contract is oracle/philox.py (pinned by the Random123
// known-answer vectors in tests/test_philox.py and, on the GPU, tests/test_gpu_kernels.py).
#pragma once
#include "common.cuh"

namespace tts {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

enum DropSite : uint32_t {
    SITE_DEC_PRENET_FC1 = 0,
    SITE_DEC_PRENET_FC2 = 1,
    SITE_ENC_PRENET_CONV0 = 2,
    SITE_POSTNET_CONV0 = 5,
    SITE_ENC_PE = 16,
    SITE_DEC_PE = 17,
    SITE_ENC_LAYER0 = 32,
    SITE_DEC_LAYER0 = 64,
};

TTS_HD uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) { k0 += kPhiloxW0; k1 += kPhiloxW1; }
        uint64_t p0 = (uint64_t)kPhiloxM0 * c.x;
        uint64_t p1 = (uint64_t)kPhiloxM1 * c.z;
        uint4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
    }
    return c;
}

// keep-bit of channel `c` at a p = 0.5 site
TTS_HD bool keep_bit(uint64_t seed, uint32_t site, uint32_t t, uint32_t b, uint32_t c) {
    uint4 w = philox4x32_10(make_uint4(site, t, b, c >> 7), (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t word = (c >> 5) & 3u;
    uint32_t v = word == 0 ? w.x : word == 1 ? w.y : word == 2 ? w.z : w.w;
    return (v >> (c & 31u)) & 1u;
}

}  // namespace tts
