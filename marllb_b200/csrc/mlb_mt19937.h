// Host-side MT19937 (numpy legacy RandomState raw stream) used to build the
// replay tables the kernels draw from.  Own implementation of the published
// algorithm (Matsumoto & Nishimura 1998, init_genrand 2002 seeding) -- the
// generator behind np.random.RandomState(seed), which the reference uses at
// reservoir.py:45,76 and env.py:127.
#pragma once
#include <stdint.h>

namespace mlb {

class MT19937 {
public:
    explicit MT19937(uint32_t seed) { reseed(seed); }
    void reseed(uint32_t seed) {
        s_[0] = seed;
        for (int i = 1; i < N; i++) s_[i] = 1812433253u * (s_[i - 1] ^ (s_[i - 1] >> 30)) + (uint32_t)i;
        pos_ = N;
    }
    uint32_t next() {
        if (pos_ >= N) twist();
        uint32_t y = s_[pos_++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

private:
    static constexpr int N = 624, M = 397;
    static uint32_t mix(uint32_t hi, uint32_t lo) {
        const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
        return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    void twist() {
        for (int k = 0; k < N; k++) s_[k] = s_[(k + M) % N] ^ mix(s_[k], s_[(k + 1) % N]);
        pos_ = 0;
    }
    uint32_t s_[N];
    int pos_;
};

}  // namespace mlb
