// Host-side Fiat-Shamir transcript: DuplexChallenger<BabyBear, Poseidon<_, CosetMds<_,16>, 16, 5>, 16>
// (basic/src/bin/valida.rs:360-365,382,397).  The transcript is strictly sequential and tiny
// (a few hundred permutations per proof), so it stays on the host and crosses to the device only
// as 8-word digests and 5-limb challenges.  Values are Montgomery words.
#pragma once
#include <cstdint>
#include <vector>
#include "../bb.cuh"

namespace vgh {

struct Poseidon16 {
    uint32_t rc[480];       // Montgomery
    uint32_t mds[16][16];   // Montgomery, out[i] = sum_j mds[i][j] in[j]
    void set(const uint32_t rc_canonical[480], const uint32_t* mds_canonical_or_null);
    void permute(uint32_t s[16]) const;
};

struct Challenger {
    const Poseidon16* perm = nullptr;
    uint32_t state[16] = {0};
    std::vector<uint32_t> input, output;   // Montgomery words
    void duplexing();
    void observe(uint32_t v_monty);
    void observe_digest_canonical(const uint32_t d[8]);
    uint32_t sample();                       // Montgomery
    bb::E5 sample_ext();
    uint32_t sample_bits(int bits);
    bool check_witness(int bits, uint32_t witness_monty);
    uint32_t grind(int bits);                // smallest canonical witness, returned in Montgomery form
};

}  // namespace vgh
