// stdrng.cpp — restatement of rand 0.8.5's StdRng as the reference uses it.
//
// The reference draws its maze from `StdRng::seed_from_u64(0)` (reference src/main.rs:381), shuffles the
// edge list with it (:382) and keeps drawing `gen::<f32>()` for mirror / light decisions (:460,467,494,501).
// rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4 (reference Cargo.lock:342-371) are NOT vendored under
// /root/reference, so this file restates their published algorithms:
//   * StdRng = ChaCha12, 256-bit key = seed, 64-bit block counter starting at 0, 64-bit stream id 0,
//     output consumed as consecutive little-endian u32 words (BlockRng over a 4-block buffer — buffering
//     does not change the word order);
//   * SeedableRng::seed_from_u64: PCG32 (XSH-RR) expander, one u32 per 4 seed bytes;
//   * Standard f32: (next_u32 >> 8) * 2^-24;
//   * UniformInt<u32>::sample_single: widening-multiply rejection with zone = (range << lz(range)) - 1.
// PARITY UNPINNED against a real `cargo run` (no Rust toolchain here, SURVEY §8 c).  The ChaCha core is
// pinned by the RFC 7539 / eSTREAM known-answer vectors in tests/test_stdrng.py.
#include "host_surface.h"

namespace mmh {

static inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }

static inline void quarter(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
    a += b; d ^= a; d = rotl(d, 16);
    c += d; b ^= c; b = rotl(b, 12);
    a += b; d ^= a; d = rotl(d, 8);
    c += d; b ^= c; b = rotl(b, 7);
}

void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
    for (int i = 0; i < 16; i++) x[i] = s[i];
    for (int r = 0; r < rounds; r += 2) {
        quarter(x[0], x[4], x[8], x[12]);
        quarter(x[1], x[5], x[9], x[13]);
        quarter(x[2], x[6], x[10], x[14]);
        quarter(x[3], x[7], x[11], x[15]);
        quarter(x[0], x[5], x[10], x[15]);
        quarter(x[1], x[6], x[11], x[12]);
        quarter(x[2], x[7], x[8], x[13]);
        quarter(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

StdRng::StdRng(uint64_t seed) {
    // rand_core 0.6.4 SeedableRng::seed_from_u64
    const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
    uint64_t state = seed;
    for (int i = 0; i < 8; i++) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        key_[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));   // bytes LE -> word LE: identity
    }
    counter_ = 0;
    pos_ = 16;
}

uint32_t StdRng::next_u32() {
    if (pos_ >= 16) {
        chacha_block(key_, counter_++, 0, 12, buf_);
        pos_ = 0;
    }
    return buf_[pos_++];
}

float StdRng::gen_f32() {
    return (float)(next_u32() >> 8) * (1.0f / 16777216.0f);
}

uint32_t StdRng::gen_range_u32(uint32_t low, uint32_t high) {
    uint32_t range = high - low;
    if (range == 0) return next_u32();
    uint32_t zone = (range << __builtin_clz(range)) - 1u;
    for (;;) {
        uint32_t v = next_u32();
        uint64_t m = (uint64_t)v * (uint64_t)range;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo <= zone) return low + hi;
    }
}

}  // namespace mmh
