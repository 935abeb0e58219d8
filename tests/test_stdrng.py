"""rand 0.8.5 StdRng restatement (reference call sites src/main.rs:381-382,460,467,494,501; crates pinned by
reference Cargo.lock:342-371 but not vendored).  The ChaCha core is pinned with published known-answer vectors
(RFC 7539 §2.3.2 ChaCha20 block; zero-key ChaCha20/12/8 keystreams of the eSTREAM test-vector set); the seed
expander and samplers are cross-checked against the independent Python restatement in oracle/host_ref.py.
PARITY UNPINNED against a real `cargo run`: no Rust toolchain or crate source exists here."""
import numpy as np


def _bytes(words):
    return np.asarray(words, dtype="<u4").tobytes().hex()


def test_chacha20_rfc7539_block(mm):
    # RFC 7539 §2.3.2: key 00..1f, counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00.  With rand_chacha's word
    # layout (64-bit counter in words 12-13, 64-bit stream in 14-15) that state is counter = 1 | 0x09000000<<32,
    # stream = 0x4a000000.
    key = bytes(range(32))
    out = mm.chacha_block(key, 1 | (0x09000000 << 32), 0x4A000000, 20)
    assert _bytes(out) == ("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                           "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")


def test_chacha_zero_key_keystreams(mm):
    z = bytes(32)
    assert _bytes(mm.chacha_block(z, 0, 0, 20)) == ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                                                    "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    assert _bytes(mm.chacha_block(z, 0, 0, 12)) == ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                                                    "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
    assert _bytes(mm.chacha_block(z, 0, 0, 8)) == ("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
                                                   "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42")


def test_cpp_matches_python_restatement(mm):
    from oracle import host_ref

    for seed in (0, 1, 0xDEADBEEF, 2 ** 64 - 1):
        a, b = mm.StdRng(seed), host_ref.StdRng(seed)
        assert [a.next_u32() for _ in range(200)] == [b.next_u32() for _ in range(200)]
        assert [a.gen_range(0, n) for n in (1, 2, 3, 7, 180, 2 ** 31 + 5, 2 ** 32 - 1)] == \
               [b.gen_range(0, n) for n in (1, 2, 3, 7, 180, 2 ** 31 + 5, 2 ** 32 - 1)]
        assert [a.gen_f32() for _ in range(50)] == [float(b.gen_f32()) for _ in range(50)]


def test_block_counter_advances_every_16_words(mm):
    from oracle import host_ref

    r = mm.StdRng(0)
    words = [r.next_u32() for _ in range(48)]
    ref = host_ref.StdRng(0)
    blocks = [host_ref.chacha_block(ref.key, c, 0, 12) for c in range(3)]
    assert words == blocks[0] + blocks[1] + blocks[2]


def test_samplers(mm):
    r = mm.StdRng(7)
    f = [r.gen_f32() for _ in range(2000)]
    assert all(0.0 <= v < 1.0 for v in f) and 0.45 < float(np.mean(f)) < 0.55
    assert all(float(v * 2 ** 24).is_integer() for v in f)          # (u32 >> 8) * 2^-24
    g = [r.gen_range(3, 10) for _ in range(2000)]
    assert min(g) == 3 and max(g) == 9
