"""The integer identities the tensor form of the matching kernel rests on (boslam_b200/csrc/bfm_tensor.cuh), restated in
numpy and checked exhaustively or on random data - no GPU needed.  The kernel itself is compared with the oracle, the POPC
kernel and live cv2 by the `-m gpu` tests; these pin WHY it can be exact:

* the s8 expansion: bit b -> 1 - 2b, so the dot product of two expanded descriptors is 256 - 2 * hamming;
* the 16-bit key of the epilogue, `16384 - 64 * dot + column` = `distance << 7 | column`, and the two IMADs that make two
  keys in one 32-bit register modulo 2^32 (`fold_chunk`);
* the two smallest of a stream of keys kept with (min, max) of a pair + a three-input minimum;
* the swizzled plane layout: chunk c of row r at chunk position c ^ (r % 8)."""
import numpy as np

from oracle import hamming_oracle as orc

U32 = np.uint64(1 << 32)


def _expand(desc):
    bits = np.unpackbits(desc, axis=1, bitorder="little").astype(np.int32)
    return 1 - 2 * bits


def test_dot_product_of_expanded_descriptors_is_256_minus_twice_the_hamming_distance():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    t[:10] = q[:10]                       # distance 0
    t[10] = ~q[10]                        # distance 256
    dot = _expand(q) @ _expand(t).T
    idx, dist = orc.knn(q, t, k=t.shape[0])          # all distances, sorted per row
    ham = (256 - dot) // 2
    assert ((256 - dot) % 2 == 0).all() and ham.min() == 0 and ham.max() == 256
    assert np.array_equal(np.sort(ham, axis=1), dist)


def test_sixteen_bit_keys_hold_distance_and_column_and_pack_two_to_a_register():
    dots = np.arange(-256, 257, 2, dtype=np.int64)                # every value a dot product can take
    cols = np.arange(128, dtype=np.int64)
    key = 16384 - 64 * dots[:, None] + cols[None, :]
    assert key.min() == 0 and key.max() == (256 << 7 | 127) < 0xFFFF      # all-ones stays free for "no key"
    assert np.array_equal(key >> 7, np.broadcast_to(((256 - dots) // 2)[:, None], key.shape))
    assert np.array_equal(key & 127, np.broadcast_to(cols[None, :], key.shape))
    # fold_chunk: P = v[j + 16] * m_hi + (v[j] * m_lo + cst) modulo 2^32, v = the s32 accumulators (two's complement)
    rng = np.random.default_rng(1)
    m_lo, m_hi = np.uint64((-64) % (1 << 32)), np.uint64(((-64) << 16) % (1 << 32))
    for ch in range(4):
        for j in range(16):
            c_lo, c_hi = ch * 32 + j, ch * 32 + j + 16
            cst = np.uint64((16384 + c_lo) + ((16384 + c_hi) << 16))
            d_lo, d_hi = rng.choice(dots, 200), rng.choice(dots, 200)
            v_lo, v_hi = (d_lo % (1 << 32)).astype(np.uint64), (d_hi % (1 << 32)).astype(np.uint64)
            P = (v_hi * m_hi % U32 + (v_lo * m_lo % U32 + cst) % U32) % U32
            assert np.array_equal(P & np.uint64(0xFFFF), (16384 - 64 * d_lo + c_lo).astype(np.uint64))
            assert np.array_equal(P >> np.uint64(16), (16384 - 64 * d_hi + c_hi).astype(np.uint64))
    # garbage in the high column (a column past the train range) never reaches the low key
    junk = rng.integers(0, 1 << 32, 200, dtype=np.uint64)
    P = (junk * m_hi % U32 + ((np.uint64(7) * m_lo) % U32 + np.uint64(16384 + 5 + ((16384 + 21) << 16))) % U32) % U32
    assert (P & np.uint64(0xFFFF) == np.uint64(16384 - 64 * 7 + 5)).all()


def test_two_smallest_through_pair_ordering_and_a_three_input_minimum():
    rng = np.random.default_rng(2)
    for _ in range(200):
        n = int(rng.integers(2, 65)) * 2
        keys = rng.choice(1 << 16, n, replace=rng.random() < 0.3).astype(np.int64)
        p1 = p2 = 0xFFFF
        for a, b in keys.reshape(-1, 2):
            lo, hi = min(a, b), max(a, b)
            p2 = min(p2, max(p1, lo), hi)
            p1 = min(p1, lo)
        s = np.sort(keys)
        assert (p1, p2) == (s[0], s[1])
    # the row-state commit: two contributors (best, second) merged with one displaced value
    for _ in range(200):
        a, b = np.sort(rng.choice(1 << 20, 2, replace=False)), np.sort(rng.choice(1 << 20, 2, replace=False))
        best, second = 0xFFFFFFFF, 0xFFFFFFFF
        for mb, ms in (a, b):
            displaced, best = best, min(best, mb)                 # atomicMin returns the old value
            second = min(second, min(max(displaced, mb), ms))
        s = np.sort(np.concatenate([a, b]))
        assert (best, second) == (s[0], s[1])


def test_swizzled_plane_layout_is_a_permutation_of_every_1024_byte_atom():
    rows = np.arange(64)
    for r in rows:
        pos = [(c ^ (r % 8)) for c in range(8)]
        assert sorted(pos) == list(range(8))                      # a row's eight 16-byte chunks stay inside its 128 bytes
    # eight consecutive rows put chunk c on eight different positions: a column of 16-byte chunks read by the MMA never
    # hits one bank group twice
    for c in range(8):
        assert sorted((c ^ (r % 8)) for r in range(8)) == list(range(8))
