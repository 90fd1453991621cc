"""CPU tests pinning the oracle's primitives against published known answers (no GPU)."""
import hashlib

import numpy as np

P = 2013265921


def test_keccak_kats(oracle):
    # Keccak-256 (pre-SHA3 padding) known answers
    assert oracle.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert oracle.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"


def test_keccak_permutation_matches_sha3(oracle):
    # same permutation + sponge with the SHA-3 domain byte must equal hashlib for every length
    rng = np.random.default_rng(1)
    for n in list(range(0, 300)) + [135, 136, 137, 271, 272, 273, 1000, 4096]:
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert oracle.keccak256(data, pad=0x06) == hashlib.sha3_256(data).digest(), n


def test_babybear_constants(oracle):
    assert P == 2**31 - 2**27 + 1
    g27 = oracle.L.orc_two_adic_generator(27)
    assert g27 == 0x1A427A41 == pow(31, 15, P)
    assert pow(g27, 1 << 27, P) == 1 and pow(g27, 1 << 26, P) == P - 1
    for bits in range(0, 28):
        g = oracle.L.orc_two_adic_generator(bits)
        assert pow(g, 1 << bits, P) == 1
        if bits:
            assert pow(g, 1 << (bits - 1), P) == P - 1
    assert (oracle.L.orc_inv(12345) * 12345) % P == 1


def test_ext5_field_axioms(oracle):
    rng = np.random.default_rng(2)
    one = np.array([1, 0, 0, 0, 0], dtype=np.uint32)
    x = np.array([0, 1, 0, 0, 0], dtype=np.uint32)
    # X^5 = 2
    acc = one
    for _ in range(5):
        acc = oracle.ext_mul(acc, x)
    assert list(acc) == [2, 0, 0, 0, 0]
    for _ in range(50):
        a = rng.integers(0, P, size=5, dtype=np.uint32)
        b = rng.integers(0, P, size=5, dtype=np.uint32)
        c = rng.integers(0, P, size=5, dtype=np.uint32)
        assert list(oracle.ext_mul(a, b)) == list(oracle.ext_mul(b, a))
        assert list(oracle.ext_mul(oracle.ext_mul(a, b), c)) == list(oracle.ext_mul(a, oracle.ext_mul(b, c)))
        assert list(oracle.ext_mul(a, oracle.ext_inv(a))) == list(one)


def test_dft_matches_textbook_definition(oracle):
    rng = np.random.default_rng(3)
    for log_h in range(0, 8):
        m = rng.integers(0, P, size=(1 << log_h, 3), dtype=np.uint32)
        fast = oracle.dft(m)
        assert np.array_equal(fast, oracle.naive_dft(m))
        assert np.array_equal(oracle.dft(fast, inverse=True), m)
    # independent pure-python check of one small case
    m = rng.integers(0, P, size=(8, 1), dtype=np.uint32)
    w = oracle.L.orc_two_adic_generator(3)
    ref = [sum(int(m[j, 0]) * pow(w, j * k, P) for j in range(8)) % P for k in range(8)]
    assert [int(v) for v in oracle.dft(m)[:, 0]] == ref


def test_coset_lde_is_polynomial_evaluation(oracle):
    # LDE on shift*K must equal Horner evaluation of the interpolated polynomial
    rng = np.random.default_rng(4)
    h, shift = 8, 31
    m = rng.integers(0, P, size=(h, 2), dtype=np.uint32)
    coeffs = oracle.dft(m, inverse=True)
    lde = oracle.coset_lde(m, 1, shift, bitrev=False)
    w16 = oracle.L.orc_two_adic_generator(4)
    for i in range(16):
        x = shift * pow(w16, i, P) % P
        for c in range(2):
            val = 0
            for k in reversed(range(h)):
                val = (val * x + int(coeffs[k, c])) % P
            assert val == int(lde[i, c])
    # committed order = bit-reversed rows
    br = oracle.coset_lde(m, 1, shift, bitrev=True)
    rev = [int(format(i, "04b")[::-1], 2) for i in range(16)]
    assert np.array_equal(br, lde[rev])


def test_merkle_root_small_tree_by_hand(oracle):
    # two matrices of heights 4 and 2: root = C( C(C(l0,l1), h(m2 row0)), C(C(l2,l3), h(m2 row1)) )
    rng = np.random.default_rng(5)
    a = rng.integers(0, P, size=(4, 3), dtype=np.uint32)
    b = rng.integers(0, P, size=(2, 2), dtype=np.uint32)

    def H(words):
        d = oracle.keccak256(np.asarray(words, dtype="<u4").tobytes())
        return [int.from_bytes(d[4 * i:4 * i + 4], "little") % P for i in range(8)]

    leaves = [H(a[i]) for i in range(4)]
    n0 = H(H(leaves[0] + leaves[1]) + H(b[0]))
    n1 = H(H(leaves[2] + leaves[3]) + H(b[1]))
    root = H(n0 + n1)
    assert [int(v) for v in oracle.merkle_root([a, b])] == root
    assert [int(v) for v in oracle.merkle_root([b, a])] == root  # order by height, not by position


def test_poseidon_structure(oracle):
    mds = oracle.coset_mds().astype(object)
    # the MDS layer is linear and invertible: check non-singularity mod p by Gaussian elimination
    M = [[int(v) for v in row] for row in mds]
    n = 16
    for col in range(n):
        piv = next(r for r in range(col, n) if M[r][col] % P)
        M[col], M[piv] = M[piv], M[col]
        inv = pow(M[col][col], P - 2, P)
        for r in range(col + 1, n):
            f = M[r][col] * inv % P
            M[r] = [(x - f * y) % P for x, y in zip(M[r], M[col])]
    assert all(M[i][i] % P for i in range(n))
    # permutation is deterministic and sensitive to every input lane
    s0 = np.arange(16, dtype=np.uint32)
    out0 = oracle.poseidon_permute(s0)
    assert np.array_equal(out0, oracle.poseidon_permute(s0))
    for i in range(16):
        s1 = s0.copy(); s1[i] += 1
        assert not np.array_equal(out0, oracle.poseidon_permute(s1))
    assert oracle.rc480.max() < P and len(set(oracle.rc480.tolist())) > 470


def test_challenger_duplex_semantics(oracle):
    # observe x16 triggers a duplexing; sample pops from the END of the state; observe clears buffered output
    ops = [0] * 16 + [1, 1]
    args = list(range(1, 17)) + [0, 0]
    out = oracle.challenger_script(ops, args)
    state = oracle.poseidon_permute(np.arange(1, 17, dtype=np.uint32))
    assert out[16] == state[15] and out[17] == state[14]
    # grinding: the returned witness satisfies the check (low bits zero), and it is the smallest one
    out = oracle.challenger_script([0, 3], [7, 8])
    w = int(out[1])
    for cand in range(w + 1):
        r = oracle.challenger_script([0, 0, 2], [7, cand, 8])
        assert (int(r[2]) == 0) == (cand == w)
