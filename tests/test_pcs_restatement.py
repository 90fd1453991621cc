"""An independent, plain-Python reading of the Plonky3-side conventions the proof bytes depend on, written
from the published design of the p3 crates and the Rust call sites — NOT from oracle/*.h — and in a different formulation
on purpose (recursive Merkle tree, transforms by the defining sums, FRI folding on COEFFICIENTS rather than on evaluation
pairs, a sponge written as a small state machine), then compared byte for byte with the oracle on 2^2 .. 2^7-row inputs.
The oracle (C++) and the product's host code (C++) restate these conventions twice by the same hand; this third text is
the one whose disagreement would show a slip in either.  Parity with the real Plonky3 stays UNPINNED (no Rust toolchain,
no golden vectors in the reference: SURVEY D1, D7): each convention below is a named switch, to be flipped the day
`valida prove` output can be diffed.

App. A items covered: 2 (generators), 4/5 (coset LDE, shift = generator / coset_shift, bit-reversed rows), 6/7/8 (mixed-height
FieldMerkleTree, SerializingHasher32<Keccak256>, CompressionFunctionFromHasher), 9 (ExtensionMmcs leaves of FRI layers),
10/11/12 (Poseidon round structure, CosetMds, DuplexChallenger), 14/15 (open_multi_batches, FRI prover, PoW, query index ->
row), 17 (evaluation at a point), 20 (serde/ciborium shapes; field names and order from machine/src/proof.rs:13-44)."""
import cbor2
import numpy as np
import pytest

from test_perm_trace_restatement import P, e_add, e_inv, e_mul, e_sub

# ---- switches (App. A open items): the values below are the oracle's current reading -----------------------------------
SWITCH = {
    "generator": 31,                        # item 2: BabyBear::generator()
    "two_adic_generator_27": 0x1A427A41,    # item 2
    "keccak_pad": 0x01,                     # item 7: tiny-keccak Keccak::v256 (pre-SHA3 padding)
    "digest_word_reduction": "mod_p",       # item 7: 32-byte hash -> 8 LE u32 words, each reduced mod p (from_wrapped_u32)
    "leaf_words": "canonical_le",           # item 7: as_canonical_u32().to_le_bytes()
    "inject": "compress(compress(l,r),rows)",   # item 6
    "commit_row_order": "bit_reversed",     # item 5
    "lde_shift": "generator/coset_shift",   # item 5
    "sponge_absorb": "overwrite",           # item 12: duplexing overwrites the first |input| lanes
    "sponge_sample_from": "end",            # item 12: sample pops from the END of the output buffer
    "sponge_rate": 16,                      # item 12: 3-parameter DuplexChallenger => rate = WIDTH
    "mds_twist": "31^slot on bit-reversed slots",   # item 11
    "pow": "smallest_witness",              # item 15 (rayon find_any upstream: any witness verifies)
    "fri_fold": "(fe + beta*fo)(x^2)",      # item 15
    "reduced_opening_order": "per height: matrices in commit order, points in order, columns ascending",   # item 14
    "felt_serde": "map{value: montgomery u32}",     # items 1, 20
}
R = pow(2, 32, P)          # Montgomery radix of p3-baby-bear


# ---- field helpers -------------------------------------------------------------------------------------------------------
def two_adic_generator(bits):
    return pow(SWITCH["two_adic_generator_27"], 1 << (27 - bits), P)


def brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def e_scale(a, k):
    return [x * k % P for x in a]


def e_pow(a, n):
    r = [1, 0, 0, 0, 0]
    while n:
        if n & 1:
            r = e_mul(r, a)
        a = e_mul(a, a)
        n >>= 1
    return r


# ---- Keccak-256 (own text: lanes as Python ints) ---------------------------------------------------------------------------
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]   # [x][y]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f(a):   # a[x][y]
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) & _M64 for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data):
    rate = 136
    msg = bytearray(data)
    msg.append(SWITCH["keccak_pad"])
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        a = keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def hash_words(words):
    """SerializingHasher32<Keccak256Hash> + the [BabyBear; 8] digest."""
    out = keccak256(b"".join(int(w).to_bytes(4, "little") for w in words))
    return [int.from_bytes(out[4 * i: 4 * i + 4], "little") % P for i in range(8)]


def compress(l, r):
    return hash_words(list(l) + list(r))


# ---- coset LDE by the defining sums, committed (bit-reversed) row order ------------------------------------------------------
def coset_lde_column(col, shift):
    """coset_lde_batch(.., added_bits = 1, shift): evaluations on the subgroup -> evaluations on shift * K, |K| = 2n, natural order."""
    n = len(col)
    lg = n.bit_length() - 1
    w = two_adic_generator(lg)
    ninv = pow(n, P - 2, P)
    winv = pow(w, P - 2, P)
    coef = [sum(col[j] * pow(winv, j * k % n, P) for j in range(n)) * ninv % P for k in range(n)]
    w2 = two_adic_generator(lg + 1)
    out = []
    for i in range(2 * n):
        x = shift * pow(w2, i, P) % P
        acc = 0
        for c in reversed(coef):   # Horner
            acc = (acc * x + c) % P
        out.append(acc)
    return coef, out


def commit_ldes(mats, coset_shifts=None):
    """TwoAdicFriPcs::commit_shifted_batches up to (not including) the tree: per matrix the bit-reversed LDE rows."""
    ldes, coefs = [], []
    for i, m in enumerate(mats):
        cs = 1 if coset_shifts is None else coset_shifts[i]
        shift = SWITCH["generator"] * pow(cs, P - 2, P) % P
        h, w = len(m), len(m[0])
        cols, cfs = [], []
        for c in range(w):
            cf, ev = coset_lde_column([int(m[r][c]) for r in range(h)], shift)
            cols.append(ev); cfs.append(cf)
        lg = (2 * h).bit_length() - 1
        ldes.append([[cols[c][brev(r, lg)] for c in range(w)] for r in range(2 * h)])
        coefs.append((cfs, shift))
    return ldes, coefs


# ---- FieldMerkleTree, recursively -------------------------------------------------------------------------------------------
class Tree:
    def __init__(self, mats):
        self.mats = mats
        self.max_h = max(len(m) for m in mats)
        self.memo = {}

    def rows_hash(self, height, j):
        group = [m for m in self.mats if len(m) == height]     # caller's order within a height (stable sort upstream)
        if not group:
            return None
        return hash_words([w for m in group for w in m[j]])

    def node(self, height, j):
        """Digest of node j of the layer with `height` nodes."""
        key = (height, j)
        if key not in self.memo:
            if height == self.max_h:
                d = self.rows_hash(height, j)
            else:
                d = compress(self.node(2 * height, 2 * j), self.node(2 * height, 2 * j + 1))
                inj = self.rows_hash(height, j)
                if inj is not None:
                    d = compress(d, inj)
            self.memo[key] = d
        return self.memo[key]

    def root(self):
        return self.node(1, 0)

    def open(self, index):
        """(rows of every matrix at `index` scaled to its height, sibling digests leaf level first)."""
        lg = self.max_h.bit_length() - 1
        rows = [m[index >> (lg - (len(m).bit_length() - 1))] for m in self.mats]
        path = [self.node(self.max_h >> lvl, (index >> lvl) ^ 1) for lvl in range(lg)]
        return rows, path


# ---- Poseidon-16 / CosetMds / DuplexChallenger -----------------------------------------------------------------------------
def coset_mds_apply(v):
    """CosetMds<_, 16>: unscaled inverse transform to frequency slots in bit-reversed order, slot q times generator^q, forward
    transform.  Written on frequencies: c_f = sum_j v_j w^(-jf); slot of frequency f is rev(f); out_k = sum_f c_f g^rev(f) w^(fk)."""
    w = two_adic_generator(4)
    winv = pow(w, P - 2, P)
    c = [sum(v[j] * pow(winv, j * f % 16, P) for j in range(16)) % P for f in range(16)]
    tw = [c[f] * pow(SWITCH["generator"], brev(f, 4), P) % P for f in range(16)]
    return [sum(tw[f] * pow(w, f * k % 16, P) for f in range(16)) % P for k in range(16)]


def poseidon(state, rc):
    s = list(state)
    r = 0
    for phase, n in (("full", 4), ("partial", 22), ("full", 4)):      # Poseidon::new_from_rng(half_num_full_rounds = 4, num_partial_rounds = 22, ..)
        for _ in range(n):
            s = [(s[i] + int(rc[16 * r + i])) % P for i in range(16)]
            if phase == "full":
                s = [pow(x, 5, P) for x in s]
            else:
                s[0] = pow(s[0], 5, P)
            s = coset_mds_apply(s)
            r += 1
    return s


class Duplex:
    def __init__(self, rc):
        self.rc, self.state, self.inp, self.out = rc, [0] * 16, [], []

    def _duplex(self):
        assert SWITCH["sponge_absorb"] == "overwrite"
        for i, v in enumerate(self.inp):
            self.state[i] = v
        self.inp = []
        self.state = poseidon(self.state, self.rc)
        self.out = list(self.state)

    def observe(self, v):
        self.out = []
        self.inp.append(v % P)
        if len(self.inp) == SWITCH["sponge_rate"]:
            self._duplex()

    def observe_digest(self, d):
        for v in d:
            self.observe(v)

    def sample(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()          # from the END

    def sample_ext(self):
        return [self.sample() for _ in range(5)]

    def sample_bits(self, bits):
        return self.sample() & ((1 << bits) - 1)

    def grind(self, bits):
        w = 0
        while True:
            trial = Duplex(self.rc)
            trial.state, trial.inp, trial.out = list(self.state), list(self.inp), list(self.out)
            trial.observe(w)
            if trial.sample_bits(bits) == 0:
                self.observe(w)
                assert self.sample_bits(bits) == 0
                return w
            w += 1


# ---- open_multi_batches + p3-fri prover on COEFFICIENTS ---------------------------------------------------------------------
def poly_eval_ext(coef, z):
    acc = [0] * 5
    for c in reversed(coef):
        acc = e_mul(acc, z)
        acc = e_add(acc, c)
    return acc


def open_multi_batches(rounds_mats, rounds_shifts, points, ch, num_queries=40, pow_bits=8):
    """rounds_mats[r] = matrices of commit r (unextended, row-major lists).  Returns the python image of (opened_values, proof).
    Everything FRI sees is kept as polynomial COEFFICIENTS in the variable Y = X / generator (the committed domain g * K is the
    plain subgroup K in Y, and p3-fri folds as if on K): the reduced opening of a height is the polynomial
    sum_k alpha^k (q_k(X) - q_k(z)) / (X - z) (synthetic division), a fold is fe + beta * fo, the reduced opening of the next height
    is added coefficient-wise, and values appear only when a layer's coefficients are evaluated on its bit-reversed subgroup."""
    g = SWITCH["generator"]
    alpha = ch.sample_ext()
    trees = []
    red, count, values = {}, {}, []      # log2(LDE height) -> coefficient list in Y / number of (column, point) terms so far
    mat_no = 0
    for mats, shifts in zip(rounds_mats, rounds_shifts):
        ldes, coefs = commit_ldes(mats, shifts)
        trees.append(Tree(ldes))
        values.append([])
        for mi, m in enumerate(mats):
            h = len(m)
            lh = (2 * h).bit_length() - 1
            cfs, shift = coefs[mi]
            # the committed rows are p(shift * K); as a function on g * K that is q(X) = p(X * shift / g): c_k -> c_k (shift / g)^k
            ratio = shift * pow(g, P - 2, P) % P
            qcoefs = [[c * pow(ratio, k, P) % P for k, c in enumerate(cf)] for cf in cfs]
            values[-1].append([])
            for z in points[mat_no]:
                ys = [poly_eval_ext([[c, 0, 0, 0, 0] for c in qc], z) for qc in qcoefs]
                values[-1][-1].append(ys)
                n0 = count.get(lh, 0)
                acc = red.setdefault(lh, [[0] * 5 for _ in range(h)])
                for k, (qc, y) in enumerate(zip(qcoefs, ys)):
                    quo = [None] * h            # (q(X) - y) / (X - z), synthetic division from the top coefficient down
                    carry = [0] * 5
                    for d in range(h - 1, -1, -1):
                        quo[d] = carry
                        carry = e_add([qc[d], 0, 0, 0, 0], e_mul(carry, z))
                    assert carry == y            # the remainder is q(z)
                    ak = e_pow(alpha, n0 + k)
                    for d in range(h):
                        acc[d] = e_add(acc[d], e_scale(e_mul(ak, quo[d]), pow(g, d, P)))      # X^d = g^d Y^d
                count[lh] = n0 + len(qcoefs)
            mat_no += 1
    log_max = max(red)

    def evals_bitrev(coef, log_n):
        w = two_adic_generator(log_n)
        return [poly_eval_ext(coef, [pow(w, brev(i, log_n), P), 0, 0, 0, 0]) for i in range(1 << log_n)]

    cur = red[log_max]
    commits, layer_trees = [], []
    for lfh in range(log_max - 1, 0, -1):
        ev = evals_bitrev(cur, lfh + 1)
        t = Tree([[ev[2 * i] + ev[2 * i + 1] for i in range(len(ev) // 2)]])        # ExtensionMmcs: width-2 ext matrix, flattened to 10 words
        layer_trees.append(t)
        root = t.root()
        ch.observe_digest(root)
        commits.append(root)
        beta = ch.sample_ext()
        cur = [e_add(a, e_mul(beta, b)) for a, b in zip(cur[0::2], cur[1::2])]       # f = fe(Y^2) + Y fo(Y^2)  ->  fe + beta fo
        if lfh in red:
            cur = [e_add(a, b) for a, b in zip(cur, red[lfh])]
    assert len(cur) == 1, "the final layer is a constant"
    pow_witness = ch.grind(pow_bits)
    indices = [ch.sample_bits(log_max) for _ in range(num_queries)]
    query_proofs, query_openings = [], []
    for index in indices:
        steps = []
        for i, t in enumerate(layer_trees):
            idx_i = index >> i
            rows, path = t.open(idx_i >> 1)
            half = (idx_i ^ 1) & 1
            steps.append((rows[0][5 * half: 5 * half + 5], path))
        query_proofs.append(steps)
        qo = []
        for t in trees:
            lg = t.max_h.bit_length() - 1
            qo.append(t.open(index >> (log_max - lg)))
        query_openings.append(qo)
    return values, dict(commits=commits, query_proofs=query_proofs, final=cur[0], pow=pow_witness, query_openings=query_openings)


# ---- serde / ciborium shapes --------------------------------------------------------------------------------------------------
def felt(x):        # BabyBear { value: u32 } with the Montgomery word (p3-baby-bear derives Serialize on the struct)
    return {"value": x * R % P}


def ext(e):         # BinomialExtensionField { value: [AF; D] }
    return {"value": [felt(x) for x in e]}


def digest(d):      # [BabyBear; 8]
    return [felt(x) for x in d]


def opening_to_cbor(values, pf):
    """(Vec<Vec<Vec<Vec<Challenge>>>>, TwoAdicFriPcsProof { fri_proof: FriProof { commit_phase_commits, query_proofs: [QueryProof {
    commit_phase_openings: [CommitPhaseProofStep { sibling_value, opening_proof }] }], final_poly, pow_witness }, query_openings:
    Vec<Vec<BatchOpening { opened_values, opening_proof }>> }) — struct field order is the declaration order (serde derive)."""
    v = [[[[ext(y) for y in at_point] for at_point in mat] for mat in rnd] for rnd in values]
    fri = {
        "commit_phase_commits": [digest(c) for c in pf["commits"]],
        "query_proofs": [{"commit_phase_openings": [{"sibling_value": ext(s), "opening_proof": [digest(d) for d in path]} for s, path in steps]}
                         for steps in pf["query_proofs"]],
        "final_poly": ext(pf["final"]),
        "pow_witness": felt(pf["pow"]),
    }
    qo = [[{"opened_values": [[felt(x) for x in row] for row in rows], "opening_proof": [digest(d) for d in path]} for rows, path in q]
          for q in pf["query_openings"]]
    return cbor2.dumps([v, {"fri_proof": fri, "query_openings": qo}])


# ---- tests -------------------------------------------------------------------------------------------------------------------
def _rand(rng, h, w):
    return rng.integers(0, P, (h, w), dtype=np.uint32)


def test_keccak_restatement_matches_oracle(oracle):
    rng = np.random.default_rng(1)
    for n in (0, 1, 55, 56, 135, 136, 137, 272, 300):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert keccak256(data) == oracle.keccak256(data)


def test_coset_lde_and_commit_root(oracle):
    rng = np.random.default_rng(2)
    for heights, shifts in (([16], None), ([8, 32, 8, 2, 1], None), ([16, 4, 16, 1], [1, 961, 5, 7])):
        mats = [_rand(rng, h, 1 + (i % 3)) for i, h in enumerate(heights)]
        ldes, _ = commit_ldes([m.tolist() for m in mats], shifts)
        want_root, want_ldes = oracle.commit_batches(mats, coset_shifts=shifts, want_ldes=True)
        for got, want in zip(ldes, want_ldes):
            assert np.array_equal(np.array(got, dtype=np.uint32), want)
        assert Tree(ldes).root() == list(want_root)


def test_poseidon_and_coset_mds(oracle):
    rng = np.random.default_rng(3)
    mds = oracle.coset_mds()
    for _ in range(3):
        v = rng.integers(0, P, 16, dtype=np.uint32)
        assert coset_mds_apply([int(x) for x in v]) == [int(sum(int(mds[i][j]) * int(v[j]) for j in range(16)) % P) for i in range(16)]
        assert poseidon([int(x) for x in v], oracle.rc480) == [int(x) for x in oracle.poseidon_permute(v)]


def test_duplex_challenger_script(oracle):
    rng = np.random.default_rng(4)
    ops, args = [], []
    for _ in range(120):
        op = int(rng.integers(0, 3))
        ops.append(op)
        args.append(int(rng.integers(0, P)) if op == 0 else int(rng.integers(1, 20)))
    ops += [3, 1, 2]; args += [8, 0, 5]      # grind(8), sample, sample_bits(5)
    ch = Duplex(oracle.rc480)
    got = []
    for op, a in zip(ops, args):
        got.append(0 if op == 0 and ch.observe(a) is None else ch.sample() if op == 1 else ch.sample_bits(a) if op == 2 else ch.grind(a))
    assert got == [int(x) for x in oracle.challenger_script(ops, args)]


@pytest.mark.parametrize("case", ["one_matrix", "mixed_heights_two_rounds", "shifted_round"])
def test_open_multi_batches_bytes(oracle, case):
    """FRI on coefficients + recursive trees + the state-machine sponge reproduce the oracle's (opened_values, proof) bytes."""
    rng = np.random.default_rng({"one_matrix": 5, "mixed_heights_two_rounds": 6, "shifted_round": 7}[case])
    z = [int(x) for x in rng.integers(0, P, 5)]
    z2 = [int(x) for x in rng.integers(0, P, 5)]
    if case == "one_matrix":
        rounds, shifts, pts = [[_rand(rng, 16, 3)]], [None], [[z, z2]]
    elif case == "mixed_heights_two_rounds":
        rounds = [[_rand(rng, 8, 2), _rand(rng, 32, 1), _rand(rng, 1, 3)], [_rand(rng, 8, 5), _rand(rng, 2, 1)]]
        shifts = [None, None]
        pts = [[z, z2], [z], [z, z2], [z2], [z]]
    else:
        rounds, shifts = [[_rand(rng, 16, 2)], [_rand(rng, 16, 10), _rand(rng, 4, 10)]], [None, [961, 961]]
        pts = [[z, z2], [e_mul(z, z)], [e_mul(z, z)]]
    observe = [int(x) for x in rng.integers(0, P, 11)]
    ch = Duplex(oracle.rc480)
    for v in observe:
        ch.observe(v)
    values, pf = open_multi_batches([[m.tolist() for m in r] for r in rounds], shifts, pts, ch)
    flat_shifts = None if all(s is None for s in shifts) else [x for r, s in zip(rounds, shifts) for x in (s or [1] * len(r))]
    want = oracle.open(rounds, pts, observe, shifts=flat_shifts)
    assert opening_to_cbor(values, pf) == want


def test_machine_proof_shape_follows_proof_rs(oracle):
    """Field names and their order in the oracle's MachineProof CBOR are those of machine/src/proof.rs:13-44 (serde derive writes
    struct fields in declaration order; ciborium maps keep it)."""
    import valida_b200 as vb

    t = vb.run_program(vb.fib_program(3), initial_fp=0x1000)
    d = cbor2.loads(oracle.prove(t.main, t.preprocessed, debug_checks=False).cbor())
    assert list(d) == ["commitments", "opening_proof", "chip_proofs"]                      # MachineProof, proof.rs:15-19
    assert list(d["commitments"]) == ["main_trace", "perm_trace", "quotient_chunks"]       # Commitments, proof.rs:22-26
    assert len(d["chip_proofs"]) == 14
    for cp in d["chip_proofs"]:
        assert list(cp) == ["log_degree", "opened_values", "cumulative_sum"]               # ChipProof, proof.rs:29-33
        assert list(cp["opened_values"]) == ["preprocessed_local", "preprocessed_next", "trace_local", "trace_next",
                                             "permutation_local", "permutation_next", "quotient_chunks"]   # OpenedValues, proof.rs:36-44
        assert list(cp["cumulative_sum"]) == ["value"] and len(cp["cumulative_sum"]["value"]) == 5
    assert list(d["opening_proof"]) == ["fri_proof", "query_openings"]
    assert list(d["opening_proof"]["fri_proof"]) == ["commit_phase_commits", "query_proofs", "final_poly", "pow_witness"]
    assert cbor2.dumps(d) == oracle.prove(t.main, t.preprocessed, debug_checks=False).cbor()   # definite lengths, shortest integers
