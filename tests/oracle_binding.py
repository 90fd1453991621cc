"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (never imported by valida_b200/)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2013265921
u32p = C.POINTER(C.c_uint32)


def _p(a):
    return a.ctypes.data_as(u32p)


class Oracle:
    def __init__(self):
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        self.L = L = C.CDLL(path)
        L.orc_two_adic_generator.restype = C.c_uint32
        L.orc_mul.restype = C.c_uint32
        L.orc_inv.restype = C.c_uint32
        L.orc_prove.restype = C.c_void_p
        L.orc_proof_cbor.restype = C.c_uint64
        L.orc_proof_perm_trace.restype = C.c_uint64
        L.orc_proof_quotient_chunks.restype = C.c_uint64
        L.orc_proof_constraint_failure.restype = C.c_int64
        L.orc_proof_opened.restype = C.c_uint64
        L.orc_chip_perm_width.restype = C.c_uint32
        L.orc_chip_width.restype = C.c_uint32
        L.orc_chip_prep_width.restype = C.c_uint32
        rc = (C.c_uint32 * 480)()
        L.orc_default_round_constants(rc)
        self.rc480 = np.array(list(rc), dtype=np.uint32)

    def set_threads(self, n):
        self.L.orc_set_threads(int(n))

    def max_threads(self):
        return int(self.L.orc_max_threads())

    def tune_allocator(self):
        """Process-wide malloc settings for a process that only times the oracle (bench.py --impl reference)."""
        return int(self.L.orc_tune_allocator())

    def prefault_heap(self, nbytes):
        """Maps and touches `nbytes` of heap on all threads and frees them again (after tune_allocator the pages stay mapped)."""
        self.L.orc_prefault_heap.restype = C.c_uint64
        return int(self.L.orc_prefault_heap(C.c_uint64(int(nbytes))))

    # ---- primitives ----
    def keccak256(self, data, pad=0x01):
        out = (C.c_uint8 * 32)()
        self.L.orc_keccak256(bytes(data), C.c_uint64(len(data)), out, C.c_uint32(pad))
        return bytes(out)

    def dft(self, m, inverse=False):
        a = np.ascontiguousarray(m, dtype=np.uint32).copy()
        self.L.orc_dft(_p(a), C.c_uint64(a.shape[0]), C.c_uint64(a.shape[1]), int(inverse))
        return a

    def naive_dft(self, m):
        a = np.ascontiguousarray(m, dtype=np.uint32)
        out = np.empty_like(a)
        self.L.orc_naive_dft(_p(a), C.c_uint64(a.shape[0]), C.c_uint64(a.shape[1]), _p(out))
        return out

    def coset_lde(self, m, added_bits, shift, bitrev):
        a = np.ascontiguousarray(m, dtype=np.uint32)
        out = np.empty((a.shape[0] << added_bits, a.shape[1]), dtype=np.uint32)
        self.L.orc_coset_lde(_p(a), C.c_uint64(a.shape[0]), C.c_uint64(a.shape[1]), C.c_uint32(added_bits), C.c_uint32(shift), int(bitrev), _p(out))
        return out

    def merkle_root(self, mats):
        mats = [np.ascontiguousarray(m, dtype=np.uint32) for m in mats]
        n = len(mats)
        ptrs = (u32p * n)(*[_p(m) for m in mats])
        hs = (C.c_uint64 * n)(*[m.shape[0] for m in mats])
        ws = (C.c_uint64 * n)(*[m.shape[1] for m in mats])
        d = np.zeros(8, dtype=np.uint32)
        self.L.orc_merkle_root(n, ptrs, hs, ws, _p(d))
        return d

    def commit_batches(self, mats, coset_shifts=None, want_ldes=False):
        mats = [np.ascontiguousarray(m, dtype=np.uint32) for m in mats]
        n = len(mats)
        ptrs = (u32p * n)(*[_p(m) for m in mats])
        hs = (C.c_uint64 * n)(*[m.shape[0] for m in mats])
        ws = (C.c_uint64 * n)(*[m.shape[1] for m in mats])
        sh = (C.c_uint32 * n)(*[int(s) for s in coset_shifts]) if coset_shifts is not None else None
        d = np.zeros(8, dtype=np.uint32)
        ldes = [np.empty((2 * m.shape[0], m.shape[1]), dtype=np.uint32) for m in mats] if want_ldes else None
        lp = (u32p * n)(*[_p(l) for l in ldes]) if want_ldes else None
        self.L.orc_commit_batches(n, ptrs, hs, ws, sh, _p(d), lp)
        return (d, ldes) if want_ldes else d

    def poseidon_permute(self, state, rc=None):
        rc = self.rc480 if rc is None else rc
        s = np.ascontiguousarray(state, dtype=np.uint32).copy()
        self.L.orc_poseidon_permute(_p(np.ascontiguousarray(rc)), _p(s))
        return s

    def coset_mds(self):
        out = np.zeros(256, dtype=np.uint32)
        self.L.orc_coset_mds_matrix(_p(out))
        return out.reshape(16, 16)

    def challenger_script(self, ops, args, rc=None):
        rc = self.rc480 if rc is None else rc
        ops = np.ascontiguousarray(ops, dtype=np.uint32)
        args = np.ascontiguousarray(args, dtype=np.uint32)
        out = np.zeros(len(ops), dtype=np.uint32)
        self.L.orc_challenger_script(_p(np.ascontiguousarray(rc)), len(ops), _p(ops), _p(args), _p(out))
        return out

    def ext_mul(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint32); b = np.ascontiguousarray(b, dtype=np.uint32)
        o = np.zeros(5, dtype=np.uint32)
        self.L.orc_ext_mul(_p(a), _p(b), _p(o))
        return o

    def ext_inv(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        o = np.zeros(5, dtype=np.uint32)
        self.L.orc_ext_inv(_p(a), _p(o))
        return o

    # ---- machine ----
    def prove(self, main, preps, rc=None, debug_checks=True):
        rc = self.rc480 if rc is None else rc
        main = [np.ascontiguousarray(m, dtype=np.uint32) for m in main]
        preps = [np.ascontiguousarray(m, dtype=np.uint32) for m in preps]
        ptrs = (u32p * 14)(*[_p(m) for m in main])
        hs = (C.c_uint64 * 14)(*[m.shape[0] for m in main])
        ws = (C.c_uint64 * 14)(*[m.shape[1] for m in main])
        h = self.L.orc_prove(ptrs, hs, ws, _p(preps[0]), C.c_uint64(preps[0].shape[0]), _p(preps[1]), C.c_uint64(preps[1].shape[0]),
                             _p(np.ascontiguousarray(rc)), int(debug_checks))
        return OracleProof(self, C.c_void_p(h))

    def verify(self, proof_bytes, preps, rc=None):
        rc = self.rc480 if rc is None else rc
        preps = [np.ascontiguousarray(m, dtype=np.uint32) for m in preps]
        return int(self.L.orc_verify(bytes(proof_bytes), C.c_uint64(len(proof_bytes)), _p(preps[0]), C.c_uint64(preps[0].shape[0]),
                                     _p(preps[1]), C.c_uint64(preps[1].shape[0]), _p(np.ascontiguousarray(rc))))

    def open(self, rounds, points, observe, shifts=None, rc=None, sample_ext_first=False):
        """rounds: [[matrix, ...], ...]; points: per matrix a list of ext5 points (canonical 5-tuples); observe: words
        absorbed before opening.  Returns the CBOR of (opened_values, proof)."""
        rc = self.rc480 if rc is None else rc
        flat = [np.ascontiguousarray(m, dtype=np.uint32) for r in rounds for m in r]
        n = len(flat)
        per_round = np.array([len(r) for r in rounds], dtype=np.uint32)
        ptrs = (u32p * n)(*[_p(m) for m in flat])
        hs = (C.c_uint64 * n)(*[m.shape[0] for m in flat])
        ws = (C.c_uint64 * n)(*[m.shape[1] for m in flat])
        npts = np.array([len(p) for p in points], dtype=np.uint32)
        pts = np.array([w for p in points for z in p for w in z], dtype=np.uint32)
        obs = np.ascontiguousarray(observe, dtype=np.uint32).reshape(-1)
        sh = _p(np.ascontiguousarray(shifts, dtype=np.uint32)) if shifts is not None else None
        self.L.orc_open.restype = C.c_uint64
        args = (_p(np.ascontiguousarray(rc)), C.c_uint32(len(rounds)), _p(per_round), ptrs, hs, ws, sh, _p(npts), _p(pts), _p(obs), C.c_uint32(obs.size), C.c_uint32(1 if sample_ext_first else 0))
        size = self.L.orc_open(*args, None, C.c_uint64(0))
        buf = (C.c_uint8 * size)()
        self.L.orc_open(*args, buf, C.c_uint64(size))
        return bytes(buf)

    def chip_width(self, chip):
        return int(self.L.orc_chip_width(chip))

    def chip_perm_width(self, chip):
        return int(self.L.orc_chip_perm_width(chip))

    def chip_prep_width(self, chip):
        return int(self.L.orc_chip_prep_width(chip))

    def perm_trace(self, chip, main, prep, challenges15):
        main = np.ascontiguousarray(main, dtype=np.uint32)
        h = main.shape[0]
        out = np.empty((h, self.chip_perm_width(chip)), dtype=np.uint32)
        cs = np.zeros(5, dtype=np.uint32)
        pp = _p(np.ascontiguousarray(prep, dtype=np.uint32)) if prep is not None else None
        self.L.orc_perm_trace(chip, _p(main), C.c_uint64(h), pp, _p(np.ascontiguousarray(challenges15, dtype=np.uint32)), _p(out), _p(cs))
        return out, cs

    def quotient(self, chip, log_degree, prep_lde, main_lde, perm_lde, cumsum5, challenges15, alpha5):
        h = 1 << log_degree
        out = np.empty((h, 10), dtype=np.uint32)
        c = lambda a: _p(np.ascontiguousarray(a, dtype=np.uint32))
        self.L.orc_quotient(chip, log_degree, c(prep_lde) if prep_lde is not None else None, c(main_lde), c(perm_lde), c(cumsum5), c(challenges15), c(alpha5), _p(out))
        return out


class OracleProof:
    def __init__(self, orc, h):
        self.o, self.h = orc, h

    def cbor(self):
        bp = C.POINTER(C.c_uint8)()
        n = self.o.L.orc_proof_cbor(self.h, C.byref(bp))
        return bytes(bytearray(bp[:n]))

    def transcript(self):
        pc = np.zeros(15, dtype=np.uint32); al = np.zeros(5, dtype=np.uint32); ze = np.zeros(5, dtype=np.uint32); cm = np.zeros(32, dtype=np.uint32)
        self.o.L.orc_proof_transcript(self.h, _p(pc), _p(al), _p(ze), _p(cm))
        return dict(perm_challenges=pc, alpha=al, zeta=ze, prep_commit=cm[0:8], main_commit=cm[8:16], perm_commit=cm[16:24], quotient_commit=cm[24:32])

    def perm_trace(self, chip):
        d = u32p(); w = C.c_uint64()
        h = self.o.L.orc_proof_perm_trace(self.h, chip, C.byref(d), C.byref(w))
        return np.ctypeslib.as_array(d, shape=(h * w.value,)).reshape(h, w.value).copy()

    def quotient_chunks(self, chip):
        d = u32p()
        h = self.o.L.orc_proof_quotient_chunks(self.h, chip, C.byref(d))
        return np.ctypeslib.as_array(d, shape=(h * 10,)).reshape(h, 10).copy()

    def cumulative_sum(self, chip):
        o = np.zeros(5, dtype=np.uint32)
        self.o.L.orc_proof_cumulative_sum(self.h, chip, _p(o))
        return o

    def constraint_failures(self):
        return [int(self.o.L.orc_proof_constraint_failure(self.h, i)) for i in range(14)]

    def cumulative_sum_zero(self):
        return bool(self.o.L.orc_proof_cumulative_sum_zero(self.h))

    def opened(self, chip, which):
        n = self.o.L.orc_proof_opened(self.h, chip, which, None, C.c_uint64(0))
        out = np.zeros((n, 5), dtype=np.uint32)
        self.o.L.orc_proof_opened(self.h, chip, which, _p(out), C.c_uint64(n))
        return out

    def __del__(self):
        try:
            self.o.L.orc_proof_free(self.h)
        except Exception:
            pass
