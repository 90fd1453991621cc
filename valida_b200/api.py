"""ctypes binding of libvalida_b200.so, shaped after the reference's Rust interfaces.

Reference interfaces mirrored (names and argument meaning kept):
  * p3_dft::TwoAdicSubgroupDft: dft_batch / idft_batch / coset_lde_batch        -> Radix2Dft
  * UnivariatePcsWithLde (machine/src/config.rs:17-22): commit_batches,
    commit_shifted_batches, get_ldes, coset_shift, log_blowup                    -> TwoAdicFriPcs
  * Machine::run / Chip::generate_trace (machine/src/machine.rs:13-30)          -> run_program / MachineTraces
Errors: the reference panics (derive/src/lib.rs:319,364,396); here every non-zero status raises
VgpuError carrying vgpu_last_error().
"""
import ctypes as C
import os

import numpy as np

BABYBEAR_P = 2013265921
REPR_CANONICAL, REPR_MONTY_R32 = 0, 1
NUM_CHIPS = 14

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libvalida_b200.so")


class VgpuError(RuntimeError):
    pass


class _Matrix(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint32)), ("height", C.c_uint64), ("width", C.c_uint64)]


def _load():
    if not os.path.exists(lib_path):
        raise VgpuError(
            "libvalida_b200.so is missing (%s): build it with `python -m valida_b200.build` — there is no CPU fallback" % lib_path
        )
    L = C.CDLL(lib_path)
    vp, u32p, u64 = C.c_void_p, C.POINTER(C.c_uint32), C.c_uint64
    sig = {
        "vgpu_ctx_create": (C.c_int32, [C.c_int32, vp, C.POINTER(vp)]),
        "vgpu_ctx_destroy": (None, [vp]),
        "vgpu_last_error": (C.c_char_p, [vp]),
        "vgpu_ctx_synchronize": (C.c_int32, [vp]),
        "vgpu_ctx_launch_count": (u64, [vp]),
        "vgpu_dmat_upload": (C.c_int32, [vp, C.POINTER(_Matrix), C.c_int32, C.POINTER(vp)]),
        "vgpu_dmat_download": (C.c_int32, [vp, vp, C.c_int32, u32p]),
        "vgpu_dmat_dims": (C.c_int32, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "vgpu_dmat_free": (None, [vp]),
        "vgpu_ntt_batch": (C.c_int32, [vp, vp, C.c_int32]),
        "vgpu_coset_lde_batch": (C.c_int32, [vp, vp, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(vp)]),
        "vgpu_ntt_batch_host": (C.c_int32, [vp, u32p, u64, u64, C.c_int32, C.c_int32]),
        "vgpu_commit_batches": (C.c_int32, [vp, C.POINTER(vp), C.c_uint32, u32p, u32p, C.POINTER(vp)]),
        "vgpu_commit_batches_host": (C.c_int32, [vp, C.POINTER(_Matrix), C.c_uint32, C.c_int32, u32p, u32p, C.POINTER(vp)]),
        "vgpu_prover_data_lde": (C.c_int32, [vp, C.c_uint32, C.POINTER(vp)]),
        "vgpu_prover_data_free": (None, [vp]),
        "vgpu_machine_run": (C.c_int32, [C.POINTER(C.c_int32), u64, C.c_uint32, C.c_uint32, u64, C.POINTER(vp), C.c_char_p, u64]),
        "vgpu_traces_main": (C.POINTER(_Matrix), [vp, C.c_uint32]),
        "vgpu_traces_preprocessed": (C.POINTER(_Matrix), [vp, C.c_uint32]),
        "vgpu_traces_stats": (None, [vp, u32p, u32p, u32p]),
        "vgpu_traces_mem_cell": (C.c_int32, [vp, C.c_uint32, u32p]),
        "vgpu_traces_free": (None, [vp]),
        "vgpu_fib_program": (u64, [C.c_uint32, C.POINTER(C.c_int32)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _as_u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


def _mat(a):
    return _Matrix(a.ctypes.data_as(C.POINTER(C.c_uint32)), a.shape[0], a.shape[1])


class Context:
    """One context per device/stream (single-threaded)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        rc = lib().vgpu_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h))
        if rc != 0:
            msg = lib().vgpu_last_error(self._h).decode() if self._h else "context allocation failed"
            if self._h:
                lib().vgpu_ctx_destroy(self._h)
                self._h = None
            raise VgpuError(msg)

    def check(self, rc):
        if rc != 0:
            raise VgpuError(lib().vgpu_last_error(self._h).decode())

    def synchronize(self):
        self.check(lib().vgpu_ctx_synchronize(self._h))

    @property
    def launch_count(self):
        return int(lib().vgpu_ctx_launch_count(self._h))

    def upload(self, row_major, repr=REPR_CANONICAL):
        """RowMajorMatrix<Val> (numpy h x w uint32) -> DeviceMatrix."""
        a = _as_u32(row_major)
        if a.ndim != 2:
            raise ValueError("expected a 2-D row-major matrix")
        m = _mat(a)
        out = C.c_void_p()
        self.check(lib().vgpu_dmat_upload(self._h, C.byref(m), repr, C.byref(out)))
        return DeviceMatrix(self, out)

    def close(self):
        if getattr(self, "_h", None):
            lib().vgpu_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMatrix:
    def __init__(self, ctx, handle, owned=True):
        self.ctx, self._h, self._owned = ctx, handle, owned

    @property
    def shape(self):
        h, w = C.c_uint64(), C.c_uint64()
        lib().vgpu_dmat_dims(self._h, C.byref(h), C.byref(w))
        return int(h.value), int(w.value)

    def download(self, repr=REPR_CANONICAL):
        h, w = self.shape
        out = np.empty((h, w), dtype=np.uint32)
        self.ctx.check(lib().vgpu_dmat_download(self.ctx._h, self._h, repr, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def free(self):
        if self._h and self._owned:
            lib().vgpu_dmat_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Radix2Dft:
    """p3_dft::TwoAdicSubgroupDft over device matrices (Radix2DitParallel / Radix2Bowers give the same values)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def dft_batch(self, m):
        self.ctx.check(lib().vgpu_ntt_batch(self.ctx._h, m._h, 0))
        return m

    def idft_batch(self, m):
        self.ctx.check(lib().vgpu_ntt_batch(self.ctx._h, m._h, 1))
        return m

    def coset_lde_batch(self, m, added_bits, shift, bit_reversed=False):
        out = C.c_void_p()
        self.ctx.check(lib().vgpu_coset_lde_batch(self.ctx._h, m._h, added_bits, shift, 1 if bit_reversed else 0, C.byref(out)))
        return DeviceMatrix(self.ctx, out)


class ProverData:
    def __init__(self, ctx, handle, n):
        self.ctx, self._h, self.n = ctx, handle, n

    def free(self):
        if self._h:
            lib().vgpu_prover_data_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class TwoAdicFriPcs:
    """UnivariatePcsWithLde surface used by Machine::prove (machine/src/config.rs:17-22)."""

    GENERATOR = 31

    def __init__(self, ctx, log_blowup=1, num_queries=40, proof_of_work_bits=8):
        if log_blowup != 1:
            raise VgpuError("only log_blowup = 1 is built (basic/src/bin/valida.rs:385-390)")
        self.ctx, self._log_blowup = ctx, log_blowup
        self.num_queries, self.proof_of_work_bits = num_queries, proof_of_work_bits

    def coset_shift(self):
        return self.GENERATOR

    def log_blowup(self):
        return self._log_blowup

    def commit_batches(self, polynomials):
        return self.commit_shifted_batches(polynomials, None)

    def commit_shifted_batches(self, polynomials, coset_shifts):
        n = len(polynomials)
        digest = (C.c_uint32 * 8)()
        out = C.c_void_p()
        shifts = None
        if coset_shifts is not None:
            shifts = (C.c_uint32 * n)(*[int(s) for s in coset_shifts])
        if n and isinstance(polynomials[0], DeviceMatrix):
            arr = (C.c_void_p * n)(*[m._h for m in polynomials])
            self.ctx.check(lib().vgpu_commit_batches(self.ctx._h, arr, n, shifts, digest, C.byref(out)))
        else:
            keep = [_as_u32(p) for p in polynomials]
            arr = (_Matrix * n)(*[_mat(a) for a in keep])
            self.ctx.check(lib().vgpu_commit_batches_host(self.ctx._h, arr, n, REPR_CANONICAL, shifts, digest, C.byref(out)))
        return np.array(list(digest), dtype=np.uint32), ProverData(self.ctx, out, n)

    def get_ldes(self, prover_data):
        """Committed LDEs (rows stored bit-reversed), borrowed views."""
        out = []
        for i in range(prover_data.n):
            v = C.c_void_p()
            self.ctx.check(lib().vgpu_prover_data_lde(prover_data._h, i, C.byref(v)))
            out.append(DeviceMatrix(self.ctx, v, owned=False))
        return out


class MachineTraces:
    """Host witness of one BasicMachine run: 14 main traces (chip order) + 2 preprocessed traces."""

    CHIPS = ["cpu", "program", "mem", "add", "sub", "mul", "div", "shift", "lt", "com", "bitwise", "output", "range", "static_data"]

    def __init__(self, handle):
        self._h = handle
        L = lib()
        self.main = []
        for i in range(NUM_CHIPS):
            m = L.vgpu_traces_main(handle, i).contents
            self.main.append(np.ctypeslib.as_array(m.data, shape=(m.height * m.width,)).reshape(m.height, m.width))
        self.preprocessed = []
        for i in range(2):
            m = L.vgpu_traces_preprocessed(handle, i).contents
            self.preprocessed.append(np.ctypeslib.as_array(m.data, shape=(m.height * m.width,)).reshape(m.height, m.width))
        c, mo, ao = C.c_uint32(), C.c_uint32(), C.c_uint32()
        L.vgpu_traces_stats(handle, C.byref(c), C.byref(mo), C.byref(ao))
        self.clock, self.mem_ops, self.add_ops = c.value, mo.value, ao.value

    def mem_cell(self, addr):
        v = C.c_uint32()
        if lib().vgpu_traces_mem_cell(self._h, addr, C.byref(v)) != 0:
            return None
        return v.value

    def free(self):
        if self._h:
            self.main, self.preprocessed = [], []
            lib().vgpu_traces_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fib_program(n):
    """fib_program() of basic/tests/test_prover.rs:35-188 with the `imm32 -8(fp)` operand set to n."""
    words = (C.c_int32 * (23 * 6))()
    cnt = lib().vgpu_fib_program(n, words)
    return np.array(list(words), dtype=np.int32).reshape(int(cnt), 6)


def run_program(program, initial_fp=0x1000, initial_pc=0, max_cycles=1 << 30):
    """Machine::run + generate_trace for every chip (host)."""
    p = np.ascontiguousarray(program, dtype=np.int32)
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib().vgpu_machine_run(p.ctypes.data_as(C.POINTER(C.c_int32)), p.shape[0], initial_pc, initial_fp, max_cycles, C.byref(h), err, 512)
    if rc != 0:
        raise VgpuError(err.value.decode())
    return MachineTraces(h)
