"""ONE proof split across 2 / 4 / 8 ranks: row shards after one peer-store exchange, sub-tree per rank,
sub-roots all-gathered — proof bytes identical to the single-GPU proof and to the oracle's.

The ranks here are THREADS of this process (vgpu_comm_init_local), all on device 0 when the box has one GPU: the whole
split data path (exchanges through peer pointers, shard-local sweeps, quotient "next" rows read from a peer's shard,
sharded FRI layers, owner-reported query answers) runs exactly as on several GPUs; only the transport of the small
collectives differs from the one-process-per-GPU launch (tests/test_gpu_multi.py, NCCL + CUDA IPC)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 2013265921


def _devices(n):
    import torch

    k = torch.cuda.device_count()
    return [i % k for i in range(n)]


def _group(n, oracle):
    import valida_b200 as vb

    ctxs = [vb.Context(d) for d in _devices(n)]
    vb.comm_init_local(ctxs)
    cfgs = [vb.StarkConfig(c, oracle.rc480) for c in ctxs]
    return ctxs, cfgs


def _close(ctxs):
    for c in ctxs:
        c.close()


@pytest.fixture(scope="module")
def fib15(built):
    import valida_b200 as vb

    t = vb.run_program(vb.fib_program(((1 << 15) - 17) // 7), initial_fp=0x1000)
    assert t.main[0].shape[0] == 1 << 15 and t.main[2].shape[0] == 1 << 17
    return t


@pytest.fixture(scope="module")
def fib15_proof(ctx, oracle, fib15):
    import valida_b200 as vb

    proof = vb.prove_machine(vb.StarkConfig(ctx, oracle.rc480), fib15)
    assert proof == oracle.prove(fib15.main, fib15.preprocessed, debug_checks=False).cbor()
    return proof


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_split_prove_bytes_identical(oracle, fib15, fib15_proof, nranks):
    """vgpu_prove with host traces: every rank uploads its rows only, proves, and returns the single-GPU bytes."""
    import valida_b200 as vb

    ctxs, cfgs = _group(nranks, oracle)
    try:
        proofs = vb.run_ranks(lambda r, c: vb.prove_machine(cfgs[r], fib15), ctxs)
        assert all(p == fib15_proof for p in proofs)
        # a second proof reuses the symmetric heap and the cached buffers
        proofs = vb.run_ranks(lambda r, c: vb.prove_machine(cfgs[r], fib15), ctxs)
        assert all(p == fib15_proof for p in proofs)
        stats = ctxs[0].comm_stats()
        assert stats["exchange"][0] > 0 and stats["allgather"][0] > 0
    finally:
        _close(ctxs)


def test_split_prove_reserves_the_whole_proof(oracle, fib15, fib15_proof, monkeypatch):
    """With no floor under the symmetric heap the first proof must size it for all three commits at once."""
    import valida_b200 as vb

    monkeypatch.setenv("VGPU_SYMM_HEAP_MIN_MB", "1")
    ctxs, cfgs = _group(2, oracle)
    try:
        assert all(p == fib15_proof for p in vb.run_ranks(lambda r, c: vb.prove_machine(cfgs[r], fib15), ctxs))
    finally:
        _close(ctxs)


def test_split_prove_device_resident_inputs(oracle, fib15, fib15_proof):
    """vgpu_prove_device on whole traces (every rank holds all rows) and on row shards (vgpu_dmat_upload_rows)."""
    import valida_b200 as vb

    ctxs, cfgs = _group(4, oracle)
    try:
        def whole(r, c):
            dm = [c.upload(m) for m in fib15.main]
            dp = [c.upload(m) for m in fib15.preprocessed]
            return vb.prove_machine(cfgs[r], fib15, device_resident=(dm, dp))

        def shards(r, c):
            dm = [c.upload_rows(m) for m in fib15.main]
            dp = [c.upload_rows(m) for m in fib15.preprocessed]
            assert dm[0].shape == fib15.main[0].shape          # logical dimensions
            return vb.prove_machine(cfgs[r], fib15, device_resident=(dm, dp))

        assert all(p == fib15_proof for p in vb.run_ranks(whole, ctxs))
        assert all(p == fib15_proof for p in vb.run_ranks(shards, ctxs))
    finally:
        _close(ctxs)


def test_split_commit_mixed_heights(ctx, oracle):
    """commit_batches / commit_shifted_batches with tall (split) and short (replicated) matrices in one tree."""
    import valida_b200 as vb

    rng = np.random.default_rng(5)
    mats = [rng.integers(0, P, (1 << 14, 5), dtype=np.uint32), rng.integers(0, P, (1 << 16, 3), dtype=np.uint32),
            rng.integers(0, P, (1 << 14, 11), dtype=np.uint32), rng.integers(0, P, (1 << 9, 2), dtype=np.uint32),
            rng.integers(0, P, (1, 7), dtype=np.uint32), rng.integers(0, P, (2, 2), dtype=np.uint32)]
    shifts = [1, 31 * 31 % P, 5, 1, 1, 7]
    ref_root = oracle.commit_batches(mats)
    root1, pd1 = vb.TwoAdicFriPcs(ctx).commit_shifted_batches(mats, shifts)
    pd1.free()
    ctxs, _ = _group(4, oracle)
    try:
        def go(r, c):
            pcs = vb.TwoAdicFriPcs(c)
            root, pd = pcs.commit_batches(mats)
            roots, pds = pcs.commit_shifted_batches(mats, shifts)
            pd.free(); pds.free()
            return root, roots

        for root, roots in vb.run_ranks(go, ctxs):
            assert np.array_equal(root, ref_root)
            assert np.array_equal(roots, root1)
    finally:
        _close(ctxs)


def test_split_prove_multichip_program(ctx, oracle):
    """config 5's shape (add, sub, lt family, and/or/xor chips carry rows; several tall chips of different heights)."""
    import valida_b200 as vb
    from programs import config5_program

    t = vb.run_program(config5_program(2000), initial_fp=0x1000)
    assert t.main[0].shape[0] == 1 << 15
    single = vb.prove_machine(vb.StarkConfig(ctx, oracle.rc480), t)
    assert oracle.verify(single, t.preprocessed) == 0
    ctxs, cfgs = _group(2, oracle)
    try:
        assert all(p == single for p in vb.run_ranks(lambda r, c: vb.prove_machine(cfgs[r], t), ctxs))
    finally:
        _close(ctxs)


def test_sharding_off_behaves_as_lone_gpu(oracle, fib15, fib15_proof):
    import valida_b200 as vb

    ctxs, cfgs = _group(2, oracle)
    try:
        for c in ctxs:
            c.set_sharding(False)
        assert vb.prove_machine(cfgs[0], fib15) == fib15_proof      # no collective: one rank alone may call
    finally:
        _close(ctxs)
