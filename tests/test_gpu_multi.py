"""One PROCESS per GPU (the torchrun launch: NCCL for the small collectives, CUDA IPC for the peer pointers of the symmetric
heap): commits and whole proofs split across 2 (and, when the box has them, 4) GPUs are bit-identical to the single-GPU
results.  Needs >= 2 GPUs; the same data path with threads as ranks runs on any box in tests/test_gpu_split_local.py."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2013265921


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_binding
    import valida_b200 as vb

    orc = oracle_binding.Oracle()
    ctx = vb.Context(rank)
    cfg = vb.StarkConfig(ctx, orc.rc480)
    res = {}
    # single-GPU results first (no communicator yet)
    rng = np.random.default_rng(11)
    mats = [rng.integers(0, P, (1 << 13, 5), dtype=np.uint32), rng.integers(0, P, (1 << 15, 3), dtype=np.uint32),
            rng.integers(0, P, (1 << 10, 1), dtype=np.uint32), rng.integers(0, P, (1, 7), dtype=np.uint32),
            rng.integers(0, P, (2, 2), dtype=np.uint32)]
    pcs = vb.TwoAdicFriPcs(ctx)
    root_single, pd = pcs.commit_batches(mats)
    ldes_single = [m.download() for m in pcs.get_ldes(pd)]
    pd.free()
    t = vb.run_program(vb.fib_program(((1 << 15) - 17) // 7), initial_fp=0x1000)     # cpu 2^15, memory 2^17 rows: split at 2 and 4 ranks
    proof_single = vb.prove_machine(cfg, t)

    ctx.comm_init_from_torch()
    root_split, pd = pcs.commit_batches(mats)
    res["root_equal"] = bool(np.array_equal(root_single, root_split))
    ok = True
    for m, ref in zip(pcs.get_ldes(pd), ldes_single):     # a tall matrix comes back as this rank's run of the committed rows
        r0, n = m.local_rows()
        got = m.download()
        ok = ok and np.array_equal(got[r0:r0 + n], ref[r0:r0 + n])      # both in committed (bit-reversed) row order
    res["ldes_equal"] = bool(ok)
    pd.free()
    res["root_oracle"] = bool(np.array_equal(root_split, orc.commit_batches(mats)))
    proof_split = vb.prove_machine(cfg, t)
    res["proof_equal"] = proof_split == proof_single
    res["verifies"] = orc.verify(proof_split, t.preprocessed) == 0
    # sharding off again: the rank behaves as a lone GPU
    ctx.set_sharding(False)
    res["replica_equal"] = vb.prove_machine(cfg, t) == proof_single
    gathered = [None] * world
    dist.all_gather_object(gathered, proof_split)
    res["ranks_agree"] = all(g == proof_split for g in gathered)
    out[rank] = res
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_split_commit_and_prove_processes(built, world):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    mgr = mp.Manager()
    out = mgr.dict()
    port = 33500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        assert all(out[rank].values()), (rank, dict(out[rank]))
