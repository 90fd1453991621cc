"""N > 1 host logic on CPU (gloo, world_size 2): each rank owns an independent trace (a proof segment);
the job metric is sum(rows) / max-over-ranks time; ranks agree on it."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import valida_b200 as vb

    # rank r proves the segment fib(n_r): distinct work per rank, no data-path collective
    n = [25, 582][rank]
    t = vb.run_program(vb.fib_program(n), initial_fp=0x1000)
    rows = t.main[0].shape[0]
    ms = [40.0, 100.0][rank]                      # pretend per-rank device times
    value, tmax = bench.aggregate_throughput(dist, rows, ms)
    gathered = [None, None]
    dist.all_gather_object(gathered, (rank, rows, t.clock, t.mem_cell(0x1004)))
    out[rank] = (value, tmax, gathered)
    dist.destroy_process_group()


def test_two_rank_aggregation_gloo(built):
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    v0, v1 = out[0], out[1]
    assert v0[0] == v1[0] and v0[1] == v1[1] == 100.0            # max over ranks
    assert abs(v0[0] - (256 + 4096) / 0.1) < 1e-6                 # whole-job rows / max time
    assert v0[2] == v1[2] == [(0, 256, 192, 75025), (1, 4096, 17 + 7 * 582, v0[2][1][3])]
