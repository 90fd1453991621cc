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


def _split_worker(rank, world, port, out):
    """The split-commit plan on the host (the library's own split functions, a stand-in hash, gloo for the exchange):
    column shares tile every width, and a tree whose layers are derived share-by-share and completed with one
    all-gather per split layer equals the tree built by a single rank."""
    import hashlib

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import valida_b200 as vb

    widths = [1, 2, 3, 14, 16, 51, 64]
    mine = [vb.shard_range(w, world, rank) for w in widths]
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)

    def node(a, b):
        return hashlib.blake2s(a + b).digest()

    n_leaves = 64
    leaves = [hashlib.blake2s(bytes([i])).digest() for i in range(n_leaves)]
    whole = [leaves]
    while len(whole[-1]) > 1:
        p = whole[-1]
        whole.append([node(p[2 * i], p[2 * i + 1]) for i in range(len(p) // 2)])

    layers, pending = [], []
    b, c, split = vb.tree_share(n_leaves, world, rank)
    layer = {i: leaves[i] for i in range(b, b + c)}
    layers.append(layer)
    if split:
        pending.append(0)
    length = n_leaves

    def complete():
        for li in pending:
            parts = [None] * world
            dist.all_gather_object(parts, layers[li])
            for part in parts:
                layers[li].update(part)
        pending.clear()

    while length > 1:
        nxt = length // 2
        b, c, split = vb.tree_share(nxt, world, rank)
        if not split:
            complete()
        prev = layers[-1]
        layers.append({i: node(prev[2 * i], prev[2 * i + 1]) for i in range(b, b + c)})   # KeyError = plan needs data it lacks
        if split:
            pending.append(len(layers) - 1)
        length = nxt
    complete()
    same = all([layers[k][i] for i in range(len(whole[k]))] == whole[k] for k in range(len(whole)))
    out[rank] = (everyone, same)
    dist.destroy_process_group()


def test_split_commit_plan_gloo(built):
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_split_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in range(2):
        everyone, same = out[rank]
        assert same
        for k, w in enumerate([1, 2, 3, 14, 16, 51, 64]):
            (b0, e0), (b1, e1) = everyone[0][k], everyone[1][k]
            assert b0 == 0 and e0 == b1 and e1 == w and abs((e0 - b0) - (e1 - b1)) <= 1
