"""N > 1 host logic on CPU (gloo, world_size 2): each rank owns an independent trace (a proof segment);
the job metric is sum(rows) / max-over-ranks time; ranks agree on it."""
import os
import sys

import torch
import torch.distributed as dist
import pytest
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


def _brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def _split_worker(rank, world, port, out):
    """The split commit on the host, with the library's own planning functions, a stand-in hash and gloo for the exchanges:
    columns go to the ranks the water-filling plan names, every rank "extends" its columns (here: a per-column permutation of the
    rows into committed order), ONE all-to-all hands every rank its contiguous run of committed rows, the rank hashes its leaves and
    builds its sub-tree, and only the sub-root layer is gathered.  The root must equal the one a single rank computes from the
    whole matrices; and the ranks must agree on the plan."""
    import hashlib

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    import valida_b200 as vb

    shapes = [(64, 5), (256, 3), (64, 7)]                       # two heights in one tree: the 64-row matrices are injected two layers up
    plan = vb.split_column_plan(world, shapes)
    plans = [None] * world
    dist.all_gather_object(plans, plan)

    rng = np.random.default_rng(3)
    mats = [rng.integers(0, 1 << 30, s, dtype=np.int64) for s in shapes]          # every rank generates the same "traces"

    def h(*parts):
        return hashlib.blake2s(b"".join(parts)).digest()

    def committed(m):                                           # stand-in for LDE + bit-reversed rows: height doubles
        lg = (2 * m.shape[0]).bit_length() - 1
        ext = np.concatenate([m, m * 3 + 1])
        return ext[[_brev(i, lg) for i in range(ext.shape[0])]]

    # single-rank reference tree
    ldes = [committed(m) for m in mats]
    max_h = max(l.shape[0] for l in ldes)

    def rows_hash(height, j, mats_at):
        return h(*[l[j].tobytes() for l in mats_at])

    def build(first, count, leaf_fn, inject_fn, height):       # nodes [first, first + count) of the layer with `height` nodes, down to one node per run
        layer = {j: leaf_fn(j) for j in range(first, first + count)}
        layers = [layer]
        while count > 1:
            height //= 2; first //= 2; count //= 2
            prev = layers[-1]
            nxt = {}
            for j in range(first, first + count):
                d = h(prev[2 * j], prev[2 * j + 1])
                inj = inject_fn(height, j)
                nxt[j] = h(d, inj) if inj is not None else d
            layers.append(nxt)
        return layers

    tall = [l for l in ldes if l.shape[0] == max_h]
    inject_whole = lambda height, j: rows_hash(height, j, [l for l in ldes if l.shape[0] == height]) if any(l.shape[0] == height for l in ldes) else None
    whole = build(0, max_h, lambda j: rows_hash(max_h, j, tall), inject_whole, max_h)
    root_ref = whole[-1][0]

    # ---- the split path ----
    # (1) my columns of every matrix (the plan), (2) "extend" them, (3) all-to-all: rank d gets its run of committed rows of my columns
    shards = []
    for m, pl in zip(mats, plan):
        c0, c1 = pl[rank], pl[rank + 1]
        ext = committed(m[:, c0:c1])                             # committed order is a row permutation: it commutes with taking columns
        hs = ext.shape[0] // world
        send = [ext[d * hs:(d + 1) * hs].copy() for d in range(world)]
        recv = [None] * world
        dist.all_to_all_object(recv, send) if hasattr(dist, "all_to_all_object") else None
        if recv[0] is None:                                      # portable fall-back: gather everything, keep mine
            allp = [None] * world
            dist.all_gather_object(allp, send)
            recv = [allp[src][rank] for src in range(world)]
        shards.append(np.concatenate(recv, axis=1))             # my rows x all columns (sources are in column order)
    # every shard equals the matching rows of the whole committed matrix
    ok_shards = all(np.array_equal(sh, l[rank * (l.shape[0] // world):(rank + 1) * (l.shape[0] // world)]) for sh, l in zip(shards, ldes))
    # (4) leaves and sub-tree over my run of nodes, injecting my rows of the shorter matrices
    def local_rows_hash(height, j):
        parts = []
        for sh, l in zip(shards, ldes):
            if l.shape[0] == height:
                parts.append(sh[j - rank * (height // world)].tobytes())
        return h(*parts) if parts else None
    b, c, split = vb.tree_share(max_h, world, rank)
    assert split and (b, c) == (rank * max_h // world, max_h // world)
    mine = build(b, c, lambda j: local_rows_hash(max_h, j), lambda height, j: local_rows_hash(height, j), max_h)
    # (5) the sub-roots meet; the top of the tree is computed by every rank
    subroots = [None] * world
    dist.all_gather_object(subroots, mine[-1][rank])
    top = {j: subroots[j] for j in range(world)}
    height = world
    while height > 1:
        height //= 2
        top = {j: h(top[2 * j], top[2 * j + 1]) if inject_whole(height, j) is None else h(h(top[2 * j], top[2 * j + 1]), inject_whole(height, j)) for j in range(height)}
    out[rank] = (plans, ok_shards, top[0] == root_ref)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_split_commit_plan_gloo(built, world):
    # 4 ranks over matrices of 5 / 3 / 7 columns: some ranks extend no column of some matrix (empty shares in the exchange)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000) + 7 * world
    mp.spawn(_split_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        plans, ok_shards, same_root = out[rank]
        assert all(pl == plans[0] for pl in plans)                 # the ranks agree on who extends what
        assert ok_shards and same_root


def test_column_plan_balances_a_commit(built):
    """Water-filling over a whole commit: contiguous ranges that tile every width, and no rank carries more than one column
    of the tallest matrix above the mean (an even split of every matrix on its own is 29 - 60 % above it at 8 ranks)."""
    import valida_b200 as vb

    for shapes in ([(1 << 22, 51), (1 << 24, 14), (1 << 22, 16)], [(1 << 22, 25), (1 << 24, 10), (1 << 22, 30)], [(1 << 22, 10), (1 << 24, 10), (1 << 22, 10)],
                   [(1 << 20, 79), (1 << 20, 45), (1 << 22, 51), (1 << 24, 14), (1 << 21, 16)]):
        for world in (2, 4, 8, 16):
            plan = vb.split_column_plan(world, shapes)
            load = [0] * world
            for (hh, w), pl in zip(shapes, plan):
                assert pl[0] == 0 and pl[-1] == w and all(pl[r] <= pl[r + 1] for r in range(world))
                for r in range(world):
                    load[r] += (pl[r + 1] - pl[r]) * hh
            mean = sum(load) / world
            assert max(load) <= mean + max(hh for hh, _ in shapes), (shapes, world, load)


def test_next_rows_of_a_shard_live_on_one_rank():
    """The quotient sweep reads natural row i + 2 next to row i (machine/src/quotient.rs:124-197).  A rank holds a contiguous run of the
    COMMITTED (bit-reversed) rows; this is the host formula of quotient.cu for the one rank that holds all its "next" rows."""
    for lg_g in (1, 2, 3, 4):
        g = 1 << lg_g
        for lg_h in (lg_g + 2, lg_g + 5):
            hh = 1 << lg_h                                          # LDE height
            per = hh // g
            for r in range(g):
                owners = {_brev((_brev(s, lg_h) + 2) % hh, lg_h) // per for s in range(r * per, (r + 1) * per)}
                rho = _brev(r, lg_g)
                assert owners == {_brev((rho + 2) % g, lg_g)}, (g, hh, r, owners)
