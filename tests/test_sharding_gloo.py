"""The N>1 path on CPU: world_size-2 `gloo` processes shard the table exactly as bench.py does (one
contiguous row range per rank of a generator keyed by global row index), aggregate their shard with the
CPU oracle, exchange partials and merge them the way kq_hashagg_merge_allreduce / _repartition_alltoall
do (SUM of SUMs, MIN of MINs, MAX of MAXs, SUM of COUNTs; hash(key) % world owners). The merged result
must equal the oracle on the whole table — the decomposition main() relies on (Main.kt:1309-1325)."""
import os
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATES = "ALAKAZARCACOCTDEFLGAHIIDILINIAKSKYLAMEMDMAMIMNMSMOMTNENVNHNJNMNYNCNDOHOKORPARISCSDTNTXUTVTVAWAWVWIWY"
SPECS = [dict(kind=5, col_id=0, dict=STATES, dict_width=2, null_per_10k=200), dict(kind=3, col_id=1, ilo=0, ihi=1000, null_per_10k=300)]
N_PER_RANK = 40_000


def partial(O, lo, hi):
    v = O.col(1)
    a = O.HashAggregate([O.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)])
    a.update(O.generate(SPECS, 3, lo, hi))
    return [tuple(r) for r in zip(*[c.to_pylist() for c in a.finalize().to_arrow()])]


def merge(rows):
    out = {}
    for k, s, mn, mx, c in rows:
        if k not in out:
            out[k] = [s, mn, mx, c]
            continue
        o = out[k]
        o[0] = s if o[0] is None else (o[0] if s is None else o[0] + s)
        o[1] = mn if o[1] is None else (o[1] if mn is None else min(o[1], mn))
        o[2] = mx if o[2] is None else (o[2] if mx is None else max(o[2], mx))
        o[3] += c
    return sorted(((k, *v) for k, v in out.items()), key=lambda t: (t[0] is not None, t[0] or ""))


def worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import oracle as O
    O.build()
    lo, hi = bench.shard_range(rank, N_PER_RANK)
    mine = partial(O, lo, hi)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)                     # low cardinality: everybody gets every partial
    full = merge([r for part in gathered for r in part])
    # high cardinality: every key has exactly one owner rank, owners merge what they are sent
    owner = lambda k: (hash(k) if k is None else sum(k.encode())) % world
    sent = [[r for r in mine if owner(r[0]) == dst] for dst in range(world)]
    boxes = [None] * world              # gloo has no object all-to-all: gather the per-destination lists instead
    dist.all_gather_object(boxes, sent)
    owned = merge([r for src in range(world) for r in boxes[src][rank]])
    q.put((rank, full, owned))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_and_merge_equals_whole_table():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=240) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    sys.path[:0] = [ROOT]
    from oracle import oracle as O
    O.build()
    want = merge(partial(O, 0, world * N_PER_RANK))
    for rank, full, owned in res:
        assert full == want, f"rank {rank}: merged partials differ from the whole-table aggregate"
    keys = [set(r[0] for r in owned) for _, _, owned in res]
    assert not (keys[0] & keys[1])
    assert merge([r for _, _, owned in res for r in owned]) == want
