"""CPU suite: the N>1 path of the batch sharder with world_size-2 gloo process groups, plus the
CompressBatch worker-loop semantics the reference tests (fennec_test.go:844-934)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fennec_b200 import batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = batch.shard_range(n_items, world, rank)
        # each shard "scores" its own items: score(i) = i + 0.25, a stand-in for per-item SSIM
        local = torch.arange(b, e, dtype=torch.float64) + 0.25
        full = batch.gather_scores(local, n_items, world, rank)
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [10, 7, 2, 1])
def test_two_rank_gather_keeps_input_order(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [i + 0.25 for i in range(n_items)]
    assert got[0] == want and got[1] == want


def test_empty_batch_returns_nothing():  # fennec_test.go:844-849
    assert batch.run_sharded([], lambda x: x) == []


def test_results_keep_order_index_and_progress():  # fennec_test.go:851-900
    items = [f"item{i}" for i in range(5)]
    seen = []
    res = batch.run_sharded(items, lambda s: s.upper(), on_item=lambda c, t: seen.append((c, t)))
    assert [r.index for r in res] == list(range(5))
    assert [r.result for r in res] == [s.upper() for s in items]
    assert all(r.err is None and r.item == items[i] for i, r in enumerate(res))
    assert seen == [(i + 1, 5) for i in range(5)]


def test_pre_cancelled_context_marks_every_item():  # fennec_test.go:902-920
    res = batch.run_sharded(["a", "b"], lambda s: s, cancelled=lambda: True)
    assert all(r.err is not None and r.result is None for r in res)


def test_one_bad_item_does_not_stop_the_batch():  # fennec_test.go:922-934
    def work(s):
        if s == "bad":
            raise FileNotFoundError(s)
        return s
    res = batch.run_sharded(["ok", "bad", "ok2"], work)
    assert res[0].err is None and isinstance(res[1].err, FileNotFoundError) and res[2].result == "ok2"


def test_shards_leave_foreign_items_for_the_gather():
    items = list(range(10))
    r0 = batch.run_sharded(items, lambda x: x * 2, rank=0, world=2)
    r1 = batch.run_sharded(items, lambda x: x * 2, rank=1, world=2)
    merged = [a or b for a, b in zip(r0, r1)]
    assert [m.result for m in merged] == [x * 2 for x in items]
    assert sum(x is not None for x in r0) == 5 and sum(x is not None for x in r1) == 5
