"""CPU, world_size 2, gloo: the multi-process host logic of the data-parallel path —
shard ranges tile the batch, the packed-weight broadcast replicates rank 0's blobs, and running a
shard equals the corresponding slice of the full-batch result (no cross-image dependence)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qnnpack_b200 import shard as S


def test_shard_ranges_tile_the_batch():
    for total in (0, 1, 7, 8, 4096, 32768, 32771):
        for world in (1, 2, 3, 4, 8):
            spans = [S.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import q8_oracle as O
        from tests import cases as CS, util as U

        co = O.COracle()
        case = CS.conv_case("shard_case", 6, 9, 9, 1, 16, 24, ks=(3, 3), pad=(1, 1, 1, 1))
        x, k, b, kw = U.conv_setup(case)
        # rank 0 owns the weights; the others start from zeros and receive them by broadcast
        kt = torch.from_numpy(k.copy() if rank == 0 else np.zeros_like(k))
        bt = torch.from_numpy(b.copy() if rank == 0 else np.zeros_like(b))
        moved = S.replicate_from_rank0([kt, bt])
        assert moved == k.nbytes + b.nbytes
        assert np.array_equal(kt.numpy(), k) and np.array_equal(bt.numpy(), b)
        # the form bench.py uses: a whole model's raw parameters in ONE broadcast, NumPy arrays updated in place
        k2 = k.copy() if rank == 0 else np.zeros_like(k)
        b2 = b.copy() if rank == 0 else np.zeros_like(b)
        assert S.replicate_params_from_rank0([k2, b2]) == k.nbytes + b.nbytes
        assert np.array_equal(k2, k) and np.array_equal(b2, b) and b2.dtype == np.int32
        lo, hi = S.shard_range(case["n"], world, rank)
        mine = co.convolution(np.ascontiguousarray(x[lo:hi]), kt.numpy(), bt.numpy(), **kw)
        full = co.convolution(x, k, b, **kw)
        ok = np.array_equal(mine, full[lo:hi])
        # gather the shards on rank 0 (optional all-gather of outputs; not part of the metric)
        outs = [torch.empty(S.shard_range(case["n"], world, r)[1] - S.shard_range(case["n"], world, r)[0], *mine.shape[1:],
                            dtype=torch.uint8) for r in range(world)]
        dist.all_gather(outs, torch.from_numpy(mine))
        ok = ok and np.array_equal(torch.cat(outs).numpy(), full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_broadcast_and_sharded_run():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(2))
    assert results == {0: True, 1: True}
