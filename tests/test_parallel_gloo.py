"""CPU, world_size 2 (gloo): the multi-GPU host logic — row partition + gather of output slices.
Each rank computes its row shard of a quantized MUL_MAT (with the CPU oracle standing in for the device kernel —
this test checks the partition/exchange, not the kernel) and the gathered result must equal the unsharded product."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_shard_rows_properties():
    from ggml_b200.parallel import row_granule, shard_rows
    assert row_granule(2304) == 1 and row_granule(210) == 8 and row_granule(630) == 8 and row_granule(34) == 8 and row_granule(18 * 3) == 8
    for M in (0, 1, 7, 16, 1000, 11008, 28672, 50257):
        for world in (1, 2, 3, 4, 8):
            for g in (1, 8, 16):
                s = shard_rows(M, world, g)
                assert s[0][0] == 0 and s[-1][1] == M and all(a[1] == b[0] for a, b in zip(s, s[1:]))
                assert all(lo % g == 0 for lo, hi in s if lo < M)
                sizes = [hi - lo for lo, hi in s]
                assert max(sizes) - min(sizes) <= 2 * g or M < world * g


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from ggml_b200.parallel import gather_rows, row_granule, shard_rows
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = O.Oracle()
        for t, M, N, K in [(O.Q4_K, 203, 1, 512), (O.Q6_K, 77, 3, 256), (O.Q4_0, 64, 2, 96)]:
            rng = np.random.default_rng(42)                       # same on every rank
            W = O.random_blocks(t, M * K // orc.blck_size(t), rng)
            X = rng.uniform(-1, 1, N * K).astype(np.float32)
            rb = orc.row_size(t, K)
            shards = shard_rows(M, world, row_granule(rb))
            lo, hi = shards[rank]
            y_loc = orc.mul_mat(t, W[lo * rb:hi * rb], X, hi - lo, N, K) if hi > lo else np.zeros((N, 0), np.float32)
            y = gather_rows(torch.from_numpy(np.ascontiguousarray(y_loc)), M, shards).numpy()
            want = orc.mul_mat(t, W, X, M, N, K)
            assert np.array_equal(y, want), (rank, t)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_row_shard_gather_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
