"""torchrun --nproc-per-node N tests/gpu_multi_check.py (run by tests/test_gpu_multi.py): the row-sharded mat-vec whose exchange is fused into
the kernel (every lane stores its row into every rank's gathered y over NVLink, the last CTA publishes the epoch) against the CPU
oracle and against an NCCL all-gather of the same slices, on every rank.  Cases: both operating points of the kernel (dependent /
independent launches), several overlapping launches with ONE wait at the end (publication order), ragged shards, and
BASELINE.json configs[4] (Q4_K 8192 x 28672 split over the N GPUs)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ggml_b200 as g  # noqa: E402
from ggml_b200.parallel import row_granule, shard_rows  # noqa: E402
from oracle import oracle as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
orc = O.Oracle()
worst = 0.0


def check(t, M_total, K, slots, seed):
    """`slots` overlapping ops (distinct weights), one wait; every rank must end with the full, correct y of every op"""
    global worst
    rb = g.row_size(t, K)
    rng = np.random.default_rng(seed)                             # same data on every rank
    Ws = [O.random_blocks(t, M_total * K // orc.blck_size(t), rng) for _ in range(slots)]
    X = rng.uniform(-1, 1, K).astype(np.float32)
    shards = shard_rows(M_total, world, max(16, row_granule(rb)))
    lo, hi = shards[rank]
    Wd = [torch.from_numpy(w[lo * rb:hi * rb]).cuda() for w in Ws]
    Xd = torch.from_numpy(X).cuda()
    Yl = torch.empty((1, 1, 1, hi - lo), device="cuda")
    ex = g.PeerExchange(M_total, rank, world, lo, slots=slots)
    rows = np.sort(np.random.default_rng(seed + 1).choice(M_total, min(M_total, 256), replace=False))
    for fl in (0, g.MM_SRC0_STATIC, g.MM_SRC0_STATIC | g.MM_SRC1_STATIC):
        args = [g.mul_mat_args(t, Wd[s], Xd, Yl, hi - lo, 1, K, flags=fl) for s in range(slots)]
        for it in range(3):
            for s in range(slots):
                ex.y_full(s).zero_() if it == 0 else None
            dist.barrier()                                        # nobody zeroes a buffer a peer is already storing into
            for s in range(slots):
                ex.mul_mat_gather(args[s], slot=s)
            ex.wait()
            torch.cuda.synchronize()
            for s in range(slots):
                y = ex.y_full(s).cpu().numpy()
                Wsub = np.concatenate([Ws[s][r * rb:(r + 1) * rb] for r in rows])
                err = O.nmse(y[rows], orc.mul_mat(t, Wsub, X, len(rows), 1, K)[0])
                worst = max(worst, err)
                assert err < 1e-10 and np.isfinite(y).all(), (O.TYPE_NAMES[t], M_total, K, fl, it, s, err)
            dist.barrier()
    # NCCL reference exchange of the same kernel's slices: bit-identical
    yl = g.mul_mat(t, Wd[0], Xd, hi - lo, 1, K).view(-1)
    width = max(h - l for l, h in shards)
    pad = torch.zeros(width, device="cuda"); pad[: hi - lo] = yl
    buf = torch.empty(world * width, device="cuda")
    dist.all_gather_into_tensor(buf, pad)
    full = torch.cat([buf[q * width: q * width + (h - l)] for q, (l, h) in enumerate(shards)])
    assert torch.equal(full, ex.y_full(0)), "fused gather differs from the NCCL all-gather of the same slices"
    ex.close()


check(g.Q4_K, 1024 * world, 4096, slots=1, seed=7)
check(g.Q4_K, 11008 * world, 4096, slots=5, seed=8)               # the bench's weak-scaling shape, overlapping launches
check(g.Q6_K, 1000 * world + 24, 2048, slots=3, seed=9)           # ragged shards
check(g.Q8_0, 4096, 4096, slots=2, seed=10)
check(g.Q4_K, 28672, 8192, slots=2, seed=11)                      # BASELINE.json configs[4]
print(f"rank {rank}/{world}: fused NVLink gather OK on every case (worst NMSE vs oracle {worst:.2e}), identical to the NCCL all-gather", flush=True)
dist.destroy_process_group()
