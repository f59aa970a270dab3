"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : row-sharded Q4_K mat-vec, exchange fused into the kernel
(NVLink peer stores) vs the NCCL all-gather of the same slices; checks every rank ends with the identical full y."""
import os, sys
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ggml_b200 as g
from ggml_b200.parallel import shard_rows
from oracle import oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
orc = O.Oracle()
t, K = g.Q4_K, 4096
M_total = 1024 * world
rb = g.row_size(t, K)
rng = np.random.default_rng(7)
W = O.random_blocks(t, M_total * K // 256, rng)
X = rng.uniform(-1, 1, K).astype(np.float32)
shards = shard_rows(M_total, world, 16)
lo, hi = shards[rank]
Wd = torch.from_numpy(W[lo * rb:hi * rb]).cuda(); Xd = torch.from_numpy(X).cuda()
ex = g.PeerExchange(M_total, rank, world, lo)
Yl = torch.empty((1, 1, 1, hi - lo), device="cuda")
want = orc.mul_mat(t, W, X, M_total, 1, K)[0]
# both operating points of the kernel: dependent launch (8-warp CTAs) and independent launch (4-warp CTAs, what bench.py uses)
for fl in (0, g.MM_SRC0_STATIC | g.MM_SRC1_STATIC):
    a = g.mul_mat_args(t, Wd, Xd, Yl, hi - lo, 1, K, flags=fl)
    for it in range(3):
        ex.mul_mat_gather(a); ex.wait()
    torch.cuda.synchronize()
    y = ex.y_full().cpu().numpy()
    err = O.nmse(y, want)
    assert err < 1e-10, (fl, err)
# NCCL reference exchange
yl = g.mul_mat(t, Wd, Xd, hi - lo, 1, K).view(-1)
buf = torch.empty(M_total, device="cuda")
dist.all_gather_into_tensor(buf, yl)
same = bool(torch.equal(buf.cpu(), torch.from_numpy(y)))
print(f"rank {rank}/{world}: fused-gather nmse vs oracle {err:.2e}, identical to NCCL all-gather: {same}", flush=True)
assert err < 1e-10 and same
ex.close()
dist.destroy_process_group()
