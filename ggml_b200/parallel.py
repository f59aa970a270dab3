"""Row-sharding of one MUL_MAT across ranks (one process per GPU), SURVEY.md §8e option 1.

The path shards by OUTPUT rows: rank r owns rows [lo, hi) of W (every row is an independent dot product), the
activation vector(s) are replicated, and the only exchange is the gather of the y slices — the same partition as the
reference's split buffer (`get_row_split`, src/ggml-cuda/ggml-cuda.cu:729-742), which also gathers and never reduces.
Row boundaries are rounded to `granule` rows so that every shard starts on a 16-byte boundary whatever the block size
(the TMA mat-vec kernel's requirement) — the reference rounds to its kernel tile for the same reason.

Host-side logic only (torch.distributed for the plumbing: NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple


def row_granule(row_bytes: int) -> int:
    """smallest row count whose byte size is a multiple of 16"""
    g = 1
    while (g * row_bytes) % 16 != 0:
        g *= 2
    return g


def shard_rows(M: int, world: int, granule: int = 1) -> List[Tuple[int, int]]:
    """[lo, hi) per rank: contiguous, covering [0, M), boundaries multiples of `granule` (except the end), balanced to one granule."""
    assert M >= 0 and world >= 1 and granule >= 1
    units = (M + granule - 1) // granule
    out = []
    lo = 0
    for r in range(world):
        n = units // world + (1 if r < units % world else 0)
        hi = min(M, lo + n * granule)
        out.append((lo, hi))
        lo = hi
    assert out[-1][1] == M
    return out


def gather_rows(y_local, M: int, shards: List[Tuple[int, int]], group=None):
    """all-gather the per-rank y slices [N, hi-lo] into the full [N, M] on every rank (one collective).
    Slices may differ by one granule, so they are padded to the largest slice for the fixed-size collective."""
    import torch
    import torch.distributed as dist
    world = len(shards)
    n = y_local.shape[0]
    width = max(hi - lo for lo, hi in shards)
    pad = torch.zeros((n, width), dtype=y_local.dtype, device=y_local.device)
    pad[:, : y_local.shape[1]] = y_local
    buf = torch.empty((world, n, width), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(buf.view(world * n, width), pad, group=group)
    out = torch.empty((n, M), dtype=y_local.dtype, device=y_local.device)
    for r, (lo, hi) in enumerate(shards):
        out[:, lo:hi] = buf[r, :, : hi - lo]
    return out
