"""Row-sharding plan for the exact scan across the GPUs of one box.

Host-side bookkeeping only: which contiguous block of rows a rank owns and how per-shard results
are laid out for the all-gather. The search and the merge themselves are CUDA
(`scn_search_flat_shard_dev`, `scn_merge_topk_dev`)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) held by `rank`: contiguous blocks of ceil(n/world) rows (the last may be short
    or empty). Contiguity is what makes `row_base + local row` a global row, so that ascending
    64-bit merge keys reproduce the single-GPU (distance, row) order exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def gather_shape(world: int, nq: int, k: int) -> Tuple[int, int, int]:
    """Layout of the all-gathered key / id tensors expected by scn_merge_topk_dev: [G][nq][k]."""
    return (world, nq, k)
