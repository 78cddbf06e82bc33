"""CPI-stream sharding across GPUs: one process per GPU, CPIs are independent (fun_process_single_frame
keeps no state between calls, fsf:13-158), so ranks share nothing on the data path.  The one exchange
step is the gather of the fixed-capacity detection rings at the end of a batch.

Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from ._abi import DETECTION_DTYPE
from .frame import sort_detections

REC_BYTES = np.dtype(DETECTION_DTYPE).itemsize      # 40


def shard_range(n_cpi: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the CPI index space owned by ``rank`` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    base, extra = divmod(n_cpi, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_detections(lists: List[np.ndarray], slots: int, cap: int):
    """Per-CPI detection tables -> (counts[slots] int32, records[slots, cap*40] uint8), the layout of the
    device ring.  A list longer than ``cap`` is an error (never truncated)."""
    import torch
    counts = torch.zeros(slots, dtype=torch.int32)
    recs = torch.zeros((slots, cap * REC_BYTES), dtype=torch.uint8)
    for i, d in enumerate(lists):
        if len(d) > cap:
            raise OverflowError(f"CPI {i}: {len(d)} detections exceed the gather capacity {cap}")
        counts[i] = len(d)
        if len(d):
            raw = np.ascontiguousarray(d, dtype=DETECTION_DTYPE).view(np.uint8).reshape(-1)
            recs[i, : raw.size] = torch.from_numpy(raw.copy())
    return counts, recs


def gather_detections(counts, recs, n_local: List[int] = None):
    """all_gather the ring tensors; returns on every rank a list (one entry per rank) of per-CPI
    detection tables in the reference order.  ``counts``: int32 [slots]; ``recs``: uint8 [slots, cap*40]
    (device tensors for NCCL, CPU tensors for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        all_counts, all_recs = counts.unsqueeze(0), recs.unsqueeze(0)
    else:
        # concatenated output layout: accepted by both NCCL and gloo
        flat_c = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
        flat_r = torch.empty((world * recs.shape[0], recs.shape[1]), dtype=recs.dtype, device=recs.device)
        dist.all_gather_into_tensor(flat_c, counts.contiguous())
        dist.all_gather_into_tensor(flat_r, recs.contiguous())
        all_counts = flat_c.view(world, counts.shape[0])
        all_recs = flat_r.view(world, recs.shape[0], recs.shape[1])
    all_counts, all_recs = all_counts.cpu().numpy(), all_recs.cpu().numpy()
    cap = all_recs.shape[-1] // REC_BYTES
    out = []
    for r in range(world):
        per_rank = []
        for s in range(all_counts.shape[1]):
            n = int(all_counts[r, s])
            if n > cap:
                raise OverflowError(f"rank {r} slot {s}: {n} detections exceed the gather capacity {cap}")
            d = np.frombuffer(all_recs[r, s, : n * REC_BYTES].tobytes(), dtype=DETECTION_DTYPE)
            per_rank.append(sort_detections(d) if n else d)
        out.append(per_rank)
    return out
