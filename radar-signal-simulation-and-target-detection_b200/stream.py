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


class AsyncDetectionGather:
    """The path's one exchange step without stalling the compute stream: the detection ring slots of a finished batch are
    packed into ONE message per rank -- per CPI a 16-byte header carrying the count, then the first ``cap`` records -- and
    all-gathered on a side stream behind an event, so the collective of batch k overlaps the kernels of batch k + 1.

    ``counts_u8`` / ``recs_u8`` are uint8 views of the context's device ring (RadarChain.stream_device_buffers):
    [slots * 4] and [slots, max_detections * 40].  Two message buffers alternate, so a batch may be launched while the
    previous one is still in flight; ``wait`` returns, for every rank, the per-CPI tables in the reference order and raises
    OverflowError when a CPI found more detections than ``cap`` (never truncates).  With one rank no collective is issued.
    """
    HEADER = 16

    def __init__(self, counts_u8, recs_u8, batch_slots: int, cap: int):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.cap, self.batch = int(cap), int(batch_slots)
        self.counts = counts_u8.view(-1, 4)
        self.recs = recs_u8
        dev = recs_u8.device
        self.msg_bytes = self.HEADER + self.cap * REC_BYTES
        self.send = [torch.zeros((self.batch, self.msg_bytes), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.recv = [torch.empty((self.world, self.batch, self.msg_bytes), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.ready = [None, None]
        self.done = [None, None]
        self.turn = 0

    def launch(self, first_slot: int, compute_stream=None) -> int:
        """Gather slots [first_slot, first_slot + batch) once everything enqueued so far on ``compute_stream`` has run."""
        torch = self.torch
        k = self.turn
        self.turn ^= 1

        def body():
            send = self.send[k]
            send[:, :4] = self.counts[first_slot:first_slot + self.batch]
            send[:, self.HEADER:] = self.recs[first_slot:first_slot + self.batch, : self.cap * REC_BYTES]
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.recv[k].view(self.world * self.batch, self.msg_bytes), send)
            else:
                self.recv[k][0].copy_(send)
        if self.side is not None:
            if self.done[k] is not None:
                self.side.wait_event(self.done[k])              # buffer k's previous message has been delivered
            ev = torch.cuda.Event()
            ev.record(compute_stream if compute_stream is not None else torch.cuda.current_stream())
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                body()
                self.done[k] = torch.cuda.Event()
                self.done[k].record(self.side)
        else:
            body()
        return k

    def fence(self, k: int, compute_stream):
        """Make ``compute_stream`` wait until message k has left the ring (call before its slots are reused)."""
        if self.side is not None and self.done[k] is not None:
            compute_stream.wait_event(self.done[k])

    def wait(self, k: int):
        if self.side is not None and self.done[k] is not None:
            self.done[k].synchronize()
        buf = self.recv[k].cpu().numpy()
        out = []
        for r in range(self.world):
            per_rank = []
            for s in range(self.batch):
                n = int(buf[r, s, :4].view(np.int32)[0])
                if n > self.cap:
                    raise OverflowError(f"rank {r} slot {s}: {n} detections exceed the gather capacity {self.cap}")
                d = np.frombuffer(buf[r, s, self.HEADER:self.HEADER + n * REC_BYTES].tobytes(), dtype=DETECTION_DTYPE)
                per_rank.append(sort_detections(d) if n else d)
            out.append(per_rank)
        return out
