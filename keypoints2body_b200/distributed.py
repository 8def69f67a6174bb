"""Frame sharding across ranks and the one-frame halo exchange of schedule S2.

One process per GPU (torchrun); frames are contiguous shards ``[r*F/G, (r+1)*F/G)``.  The fit
itself needs no collective -- frames are independent given their initialisation.  Schedule S2's
sweep 1 initialises frame t from frame t-1's sweep-0 result, so at a shard boundary rank r needs
the LAST sweep-0 frame of rank r-1: one packed row (<= 95 floats) per boundary, sent with
``torch.distributed`` point-to-point ops (NCCL over NVLink on GPUs, gloo in the CPU tests).
The reference has no distributed code at all.
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

HALO_KEYS = ("global_orient", "body_pose", "betas", "transl", "expression")
HALO_DIMS = {"global_orient": 3, "body_pose": 69, "betas": 10, "transl": 3, "expression": 10}


def shard_range(num_frames: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous shard of rank ``rank``; remainders go to the lowest ranks."""
    base, rem = divmod(num_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_halo(last: dict) -> torch.Tensor:
    """dict of (1,dim) tensors -> one (95,) row (absent expression packs as zeros)."""
    ref = last["body_pose"]
    parts = []
    for k in HALO_KEYS:
        v = last.get(k)
        parts.append(v.reshape(-1) if v is not None else torch.zeros(HALO_DIMS[k], dtype=ref.dtype, device=ref.device))
    return torch.cat(parts).contiguous()


def unpack_halo(row: torch.Tensor, with_expression: bool) -> dict:
    out, o = {}, 0
    for k in HALO_KEYS:
        d = HALO_DIMS[k]
        if k != "expression" or with_expression:
            out[k] = row[o:o + d].reshape(1, d)
        o += d
    return out


def exchange_halo(last: dict, group: Optional[dist.ProcessGroup] = None) -> Optional[dict]:
    """Send this rank's last sweep-0 frame to rank+1; return rank-1's (None on rank 0).

    Every rank calls this once between the two sweeps.  Without an initialised process group
    (single GPU) it returns None.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return None
    send = pack_halo(last)
    recv = torch.empty_like(send)
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, send, rank + 1, group))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv, rank - 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return unpack_halo(recv, last.get("expression") is not None) if rank > 0 else None


def plan_two_sweep(seq_ind: torch.Tensor, num_iters_first: int, num_iters_followup: int):
    """Per-frame budgets / flags of the two sweeps from the per-frame index inside its sequence.

    Returns (iters0, iters1, preserve1, starts): sweep 0 fits every frame with the first-frame
    budget and no temporal term; sweep 1 re-fits frames with ``seq_ind > 0`` with the follow-up
    budget and the temporal term, sequence starts (``seq_ind == 0``) keep their sweep-0 fit.
    """
    first = seq_ind == 0
    iters0 = torch.full_like(seq_ind, num_iters_first, dtype=torch.int32)
    iters1 = torch.where(first, 0, num_iters_followup).to(torch.int32)
    preserve1 = (~first).to(torch.uint8)
    starts = torch.nonzero(first).reshape(-1)
    return iters0, iters1, preserve1, starts
