"""Data-parallel sharding of the inference batch: images are independent, so rank r owns
the contiguous slice [r*B/W, (r+1)*B/W) with the full weights replicated; there is no collective
inside the decode loop and exactly ONE all-gather of a packed per-rank result buffer at the end
(NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous, balanced slices; the first (batch % world) ranks get one extra image."""
    base, rem = divmod(int(batch), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_results(tokens, confs, boxes=None, max_iou=None):
    """int32 tokens (b,T1), f32 confs (b,C) [, f32 boxes (b,N,4), f32 max_iou (b,N)] -> one int32 buffer (b, T1 + C + 4N + N):
    the token ids as they are, the float fields as their bit patterns (a view, no conversion in either direction)."""
    parts = [tokens.to(torch.int32), confs.to(torch.float32).contiguous().view(torch.int32)]
    if boxes is not None:
        parts.append(boxes.reshape(boxes.shape[0], -1).to(torch.float32).contiguous().view(torch.int32))
    if max_iou is not None:
        parts.append(max_iou.to(torch.float32).contiguous().view(torch.int32))
    return torch.cat(parts, dim=1).contiguous()


def unpack_results(buf, T1, C, N=0):
    tokens = buf[:, :T1]
    confs = buf[:, T1:T1 + C].contiguous().view(torch.float32)
    off = T1 + C
    boxes = buf[:, off:off + 4 * N].contiguous().view(torch.float32).reshape(buf.shape[0], N, 4) if N else None
    max_iou = buf[:, off + 4 * N:off + 5 * N].contiguous().view(torch.float32) if N else None
    return tokens, confs, boxes, max_iou


def all_gather_results(packed, batch, group=None):
    """ONE collective: gathers every rank's (b_r, F) buffer (padded to the largest shard) and returns the
    global (batch, F) buffer on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return packed
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-int(batch) // world)
    F = packed.shape[1]
    send = torch.zeros((per, F), dtype=packed.dtype, device=packed.device)
    send[:packed.shape[0]] = packed
    recv = torch.empty((world * per, F), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    rows = []
    for r in range(world):
        s, e = shard_bounds(batch, r, world)
        rows.append(recv[r * per:r * per + (e - s)])
    return torch.cat(rows, dim=0)
