"""Sequence-parallel (Ulysses-style) plumbing for the denoising step: one process per GPU, torch.distributed (NCCL over
NVLink) for the exchanges.  No reference counterpart — the reference is strictly single-GPU; the
oracle for this path is "same result as one GPU".

Sharding: rank r owns rows [r*N/P, (r+1)*N/P) of the concatenated [text; video] token axis and every weight.
Per DiT layer two all-to-alls: (1) fused q|k|v, sequence-shard -> head-shard, (2) attention output, head-shard ->
sequence-shard.  The fused-QKV GEMM writes buffer (1) in its final layout and the out-projection consumes buffer (2) as
a K-blocked operand (`include/bya.h`: col_block / a_kblock), so no pack or unpack kernel runs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class RowShard:
    """This rank's slice of the [text; video] rows."""

    rows: int        # local rows R = N / P
    first: int       # global index of the first local row
    text_rows: int   # local rows that are text tokens (they come first)
    video_rows: int
    video_first: int  # global video-token index of the first local video row


def shard_rows(n_tokens: int, text_len: int, world: int, rank: int) -> RowShard:
    if n_tokens % world:
        raise RuntimeError(f"bya_b200: {n_tokens} tokens are not divisible by the sequence-parallel size {world}")
    R = n_tokens // world
    n0 = rank * R
    tl = min(max(text_len - n0, 0), R)
    return RowShard(R, n0, tl, R - tl, max(n0 - text_len, 0))


def qkv_rows_by_destination(w_qkv: torch.Tensor, dim: int, world: int) -> torch.Tensor:
    """Reorders the rows of the fused [q; k; v] projection ([3*dim, K] or [3*dim]) to [dest rank][q|k|v][heads of dest]."""
    dl = dim // world
    return torch.cat([w_qkv[t * dim + r * dl: t * dim + (r + 1) * dl] for r in range(world) for t in range(3)], 0).contiguous()


# ---- sequence-parallel router layout (engine.RouterShard / run_router_sp): rank r owns the spatial positions
# [r*hwl, (r+1)*hwl), hwl = ceil(hw / world), of every (character, frame); positions >= hw are padding
def router_local_tokens(frames: int, hw: int, world: int, rank: int, device=None) -> torch.Tensor:
    """Global video-token index (f*hw + position) of each local (frame, local position) row; padding rows repeat the
    last real position of the frame."""
    hwl = (hw + world - 1) // world
    j = (torch.arange(hwl, device=device) + rank * hwl).clamp(max=hw - 1)
    f = torch.arange(frames, device=device)
    return (f[:, None] * hw + j[None, :]).reshape(-1)


def router_gather_positions(recv: torch.Tensor, cf: int, world: int, hwl: int) -> torch.Tensor:
    """All-to-all result [src][(c,f), src's positions][cols] -> view ordered [(c,f)][all positions (padded)][cols]."""
    return recv.view(world, cf, hwl, -1).permute(1, 0, 2, 3)


def router_scatter_positions(full: torch.Tensor, cf: int, world: int, hwl: int) -> torch.Tensor:
    """[(c,f)][all positions (padded)][cols] -> view ordered [dest][(c,f), dest's positions][cols] (all-to-all input)."""
    return full.view(cf, world, hwl, -1).permute(1, 0, 2, 3)


def enable(model, group=None, cfg_parallel: bool = False):
    """Turns on multi-GPU execution of `model`.

    cfg_parallel=False: Ulysses sequence parallelism over `group` (default: WORLD).
    cfg_parallel=True : the two classifier-free-guidance branches (batch elements 0 / 1 of every forward input; they are
        independent inside the reference's forward — transformer.py:779, :870 loop over them) run batch-parallel: the
        first half of WORLD takes element 0, the second half element 1, each half sequence-parallel inside; the two
        predictions are exchanged once per step (pipeline_bindyouravatar.py:932-933 needs both).  SURVEY.md §8e."""
    import torch.distributed as dist

    if not cfg_parallel:
        model._sp_group = group if group is not None else dist.group.WORLD
        model._cfg = None
        model.invalidate()
        return model
    world, rank = dist.get_world_size(), dist.get_rank()
    if group is not None or world % 2:
        raise RuntimeError("bya_b200: cfg_parallel splits WORLD in two halves (even world size, no sub-group)")
    half = world // 2
    halves = [dist.new_group(list(range(h * half, (h + 1) * half))) for h in range(2)]   # every rank creates both
    model._cfg = dict(branch=rank // half, half=half, world=world)
    model._sp_group = halves[rank // half] if half > 1 else None
    model.invalidate()
    return model


def cfg_slice(x, branch: int):
    """Batch element `branch` of a forward input (tensor, or nested list / tuple of tensors), keeping the batch axis."""
    if isinstance(x, (list, tuple)):
        return type(x)(cfg_slice(y, branch) for y in x)
    if torch.is_tensor(x) and x.ndim > 0 and x.shape[0] == 2:
        return x[branch:branch + 1]
    return x


def cfg_gather(out: torch.Tensor, cfg: dict) -> torch.Tensor:
    """[1, ...] prediction of this rank's branch -> [2, ...] on every rank (one all-gather over WORLD; 2.2 MB per rank)."""
    import torch.distributed as dist

    allo = torch.empty((cfg["world"],) + tuple(out.shape[1:]), device=out.device, dtype=out.dtype)
    dist.all_gather_into_tensor(allo, out.contiguous())
    return torch.stack([allo[0], allo[cfg["half"]]])


# ---- pure-torch statements of the two exchanges (any backend; used by the gloo tests as the layout specification)
def exchange_qkv(qkv_send: torch.Tensor, group=None) -> torch.Tensor:
    """qkv_send [P, R, 3*Dl] ([dest][local row][q|k|v heads of dest]) -> [P*R, 3*Dl]: every row, this rank's heads."""
    import torch.distributed as dist

    out = torch.empty_like(qkv_send)
    dist.all_to_all_single(out, qkv_send.contiguous(), group=group)
    return out.view(-1, qkv_send.shape[-1])


def exchange_out(o_send: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """o_send [N, Dl] (every row, this rank's heads) -> [P, R, Dl]: local rows, K-blocked by source rank."""
    import torch.distributed as dist

    R = o_send.shape[0] // world
    out = torch.empty(world, R, o_send.shape[1], dtype=o_send.dtype, device=o_send.device)
    dist.all_to_all_single(out, o_send.contiguous().view(world, R, -1), group=group)
    return out
