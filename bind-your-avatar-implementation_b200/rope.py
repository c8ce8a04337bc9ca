"""3-D rotary tables for the joint self-attention (what `pipeline_bindyouravatar.py:586-610` hands to the
transformer as `image_rotary_emb`).  Host-side, once per generation; plain torch."""
import torch


def _axis(dim: int, pos: torch.Tensor, theta: float = 10000.0):
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.outer(pos.float(), freqs)
    return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)


def rope_3d_tables(head_dim: int, frames: int, grid_h: int, grid_w: int):
    """(cos, sin), each [frames*grid_h*grid_w, head_dim] fp32; per-token layout [t | h | w] = head_dim/4, 3/8, 3/8,
    every frequency repeated twice (interleaved-pair rotation)."""
    dt, dh, dw = head_dim // 4, head_dim // 8 * 3, head_dim // 8 * 3
    ct, st = _axis(dt, torch.arange(frames))
    ch, sh = _axis(dh, torch.arange(grid_h))
    cw, sw = _axis(dw, torch.arange(grid_w))

    def mix(a, b, c):
        a = a[:, None, None, :].expand(frames, grid_h, grid_w, -1)
        b = b[None, :, None, :].expand(frames, grid_h, grid_w, -1)
        c = c[None, None, :, :].expand(frames, grid_h, grid_w, -1)
        return torch.cat([a, b, c], -1).reshape(frames * grid_h * grid_w, head_dim).contiguous()

    return mix(ct, ch, cw), mix(st, sh, sw)
