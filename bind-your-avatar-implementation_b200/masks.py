"""Tracking-mask files -> forced routing logits on the GPU.

Drop-in for `util/utils.py:853-936` of the reference (`process_single_mask_dir`, `process_masks_to_routing_logits`): the
SAM-2 stage writes one PNG per frame and character under `<dir>/1`, `<dir>/2` (`annotated_frame_%05d.png`,
tools/sam2_tools.py:150-165); the reference then binarises, trilinearly resizes to the 13x30x45 latent grid on the CPU,
thresholds at 0.5 and builds one-hot routing logits.  Here the PNGs are decoded on the host (that part is I/O), and
binarise -> resize -> threshold -> label -> one-hot runs in ONE kernel (`bya_masks_to_routing`), bit-exact with the
reference (tests: the reference's own outputs on four mask sets)."""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch

from . import ops


def load_tracking_masks(base_dir: str, chars: int = 2) -> torch.Tensor:
    """`<base_dir>/<c+1>/annotated_frame_%05d.png` -> uint8 [C, T, H, W] (0 / 1), on the CPU.
    Same file discovery as the reference: the number of frames is the number of .png files of the directory
    (util/utils.py:855-861); a missing sub-directory raises the reference's ValueError (:878-879)."""
    from PIL import Image

    dirs = [os.path.join(base_dir, str(c + 1)) for c in range(chars)]
    if not all(os.path.exists(d) for d in dirs):
        names = " and ".join(f"'{c + 1}'" for c in range(chars))
        raise ValueError(f"both subdirectories {names} must exist in {base_dir}")
    out = []
    for d in dirs:
        n = len([f for f in os.listdir(d) if f.endswith(".png")])
        frames = []
        for t in range(n):
            a = np.array(Image.open(os.path.join(d, f"annotated_frame_{t:05d}.png")))
            if a.ndim != 2:
                raise ValueError(f"{d}: expected single-channel mask PNGs, got an array of shape {a.shape}")
            frames.append((a > 0).astype(np.uint8))
        out.append(np.stack(frames))
    if len({o.shape for o in out}) != 1:
        raise ValueError(f"{base_dir}: the characters' mask sequences differ in shape: {[o.shape for o in out]}")
    return torch.from_numpy(np.stack(out))


def process_masks_to_routing_logits(base_dir: str, frames: int = 13, grid: Tuple[int, int] = (30, 45), chars: int = 2,
                                    device="cuda") -> torch.Tensor:
    """Reference signature (`process_masks_to_routing_logits(base_dir)` -> float32 [1, frames*gh*gw, chars]) with the
    geometry the reference hard-codes as defaults (T=13, 60/2 x 90/2; util/utils.py:886-890).  The result stays on
    `device`: it is what `forward(routing_logits_forcing=...)` consumes."""
    masks = load_tracking_masks(base_dir, chars).to(device)
    _, logits = ops.masks_to_routing(masks, frames, grid[0], grid[1])
    return logits[None]
