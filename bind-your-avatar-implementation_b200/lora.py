"""LoRA checkpoints of the reference without peft.

The reference wraps the transformer with peft (`util/utils.py:1027-1049` `load_mixed_lora_weights`, rank-128 adapters on
`attn1.to_q` / `attn1.to_k` of every block, lora_alpha 128), loads the safetensors after three key rewrites, and later
calls `pipe.fuse_lora(lora_scale=1 / lora_rank)` (`infer.py:279`), after which the adapters are gone and only the
merged `to_q` / `to_k` weights matter.  With peft installed that flow works on the drop-in class unchanged (the engine
notices the in-place merge and repacks).  This module gives the same end state without peft / diffusers:

    transformer = load_mixed_lora_weights(transformer, lora_paths, lora_rank)     # same name and arguments
    fuse_lora(transformer, lora_scale=1 / lora_rank)                              # stands in for pipe.fuse_lora

Merge arithmetic is peft's (`LoraLayer.get_delta_weight` + `merge`): delta = (B @ A) * (lora_alpha / r * lora_scale) in
the weight's dtype (computed in fp32 when the weight is a bf16 / fp16 CPU tensor), then `weight += delta`.  peft is not
in the build image, so this is restated from its published behaviour — parity unpinned; the test checks the algebra.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

import torch

_KEY = re.compile(r"transformer_blocks\.(\d+)\.attn1\.(to_q|to_k)\.lora_(A|B)(?:\.default)?\.weight$")


def parse_lora_state(state: Dict[str, torch.Tensor]) -> Tuple[Dict[Tuple[int, str], Dict[str, torch.Tensor]], List[str]]:
    """Groups a LoRA state dict by (block, projection).  Accepts every spelling the reference's rewrites accept
    (`transformer.module.`, `transformer.`, `base_model.model.` prefixes; with or without `.default`).  Returns the
    groups and the keys that are not attn1.to_q / to_k adapters (the reference reports those as unexpected keys)."""
    groups: Dict[Tuple[int, str], Dict[str, torch.Tensor]] = {}
    unexpected = []
    for k, v in state.items():
        m = _KEY.search(k)
        if m is None:
            unexpected.append(k)
            continue
        groups.setdefault((int(m.group(1)), m.group(2)), {})[m.group(3)] = v
    return groups, unexpected


def load_mixed_lora_weights(transformer, lora_paths, lora_rank: int = 128, log_file_path: Optional[str] = None,
                            lora_alpha: float = 128.0):
    """Reads every file of `lora_paths` and parks its adapters on the transformer until `fuse_lora`."""
    from safetensors.torch import load_file

    pending = getattr(transformer, "_pending_lora", [])
    n_blocks = len(transformer.transformer_blocks)
    for path in lora_paths:
        groups, unexpected = parse_lora_state(load_file(path))
        if not groups:
            raise ValueError(f"{path}: no attn1.to_q / attn1.to_k LoRA tensors found")
        for (blk, proj), ab in groups.items():
            if blk >= n_blocks:
                raise ValueError(f"{path}: adapter for block {blk}, the model has {n_blocks}")
            if set(ab) != {"A", "B"}:
                raise ValueError(f"{path}: block {blk} {proj} needs both lora_A and lora_B")
            w = getattr(transformer.transformer_blocks[blk].attn1, proj).weight
            a, b = ab["A"], ab["B"]
            if a.shape != (lora_rank, w.shape[1]) or b.shape != (w.shape[0], lora_rank):
                raise ValueError(f"{path}: block {blk} {proj}: lora_A {tuple(a.shape)} / lora_B {tuple(b.shape)} do not "
                                 f"match rank {lora_rank} on a {tuple(w.shape)} weight")
        missing = 2 * n_blocks - len(groups)
        msg = (f"--------------------------------start loading lora from safetensors--------------------------------\n"
               f"unexpected_keys: {unexpected}\nmissing adapters: {missing}\n"
               f"--------------------------------complete loading lora from safetensors--------------------------------\n")
        if log_file_path:
            with open(log_file_path, "a") as f:
                f.write(msg)
        pending.append(dict(groups=groups, rank=lora_rank, alpha=float(lora_alpha), path=path))
    transformer._pending_lora = pending
    return transformer


@torch.no_grad()
def fuse_lora(transformer, lora_scale: float = 1.0) -> int:
    """Merges every parked adapter into `attn1.to_q / to_k` in place and drops it; returns the number of merged
    projections.  The in-place update bumps the parameter versions, so the packed engine is rebuilt on the next step."""
    pending = getattr(transformer, "_pending_lora", [])
    merged = 0
    for ad in pending:
        scaling = ad["alpha"] / ad["rank"] * lora_scale
        for (blk, proj), ab in ad["groups"].items():
            w = getattr(transformer.transformer_blocks[blk].attn1, proj).weight
            a, b = ab["A"].to(w.device, w.dtype), ab["B"].to(w.device, w.dtype)
            to_fp32 = w.device.type == "cpu" and w.dtype in (torch.float16, torch.bfloat16)
            if to_fp32:
                a, b = a.float(), b.float()
            delta = (b @ a) * scaling
            w.add_(delta.to(w.dtype))
            merged += 1
    transformer._pending_lora = []
    if hasattr(transformer, "invalidate"):
        transformer.invalidate()
    return merged
