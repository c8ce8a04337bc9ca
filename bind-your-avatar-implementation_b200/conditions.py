"""Host-side preparation of the step-invariant conditions, as the reference does it before the denoising loop:
classifier-free-guidance batching (`models/pipeline_bindyouravatar.py:877-884`, `models/utils.py:630-657`) and the
audio-to-face assignment matrix (`models/utils.py:660-670`, `infer.py:148,164`).  A few small tensors per generation:
plain torch, no kernels."""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch


def _uncond_then_cond(t: torch.Tensor, zero_uncond: bool) -> torch.Tensor:
    return torch.cat([torch.zeros_like(t) if zero_uncond else t, t], dim=0)


def cfg_id_vit_hidden(id_vit_hidden, zero2cond_cfg_flag: bool = False):
    """list(characters) of list(5) of [B, 577, 1024] -> the same nesting with batch 2B ([uncond | cond])."""
    if id_vit_hidden is None:
        raise ValueError("id_vit_hidden is None")
    return [[_uncond_then_cond(t, zero2cond_cfg_flag) for t in per_char] for per_char in id_vit_hidden]


def cfg_id_cond(id_cond, zero2cond_cfg_flag: bool = False):
    """list(characters) of [B, 1280] -> batch 2B."""
    if id_cond is None:
        raise ValueError("id_cond is None")
    return [_uncond_then_cond(t, zero2cond_cfg_flag) for t in id_cond]


def cfg_af_matrix(af_matrix: Optional[torch.Tensor], zero2cond_cfg_flag: bool = False):
    """[B, C, C] -> [2B, C, C]: repeated, or zeros for the unconditional branch (pipeline :881-882)."""
    if af_matrix is None:
        return None
    return af_matrix.repeat(2, 1, 1) if not zero2cond_cfg_flag else _uncond_then_cond(af_matrix, True)


def cfg_audio_embeds(audio_embs: Optional[torch.Tensor], zero2cond_cfg_flag: bool = False):
    """The unconditional branch always gets silent (zero) audio, whatever the flag says (pipeline :883-884)."""
    if audio_embs is None:
        return None
    return _uncond_then_cond(audio_embs, True)


def get_af_matrix_infer(speaker_pos: Union[str, int, Sequence[int]], chars: int = 2) -> torch.Tensor:
    """af[a, c] = 1 when audio stream a drives character c.  "left" -> identity, "right" -> the swap (the reference's
    two cases); beyond the reference: an int rotates the streams by that many characters, a sequence gives the
    character of every stream explicitly (C > 2)."""
    if isinstance(speaker_pos, str):
        if speaker_pos == "left":
            return torch.eye(chars)
        if speaker_pos == "right":
            if chars != 2:
                raise ValueError('"right" is only defined for two characters; pass the stream -> character list')
            return 1 - torch.eye(2)
        raise ValueError("speaker is not left or right")
    if isinstance(speaker_pos, int):
        speaker_pos = [(a + speaker_pos) % chars for a in range(chars)]
    speaker_pos = list(speaker_pos)
    if sorted(speaker_pos) != list(range(chars)):
        raise ValueError(f"speaker_pos must be a permutation of 0..{chars - 1}")
    af = torch.zeros(chars, chars)
    for a, c in enumerate(speaker_pos):
        af[a, c] = 1.0
    return af


def prepare_cfg_conditions(id_cond, id_vit_hidden, audio_embs, af_matrix, do_classifier_free_guidance: bool = True,
                           zero2cond_cfg_flag: bool = False):
    """pipeline_bindyouravatar.py:877-884 in one call: returns (id_cond, id_vit_hidden, audio_embs, af_matrix)."""
    if not do_classifier_free_guidance:
        return id_cond, id_vit_hidden, audio_embs, af_matrix
    return (cfg_id_cond(id_cond, zero2cond_cfg_flag), cfg_id_vit_hidden(id_vit_hidden, zero2cond_cfg_flag),
            cfg_audio_embeds(audio_embs, zero2cond_cfg_flag), cfg_af_matrix(af_matrix, zero2cond_cfg_flag))
