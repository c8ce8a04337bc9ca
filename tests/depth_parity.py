"""Full-depth numeric parity of the denoising step on the GPU (helper shared by `tests/test_gpu_depth.py` and
`tools/gpu_depth_report.py`; TEST INFRASTRUCTURE — it drives `oracle/restated.py`).

Three passes over the same seeded weights (bf16-rounded) and inputs, all on one B200:
  1. the product (`BindyouravatarTransformer3DModel.forward`, libbya.so kernels), taps kept on the device in bf16;
  2. the restated oracle evaluated by torch in bf16 (cuBLAS + SDPA) — the "reference as a user runs it today"
     (`infer.py:275` casts the whole transformer to bf16): its distance from fp32 is the NOISE FLOOR of the format;
  3. the restated oracle in fp32 (truth; TF32 off), whose taps are compared on the fly with both of the above.
Per tap: cosine and max-abs relative to the truth's abs-max, for ours and for torch-bf16.  Follows the layer loop of
/root/reference/models/transformer.py:727-936 (incl. the `i % cross_attn_interval` branch that re-uses the routing of
the previous cross-attention layer, :858-863).
"""
from __future__ import annotations

import dataclasses
import time

import torch


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def build_gpu_model(cfg, seed=0):
    import bya_b200  # noqa: F401
    from bya_b200.synth import fill_module
    from bya_b200.transformer import BindyouravatarTransformer3DModel

    with torch.device("meta"):
        m = BindyouravatarTransformer3DModel(**cfg.ctor_kwargs())
    m = m.to(torch.bfloat16).to_empty(device="cuda").eval()
    m.router.frames, m.router.height, m.router.width = cfg.frames, cfg.grid_w, cfg.grid_h
    m.router.pos_emb = m.router._create_positional_embedding().to("cuda", torch.bfloat16)
    fill_module(m, seed)
    return m


def cast_inputs(inp, dtype):
    o = dict(inp)
    for k in ("hidden_states", "encoder_hidden_states", "audio_embeds", "af_matrix", "routing_logits_forcing"):
        if o.get(k) is not None:
            o[k] = o[k].to(dtype)
    o["id_cond"] = [t.to(dtype) for t in inp["id_cond"]]
    o["id_vit_hidden"] = [[t.to(dtype) for t in l] for l in inp["id_vit_hidden"]]
    return o


def want_tap(name):
    return name.endswith(".video") or name.endswith(".router") or name in ("face_tokens", "audio_ctx", "temb")


@torch.no_grad()
def depth_parity(cfg, forced_masks=False, seed=1234, zero_uncond_audio=True, log=print):
    """Returns {"taps": {name: {cos, rel_max, cos_bf16, rel_max_bf16}}, "output": {...}, "seconds": {...}}."""
    from bya_b200.synth import make_inputs
    from oracle import restated

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = build_gpu_model(cfg)
    inp = make_inputs(cfg, seed, device="cuda", dtype=torch.bfloat16, forced_masks=forced_masks)
    if cfg.batch == 2 and zero_uncond_audio:
        inp["audio_embeds"][0] = 0   # pipeline_bindyouravatar.py:884
    sec = {}

    ours, bf = {}, {}
    t0 = time.perf_counter()
    out_p = m(**inp, taps=lambda k, v: ours.__setitem__(k, v.detach().to(torch.bfloat16).clone()) if want_tap(k) else None)[0]
    torch.cuda.synchronize()
    sec["product_eager_with_taps"] = time.perf_counter() - t0

    sd_bf = dict(m.state_dict())
    t0 = time.perf_counter()
    out_b = restated.step(sd_bf, cfg, **inp,
                          taps=lambda k, v: bf.__setitem__(k, v.detach().to(torch.bfloat16).clone()) if want_tap(k) else None)
    torch.cuda.synchronize()
    sec["torch_bf16"] = time.perf_counter() - t0

    sd32 = {k: v.float() for k, v in sd_bf.items()}
    del sd_bf
    rows = {}

    def compare(k, v):
        if not want_tap(k) or k not in ours:
            return
        ref = v.detach().float()
        a = ours.pop(k).float()
        if ref.numel() != a.numel():      # the product taps batch element 0 only
            ref = ref[0]
        scale = float(ref.abs().max()) + 1e-30
        a = a.reshape(ref.shape)
        r = {"cos": _cos(a, ref), "rel_max": float((a - ref).abs().max()) / scale}
        if k in bf:
            b = bf.pop(k).float()
            b = (b[0] if b.numel() != ref.numel() else b).reshape(ref.shape)
            r["cos_bf16"], r["rel_max_bf16"] = _cos(b, ref), float((b - ref).abs().max()) / scale
        rows[k] = r

    t0 = time.perf_counter()
    out_o = restated.step(sd32, cfg, **cast_inputs(inp, torch.float32), taps=compare)
    torch.cuda.synchronize()
    sec["oracle_fp32"] = time.perf_counter() - t0
    scale = float(out_o.abs().max())
    res = {"taps": rows, "seconds": sec,
           "output": {"cos": _cos(out_p.float(), out_o), "rel_max": float((out_p.float() - out_o).abs().max()) / scale,
                      "cos_bf16": _cos(out_b.float(), out_o), "rel_max_bf16": float((out_b.float() - out_o).abs().max()) / scale,
                      "abs_max": scale, "finite": bool(torch.isfinite(out_p.float()).all())},
           "config": dataclasses.asdict(cfg), "forced_masks": forced_masks}
    del m, sd32
    torch.cuda.empty_cache()
    return res


def format_report(res, every=1):
    lines = ["| tap | cos (ours) | rel max-abs (ours) | cos (torch bf16) | rel max-abs (torch bf16) |", "|---|---|---|---|---|"]
    for k, r in res["taps"].items():
        lines.append(f"| {k} | {r['cos']:.6f} | {r['rel_max']:.4f} | {r.get('cos_bf16', float('nan')):.6f} | "
                     f"{r.get('rel_max_bf16', float('nan')):.4f} |")
    o = res["output"]
    lines.append(f"| **noise prediction** | {o['cos']:.6f} | {o['rel_max']:.4f} | {o['cos_bf16']:.6f} | {o['rel_max_bf16']:.4f} |")
    return "\n".join(lines)
