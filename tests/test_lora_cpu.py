"""CPU: LoRA checkpoints without peft — key spellings the reference's rewrites accept
(util/utils.py:1031-1036), shape / completeness errors, merge algebra W' = W + (alpha / r * lora_scale) B A, and that the
merge invalidates the packed engine."""
import pytest
import torch


@pytest.fixture()
def tiny():
    import bya_b200  # noqa: F401
    from bya_b200.synth import fill_module
    from bya_b200.transformer import BindyouravatarTransformer3DModel as M

    m = M(num_attention_heads=2, attention_head_dim=64, in_channels=48, out_channels=16, num_layers=2, text_embed_dim=64,
          time_embed_dim=32, cross_attn_interval=1, is_train_face=False, is_train_audio=False,
          use_rotary_positional_embeddings=True).eval()
    fill_module(m, 1)
    return m


def _adapters(rank, dim, blocks, prefix, suffix, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for b in blocks:
        for proj in ("to_q", "to_k"):
            sd[f"{prefix}transformer_blocks.{b}.attn1.{proj}.lora_A{suffix}.weight"] = 0.1 * torch.randn(rank, dim, generator=g)
            sd[f"{prefix}transformer_blocks.{b}.attn1.{proj}.lora_B{suffix}.weight"] = 0.1 * torch.randn(dim, rank, generator=g)
    return sd


@pytest.mark.parametrize("prefix,suffix", [("transformer.module.", ""), ("transformer.", ""), ("base_model.model.", ".default"), ("", "")])
def test_load_and_fuse_matches_the_merge_formula(tiny, tmp_path, prefix, suffix):
    from safetensors.torch import save_file

    from bya_b200.lora import fuse_lora, load_mixed_lora_weights

    rank, dim = 8, 128
    sd = _adapters(rank, dim, [0, 1], prefix, suffix)
    sd["transformer.something_else.weight"] = torch.zeros(3)
    path = str(tmp_path / "lora.safetensors")
    save_file(sd, path)
    log = str(tmp_path / "log.txt")
    before = {(b, p): getattr(tiny.transformer_blocks[b].attn1, p).weight.clone() for b in (0, 1) for p in ("to_q", "to_k", "to_v")}
    sig = tiny._signature()
    out = load_mixed_lora_weights(tiny, [path], lora_rank=rank, log_file_path=log)
    assert out is tiny and len(tiny._pending_lora) == 1
    assert "something_else" in open(log).read()
    for k, w in before.items():   # loading alone changes nothing (the reference's adapters act only after the fuse, too)
        assert torch.equal(getattr(tiny.transformer_blocks[k[0]].attn1, k[1]).weight, w)
    assert fuse_lora(tiny, lora_scale=1 / rank) == 4
    scaling = 128.0 / rank / rank
    for (b, p), w in before.items():
        now = getattr(tiny.transformer_blocks[b].attn1, p).weight
        if p == "to_v":
            assert torch.equal(now, w)
            continue
        A = sd[f"{prefix}transformer_blocks.{b}.attn1.{p}.lora_A{suffix}.weight"]
        B = sd[f"{prefix}transformer_blocks.{b}.attn1.{p}.lora_B{suffix}.weight"]
        assert torch.equal(now, w + (B @ A) * scaling)
    assert tiny._pending_lora == [] and tiny._signature() != sig   # the engine will repack
    assert fuse_lora(tiny) == 0


def test_bf16_cpu_weights_merge_through_fp32_like_peft(tiny, tmp_path):
    from safetensors.torch import save_file

    from bya_b200.lora import fuse_lora, load_mixed_lora_weights

    tiny.to(torch.bfloat16)
    rank, dim = 4, 128
    sd = _adapters(rank, dim, [1], "transformer.", "")
    path = str(tmp_path / "l.safetensors")
    save_file(sd, path)
    w0 = tiny.transformer_blocks[1].attn1.to_k.weight.clone()
    load_mixed_lora_weights(tiny, [path], lora_rank=rank)
    fuse_lora(tiny, 0.5)
    A = sd["transformer.transformer_blocks.1.attn1.to_k.lora_A.weight"].bfloat16().float()
    B = sd["transformer.transformer_blocks.1.attn1.to_k.lora_B.weight"].bfloat16().float()
    want = w0 + ((B @ A) * (128.0 / rank * 0.5)).bfloat16()
    assert torch.equal(tiny.transformer_blocks[1].attn1.to_k.weight, want)


def test_bad_lora_files_are_rejected(tiny, tmp_path):
    from safetensors.torch import save_file

    from bya_b200.lora import load_mixed_lora_weights

    def write(sd):
        p = str(tmp_path / f"f{len(list(tmp_path.iterdir()))}.safetensors")
        save_file(sd, p)
        return p

    with pytest.raises(ValueError, match="no attn1"):
        load_mixed_lora_weights(tiny, [write({"x.weight": torch.zeros(1)})], 8)
    with pytest.raises(ValueError, match="block 5"):
        load_mixed_lora_weights(tiny, [write(_adapters(8, 128, [5], "transformer.", ""))], 8)
    with pytest.raises(ValueError, match="rank 16"):
        load_mixed_lora_weights(tiny, [write(_adapters(8, 128, [0], "transformer.", ""))], 16)
    half = {k: v for k, v in _adapters(8, 128, [0], "transformer.", "").items() if "lora_A" in k}
    with pytest.raises(ValueError, match="both"):
        load_mixed_lora_weights(tiny, [write(half)], 8)
    assert getattr(tiny, "_pending_lora", []) == []
