"""CPU: `bya_b200.conditions` against the reference's own functions.  The reference module (`models/utils.py`) imports
cv2 / insightface at the top, so the three functions are lifted out of its source with `ast` and executed as they are
(nothing is copied into the repository); skipped where /root/reference is not mounted."""
import ast
import os

import pytest
import torch

REF = os.path.join(os.environ.get("BYA_REFERENCE_ROOT", "/root/reference"), "models", "utils.py")


def _reference_functions(*names):
    tree = ast.parse(open(REF).read())
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return [ns[n] for n in names]


def _conditions(B=1, C=2):
    g = torch.Generator().manual_seed(0)
    id_cond = [torch.randn(B, 1280, generator=g) for _ in range(C)]
    vit = [[torch.randn(B, 7, 16, generator=g) for _ in range(5)] for _ in range(C)]
    return id_cond, vit


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isfile(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("zero2cond", [False, True])
def test_cfg_batching_equals_reference(zero2cond):
    import bya_b200  # noqa: F401
    from bya_b200 import conditions as c

    r_vit, r_cond, r_af = _reference_functions("cfg_id_vit_hidden", "cfg_id_cond", "get_af_matrix_infer")
    id_cond, vit = _conditions()
    for a, b in zip(c.cfg_id_cond(id_cond, zero2cond), r_cond(id_cond, zero2cond)):
        assert torch.equal(a, b) and a.shape[0] == 2
    for la, lb in zip(c.cfg_id_vit_hidden(vit, zero2cond), r_vit(vit, zero2cond)):
        assert len(la) == len(lb) == 5
        for a, b in zip(la, lb):
            assert torch.equal(a, b)
    for pos in ("left", "right"):
        assert torch.equal(c.get_af_matrix_infer(pos), r_af(pos))
    with pytest.raises(ValueError):
        r_af("middle")
    with pytest.raises(ValueError):
        c.get_af_matrix_infer("middle")
    with pytest.raises(ValueError):
        c.cfg_id_cond(None)
    with pytest.raises(ValueError):
        c.cfg_id_vit_hidden(None)


def test_cfg_audio_and_af_follow_the_pipeline_lines():
    """pipeline_bindyouravatar.py:881-884: af repeats (or zeros | af), audio is always zeros | audio."""
    import bya_b200  # noqa: F401
    from bya_b200 import conditions as c

    af = c.get_af_matrix_infer("right").unsqueeze(0)
    audio = torch.randn(1, 2, 9, 12, 8)
    assert torch.equal(c.cfg_af_matrix(af), af.repeat(2, 1, 1))
    assert torch.equal(c.cfg_af_matrix(af, True), torch.cat([torch.zeros_like(af), af]))
    for flag in (False, True):
        assert torch.equal(c.cfg_audio_embeds(audio, flag), torch.cat([torch.zeros_like(audio), audio]))
    assert c.cfg_af_matrix(None) is None and c.cfg_audio_embeds(None) is None
    id_cond, vit = _conditions()
    out = c.prepare_cfg_conditions(id_cond, vit, audio, af, True, True)
    assert out[0][0].shape[0] == 2 and out[1][1][4].shape[0] == 2 and out[2].shape[0] == 2 and out[3].shape[0] == 2
    assert float(out[0][0][0].abs().max()) == 0.0 and float(out[3][0].abs().max()) == 0.0
    same = c.prepare_cfg_conditions(id_cond, vit, audio, af, False)
    assert same[0] is id_cond and same[2] is audio


def test_af_matrix_for_more_than_two_characters():
    import bya_b200  # noqa: F401
    from bya_b200.conditions import get_af_matrix_infer

    assert torch.equal(get_af_matrix_infer("left", 3), torch.eye(3))
    assert torch.equal(get_af_matrix_infer(1, 2), 1 - torch.eye(2))          # rotate by one == "right" for two
    af = get_af_matrix_infer([2, 0, 1], 3)
    assert af.sum(0).tolist() == [1, 1, 1] and af.sum(1).tolist() == [1, 1, 1] and af[0, 2] == 1 and af[1, 0] == 1
    with pytest.raises(ValueError):
        get_af_matrix_infer([0, 0, 1], 3)
    with pytest.raises(ValueError):
        get_af_matrix_infer("right", 3)
