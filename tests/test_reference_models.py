"""The drop-in boundary at MODEL level: after `install()`, the reference's own
`model.py` builds `Transformer_Net_Cross_Attention`, `SwinTransformerV2`, `Func_Struct_Cross` and
`SwinFusion` with `main.py`'s default arguments on top of our modules, with the same state_dict
keys and shapes as the reference build, and loads a reference-built state_dict strictly.

Needs the reference checkout (this container); skipped where it is absent (the GPU box).  Each
mode runs in its own subprocess -- the two `modules.*` sets cannot share one interpreter.
"""
import importlib
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def probes(tmp_path_factory):
    d = tmp_path_factory.mktemp("refmodels")
    out = {}
    for mode in ("reference", "installed"):          # reference first: it writes the state_dicts the other loads
        js = d / f"{mode}.json"
        res = subprocess.run([sys.executable, os.path.join(HERE, "ref_harness.py"), mode, str(js), str(d)],
                             capture_output=True, text=True, cwd=str(d), timeout=900)
        assert res.returncode == 0, res.stderr[-3000:]
        out[mode] = json.loads(js.read_text())
    return out


@pytest.mark.parametrize("name", list(H.MODELS))
def test_model_builds_on_dropin_modules(probes, name):
    ref, ours = probes["reference"][name], probes["installed"][name]
    assert ours["keys"] == ref["keys"]                                   # same keys, same order, same shapes
    assert ours["strict_load"] == [[], []]                               # load_state_dict(strict=True) of a reference build
    # every attention module of the installed build is ours, none is the reference's
    att = [c for c in ours["attention_classes"] if "Attention" in c.split(".")[-1] and not c.startswith("model.")]
    assert att and all(c.startswith("multimodal_neuroimage_b200.modules.") for c in att), att
    assert len(att) == len([c for c in ref["attention_classes"] if not c.startswith("model.")])


def test_star_exports_cover_the_reference():
    """`model.py` takes everything it uses from `from modules.X import *` (model.py:15,18): every public name of a
    reference module must exist in ours (trunc_normal_, Upsample, UpsampleOneStep, np, F were missing)."""
    code = r"""
import importlib, json, sys
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.append(%r)
out = {}
for name in ["swin_v2_module", "swinfusion_module", "crossmodal_transformer", "multihead_attention", "position_embedding"]:
    ref = importlib.import_module("modules." + name)
    ours = importlib.import_module("multimodal_neuroimage_b200.modules." + name)
    out[name] = sorted(n for n in dir(ref) if not n.startswith("_") and not hasattr(ours, n))
print(json.dumps(out))
""" % (os.path.join(HERE, "golden", "_shims"), H.ROOT, H.REF)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    missing = json.loads(res.stdout.strip().splitlines()[-1])
    missing["multihead_attention"] = [n for n in missing["multihead_attention"] if n != "sys"]   # an unused import there
    assert all(not v for v in missing.values()), missing


def test_upsample_modules_match_reference_layout():
    sf = importlib.import_module("multimodal_neuroimage_b200.modules.swinfusion_module")
    up = sf.Upsample(4, 8)
    assert [type(m).__name__ for m in up] == ["Conv2d", "PixelShuffle", "Conv2d", "PixelShuffle"]
    assert up[0].weight.shape == (32, 8, 3, 3)
    assert [type(m).__name__ for m in sf.Upsample(3, 8)] == ["Conv2d", "PixelShuffle"]
    with pytest.raises(ValueError):
        sf.Upsample(5, 8)
    one = sf.UpsampleOneStep(2, 12, 1, input_resolution=(84, 84))
    assert one[0].weight.shape == (4, 12, 3, 3) and one.flops() == 84 * 84 * 12 * 3 * 9
