"""INTEGRATION.md section 1 executed for real: the reference's own ``experiments/vae.py`` on top of this package's modules.

In fresh interpreters (tests/_binding_probe.py): (a) the unmodified reference builds the BASELINE configs[0] / configs[3]
VAEs and saves their checkpoints; (b) the four hot-path modules are aliased exactly as INTEGRATION.md says, the reference's
``VAE`` class is imported on top of them, builds the same models -- identical ``state_dict`` keys, shapes and dtypes -- and
loads the reference checkpoints; (c) the reference loads the aliased models' checkpoints.  CPU only (construction and
checkpoints; the kernels have no CPU path); skipped where ``/root/reference`` does not exist (the GPU box).
"""
import os
import subprocess
import sys

import pytest
import torch

import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "tests", "_binding_probe.py")

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")


def run(*args):
    p = subprocess.run([sys.executable, PROBE, *args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-4000:]


@pytest.mark.timeout(900)
def test_reference_vae_builds_on_aliased_modules_and_checkpoints_load_both_ways(tmp_path):
    ref_pt, ali_pt = str(tmp_path / "ref.pt"), str(tmp_path / "aliased.pt")
    run("reference", ref_pt)
    run("aliased", ref_pt, ali_pt)
    out = torch.load(ali_pt)
    for name, rep in out["report"].items():
        assert rep["keys_equal"] and rep["shapes_equal"] and rep["loaded_equal"], (name, rep)
    assert out["report"]["config1"]["n_params"] == 24124                 # SURVEY.md section 8d [probed]
    assert out["report"]["config4"]["n_params"] == 5247652
    # (c) aliased checkpoints -> the unmodified reference, in this process (the reference import is test infrastructure)
    refshim.load_reference()
    from lie_vae.experiments.vae import VAE
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _binding_probe as probe
    for name, kw in probe.CONFIGS.items():
        model = VAE(**kw)
        model.load_state_dict(out["state"][name], strict=True)
        assert all(torch.equal(model.state_dict()[k], v) for k, v in out["state"][name].items())
