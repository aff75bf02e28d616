"""Helper of tests/test_reference_binding.py: run in a fresh interpreter so that module aliasing cannot leak.

    python tests/_binding_probe.py reference <out.pt>     build the UNMODIFIED reference VAEs, save their state_dicts
    python tests/_binding_probe.py aliased <in.pt> <out.pt>  alias the four hot-path modules as INTEGRATION.md section 1 says,
                                                          import the reference's own experiments/vae.py on top of them, build
                                                          the same VAEs, load the reference checkpoints, save the aliased ones
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # BASELINE configs[0]: toy SO(3) VAE, MLP decoder
    "config1": dict(latent_mode="so3", decoder_mode="mlp", mean_mode="alg", encode_mode="toy", deconv_mode="toy", degrees=3, rep_copies=3),
    # BASELINE configs[3]: conv encoder, action decoder, deconv stack (config 4; experiments/main.py:155,167 defaults)
    "config4": dict(latent_mode="so3", decoder_mode="action", mean_mode="s2s2", encode_mode="conv", deconv_mode="deconv", degrees=6,
                    rep_copies=10, deconv_hidden=200, rgb=True, batch_norm=True),
    "config4_alg_fixed_sigma": dict(latent_mode="so3", decoder_mode="action", mean_mode="alg", encode_mode="conv", deconv_mode="deconv",
                                    degrees=6, rep_copies=10, deconv_hidden=50, rgb=False, batch_norm=False, fixed_sigma=0.3),
}


def build_all(VAE):
    torch.manual_seed(0)
    out = {}
    for name, kw in CONFIGS.items():
        model = VAE(**kw)
        model.r_callback = None                      # read in encode() but never assigned by the reference
        out[name] = model
    return out


def main():
    mode = sys.argv[1]
    import refshim
    if mode == "reference":
        refshim.load_reference()                      # stand-ins for the absent third-party packages only
        from lie_vae.experiments.vae import VAE
        import lie_vae.reparameterize as rp
        assert rp.__file__.startswith(refshim.REFERENCE_ROOT)
        models = build_all(VAE)
        torch.save({k: m.state_dict() for k, m in models.items()}, sys.argv[2])
        return
    # ---- INTEGRATION.md section 1, verbatim
    import lie_vae_b200.lie_tools, lie_vae_b200.reparameterize, lie_vae_b200.decoders, lie_vae_b200.utils
    sys.modules["lie_vae.lie_tools"] = lie_vae_b200.lie_tools
    sys.modules["lie_vae.utils"] = lie_vae_b200.utils
    sys.modules["lie_vae.reparameterize"] = lie_vae_b200.reparameterize
    sys.modules["lie_vae.decoders"] = lie_vae_b200.decoders
    # ----
    sys.path.insert(0, refshim.REFERENCE_ROOT)
    from lie_vae.experiments.vae import VAE           # the reference's own model assembly, unedited
    import lie_vae.experiments.vae as vae_mod
    assert vae_mod.__file__.startswith(refshim.REFERENCE_ROOT)
    assert vae_mod.SO3reparameterize is lie_vae_b200.reparameterize.SO3reparameterize
    assert vae_mod.ActionNet is lie_vae_b200.decoders.ActionNet
    assert vae_mod.group_matrix_to_eazyz is lie_vae_b200.lie_tools.group_matrix_to_eazyz
    # the regularisers (losses/*.py) are callers of lie_tools only: unedited, they pick up the kernels too
    import lie_vae.losses.equivariance_loss as eq
    import lie_vae.losses.encoder_continuity_loss as ec
    assert eq.__file__.startswith(refshim.REFERENCE_ROOT) and eq.s2s1rodrigues is lie_vae_b200.lie_tools.s2s1rodrigues
    assert ec.EncoderContinuityLoss(None, lamb=2.0)(torch.arange(12.0).view(4, 3), 0).item() == 2.0 * 27.0
    models = build_all(VAE)
    ref_sd = torch.load(sys.argv[2])
    report = {}
    for name, model in models.items():
        sd = model.state_dict()
        report[name] = {"keys_equal": list(sd.keys()) == list(ref_sd[name].keys()),
                        "shapes_equal": all(tuple(sd[k].shape) == tuple(ref_sd[name][k].shape) and sd[k].dtype == ref_sd[name][k].dtype
                                            for k in sd if k in ref_sd[name]),
                        "n_params": sum(p.numel() for p in model.parameters())}
        model.load_state_dict(ref_sd[name], strict=True)            # reference checkpoint -> aliased model
        report[name]["loaded_equal"] = all(torch.equal(model.state_dict()[k], ref_sd[name][k]) for k in ref_sd[name])
    torch.manual_seed(1)
    fresh = build_all(VAE)                                          # different weights, to be loaded back into the reference
    torch.save({"report": report, "state": {k: m.state_dict() for k, m in fresh.items()}}, sys.argv[3])


if __name__ == "__main__":
    main()
