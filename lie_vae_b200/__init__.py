"""lie_vae_b200 -- B200-native (sm_100a) SO(3) latent hot path of pimdh/lie-vae.

Sub-modules mirror the reference package: ``lie_tools``, ``reparameterize``,
``decoders``, ``utils``.  The compute lives in ``liblievae_sm100a.so`` (C ABI in
``include/lievae.h``), built in-tree by ``lie_vae_b200._build``; importing the
package does not load it, the first kernel call does, and fails loudly if it is
missing.
"""
__version__ = "0.1.0"
