"""Synthetic weights / images: moved to vqvae_b200/synth.py (no oracle arithmetic in it); re-exported for the tests."""
from vqvae_b200.synth import make_images, make_state_dict, state_dict_shapes  # noqa: F401
