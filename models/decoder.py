"""``from models.decoder import Decoder`` -- reference models/decoder.py:9-39."""
from vqvae_b200.modules import Decoder  # noqa: F401
from models.residual import ResidualStack  # noqa: F401
