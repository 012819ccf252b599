"""``from models.encoder import Encoder`` -- reference models/encoder.py:9-43."""
from vqvae_b200.modules import Encoder  # noqa: F401
from models.residual import ResidualStack  # noqa: F401
