"""``from models.residual import ResidualLayer, ResidualStack`` -- reference models/residual.py:8-51."""
from vqvae_b200.modules import ResidualLayer, ResidualStack  # noqa: F401
