"""``from models.vqvae import VQVAE`` -- reference models/vqvae.py:10-44."""
from vqvae_b200.modules import VQVAE  # noqa: F401
from models.encoder import Encoder  # noqa: F401
from models.quantizer import VectorQuantizer  # noqa: F401
from models.decoder import Decoder  # noqa: F401
