"""``from models.quantizer import VectorQuantizer`` -- reference models/quantizer.py:10-76."""
import torch

from vqvae_b200.modules import VectorQuantizer  # noqa: F401

# the reference binds a module-global device at import time (quantizer.py:7);
# kept for callers that read it, unused by the kernels (buffers follow the input's device)
device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
