"""Shared helpers of the parity tests (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np

from oracle.make_golden import MODEL_CASES, VQ_CASES, make_vq_inputs  # noqa: F401
from oracle.weights import make_images, make_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HP_KEYS = ("h_dim", "res_h_dim", "n_res_layers", "n_embeddings", "embedding_dim")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
        out = {k: d[k] for k in d.files}
    out["case"] = json.loads(str(out["case"]))
    return out


def model_case_inputs(case):
    hp = {k: case[k] for k in HP_KEYS}
    sd = make_state_dict(seed=case["wseed"], codebook=case["codebook"],
                         codebook_scale=case["codebook_scale"], **hp)
    x = make_images(case["batch"], case["size"], case["xseed"])
    return hp, sd, x


def build_model(hp, sd, device="cuda"):
    """The product model (drop-in import path) loaded with a numpy state dict."""
    import torch
    from models.vqvae import VQVAE
    m = VQVAE(hp["h_dim"], hp["res_h_dim"], hp["n_res_layers"], hp["n_embeddings"],
              hp["embedding_dim"], 0.25)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return m.to(device).eval()
