"""Checkpoint interchange with the reference.

The reference saves ``{'model': state_dict, 'results': {...}, 'hyperparameters': vars(args)}`` with ``torch.save``
(utils.py:106-115, called from main.py:90-93) and loads it in the notebook's ``load_model`` (visualization.ipynb cell 1):
``VQVAE(params['n_hiddens'], params['n_residual_hiddens'], params['n_residual_layers'], params['n_embeddings'],
params['embedding_dim'], params['beta'])`` followed by ``load_state_dict(data['model'])``.  ``load_checkpoint`` does
the same with the drop-in model and packs every conv weight into its kernel layouts right away (not at the first forward).
"""
import torch

from .modules import VQVAE

_HP_KEYS = ("n_hiddens", "n_residual_hiddens", "n_residual_layers", "n_embeddings", "embedding_dim", "beta")


def load_checkpoint(path, device="cuda"):
    """(model, data) from a reference ``.pth``; ``data`` is the loaded dict (``results`` and ``hyperparameters`` untouched)."""
    # the file holds python lists / numpy scalars next to the tensors (main.py:59-64, 81-84): not a weights-only pickle
    data = torch.load(path, map_location="cpu", weights_only=False)
    for k in ("model", "hyperparameters"):
        if k not in data:
            raise KeyError(f"{path}: not a reference checkpoint (missing {k!r}; expected the dict of utils.py:106-115)")
    hp = data["hyperparameters"]
    missing = [k for k in _HP_KEYS if k not in hp]
    if missing:
        raise KeyError(f"{path}: hyperparameters lack {missing}")
    model = VQVAE(hp["n_hiddens"], hp["n_residual_hiddens"], hp["n_residual_layers"], hp["n_embeddings"],
                  hp["embedding_dim"], hp["beta"])
    model.load_state_dict(data["model"])          # 23 keys incl. the aliased stack.1.* 
    model = model.to(device).eval()
    if torch.device(device).type == "cuda":
        with torch.cuda.device(torch.device(device)):
            model.repack()                        # conv weights -> kernel layouts now
    return model, data


def save_checkpoint(model, results, hyperparameters, path):
    """Write the reference's checkpoint format (utils.py:106-115) so its notebook / load_model can read the file."""
    torch.save({"model": model.state_dict(), "results": results, "hyperparameters": hyperparameters}, path)
