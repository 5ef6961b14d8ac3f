"""Import path `models.util` of the reference (models/util.py): `load_batch` runs on the GPU (transformer-lm_b200/batch.py ->
bpe_batch_windows_dev, same signature, same random draws); `save_checkpoint` / `load_checkpoint` are the reference's plain
torch.save / torch.load dictionaries (models/util.py:10-34), kept so that `from models.util import save_checkpoint,
load_checkpoint, load_batch` (the reference's train.py:17) works against this package.

Not a drop-in for device="cpu": load_batch gathers on a B200 and raises BpeError for any other device (there is no CPU path)."""
import torch

from transformer_lm_b200.batch import load_batch  # noqa: F401


def save_checkpoint(model, optimizer, iteration, out):
    """{"optimizer_state_dict", "model_state_dict", "iteration"} through torch.save (models/util.py:10-21)."""
    torch.save({"optimizer_state_dict": optimizer.state_dict(), "model_state_dict": model.state_dict(), "iteration": iteration}, out)


def load_checkpoint(src, model, optimizer):
    """Restores model (and optimizer, unless None) from a save_checkpoint file; returns the iteration (models/util.py:24-34)."""
    state = torch.load(src, map_location=torch.device("cpu"))
    model.load_state_dict(state["model_state_dict"])
    if optimizer is not None:
        optimizer.load_state_dict(state["optimizer_state_dict"])
    return state["iteration"]
