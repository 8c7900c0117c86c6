"""``torchdistx.fake``: tensors without storage.  A "fake" tensor here is a ``meta`` tensor."""
import contextlib

import torch


def is_fake(tensor: torch.Tensor) -> bool:
    return isinstance(tensor, torch.Tensor) and tensor.is_meta


@contextlib.contextmanager
def fake_mode(*, fake_cuda: bool = False):
    with torch.device("meta"):
        yield


def meta_like(fake: torch.Tensor) -> torch.Tensor:
    return torch.empty_like(fake, device="meta")
