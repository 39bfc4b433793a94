"""Load the committed golden vectors (tests/golden/reference_vectors.pt)."""
import os
from types import SimpleNamespace

import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.pt")
_CACHE = {}


def golden():
    if "fx" not in _CACHE:
        _CACHE["fx"] = torch.load(_PATH, map_location="cpu", weights_only=False)
    return _CACHE["fx"]


def to_namespace(d):
    """Plain nested dict -> attribute access (what the reference's ConfigDict offers)."""
    if isinstance(d, dict):
        return SimpleNamespace(**{k: to_namespace(v) for k, v in d.items()})
    return d
