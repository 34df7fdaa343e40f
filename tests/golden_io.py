"""Helpers to read the committed golden vectors (tests/golden/*.npz)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case(dict):
    __getattr__ = dict.__getitem__


def load(file_name):
    """Returns {case_name: Case(field -> torch tensor)} for one golden file."""
    z = np.load(os.path.join(GOLDEN_DIR, file_name))
    cases = {}
    for name in z["names"]:
        name = str(name)
        c = Case()
        for key in z.files:
            if key.startswith(name + "/"):
                arr = z[key]
                c[key[len(name) + 1:]] = torch.from_numpy(arr) if arr.shape != () else arr.item()
        cases[name] = c
    return cases


def opt(t):
    """Golden files store 'absent' tensors (no mask, no target lengths) as empty arrays."""
    return None if (t is None or t.numel() == 0) else t
