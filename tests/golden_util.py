"""Helpers to read the committed golden fixtures (tests/golden/*.npz, made by oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def cfg_of(d, key="cfg"):
    return json.loads(str(d[key]))


def state_dict_of(d, prefix="sd."):
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in d.items() if k.startswith(prefix)}


def t(a):
    return torch.from_numpy(np.array(a).copy())
