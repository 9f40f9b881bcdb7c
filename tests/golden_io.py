"""Loader for the fixtures minted by oracle/make_golden.py from the unmodified reference."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case:
    def __init__(self, name):
        self.name, self.meta, self.inp, self.out = name, {}, {}, {}

    def __repr__(self):
        return f"Case({self.name})"


def load(book):
    """-> {case_name: Case} with torch tensors."""
    z = np.load(os.path.join(GOLDEN_DIR, book + ".npz"))
    cases = {}
    for key in z.files:
        name, kind, *rest = key.split("|")
        c = cases.setdefault(name, Case(name))
        if kind == "meta":
            c.meta = json.loads(bytes(z[key]).decode())
        else:
            getattr(c, "inp" if kind == "in" else "out")[rest[0]] = torch.from_numpy(np.array(z[key]))
    return cases


def bits_equal(a, b):
    """Bit-exact float comparison (distinguishes -0.0/+0.0; all NaNs count as one value)."""
    a, b = a.contiguous().float(), b.contiguous().float()
    if a.shape != b.shape:
        return False
    an, bn = torch.isnan(a), torch.isnan(b)
    if not torch.equal(an, bn):
        return False
    ai = a.masked_fill(an, 0).view(torch.int32)
    bi = b.masked_fill(bn, 0).view(torch.int32)
    return torch.equal(ai, bi)


def first_mismatch(a, b):
    a, b = a.contiguous().float().flatten(), b.contiguous().float().flatten()
    an, bn = torch.isnan(a), torch.isnan(b)
    bad = (an != bn) | ((a.masked_fill(an, 0).view(torch.int32) != b.masked_fill(bn, 0).view(torch.int32)))
    idx = torch.nonzero(bad).flatten()
    if idx.numel() == 0:
        return "none"
    i = int(idx[0])
    return f"{idx.numel()} mismatches, first at {i}: {a[i].item()!r} vs {b[i].item()!r}"
