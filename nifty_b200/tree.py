"""Flat-vector view of latent parameter dictionaries.

``nifty.re`` works on pytrees (``tree_math.Vector``); JAX orders dict leaves by sorted key, which
fixes the ``random_like`` sub-key assignment (tree_math/forest_math.py:60-72).  The B200 path keeps
every latent vector as ONE flat device buffer in that same order so that conjugate gradient runs
fused vector kernels; ``Layout`` converts between the two views.
"""

from __future__ import annotations

from typing import Dict

import numpy as np
import torch


class Layout:
    def __init__(self, domain: Dict[str, tuple]):
        self.keys = sorted(domain)
        self.shapes = {k: tuple(int(s) for s in domain[k]) for k in self.keys}
        self.offsets, off = {}, 0
        for k in self.keys:
            self.offsets[k] = off
            off += int(np.prod(self.shapes[k], dtype=np.int64))
        self.size = off

    def numel(self, k):
        return int(np.prod(self.shapes[k], dtype=np.int64))

    def pack(self, tree, dtype, device) -> torch.Tensor:
        out = torch.empty(self.size, dtype=dtype, device=device)
        for k in self.keys:
            v = tree[k]
            v = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
            if tuple(v.shape) != self.shapes[k] and v.numel() != self.numel(k):
                raise ValueError(f"leaf {k!r}: shape {tuple(v.shape)} does not match {self.shapes[k]}")
            out[self.offsets[k]:self.offsets[k] + self.numel(k)] = v.reshape(-1).to(device=device, dtype=dtype)
        return out

    def unpack(self, vec: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {k: vec[self.offsets[k]:self.offsets[k] + self.numel(k)].reshape(self.shapes[k]) for k in self.keys}

    def unpack_numpy(self, vec: np.ndarray) -> Dict[str, np.ndarray]:
        return {k: np.asarray(vec[self.offsets[k]:self.offsets[k] + self.numel(k)]).reshape(self.shapes[k]) for k in self.keys}

    def pack_numpy(self, tree) -> np.ndarray:
        out = np.empty(self.size, dtype=np.result_type(*[np.asarray(tree[k]).dtype for k in self.keys]))
        for k in self.keys:
            out[self.offsets[k]:self.offsets[k] + self.numel(k)] = np.asarray(tree[k]).reshape(-1)
        return out

    def random(self, seed_or_rng, dtype, device) -> torch.Tensor:
        """Leaf-by-leaf N(0,1) draws in sorted-key order from a numpy Generator."""
        rng = seed_or_rng if isinstance(seed_or_rng, np.random.Generator) else np.random.default_rng(seed_or_rng)
        tree = {k: rng.standard_normal(self.shapes[k]) for k in self.keys}
        return self.pack(tree, dtype, device)
