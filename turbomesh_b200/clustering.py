"""1-D point distributions on [0, 1] -- host-side mirror of the reference's ``src/core/clustering.zig``.

These feed the clustering arrays ``s1, s2, t1, t2`` consumed by the TFI kernel
(``tfi.zig:112-208``).  Scalar Python floats are IEEE doubles and the expressions keep the
reference's operation order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Uniform:
    """``clustering.zig:9-17``."""

    def compute(self, n: int) -> np.ndarray:
        return np.array([i / (n - 1) for i in range(n)], dtype=np.float64)


@dataclass(frozen=True)
class Roberts:
    """Roberts stretching, ``clustering.zig:24-42``; alpha = 0.5 clusters at both ends."""

    alpha: float
    beta: float

    def compute(self, n: int) -> np.ndarray:
        assert n > 1
        alpha, beta = self.alpha, self.beta
        out = np.empty(n, dtype=np.float64)
        for i in range(n):
            u = i / (n - 1)
            tmp = math.pow((beta + 1.0) / (beta - 1.0), (u - alpha) / (1.0 - alpha))
            tbar = (beta + 2.0 * alpha) * tmp - beta + 2.0 * alpha
            out[i] = tbar / ((2.0 * alpha + 1.0) * (1.0 + tmp))
        return out


@dataclass(frozen=True)
class SingleHyperbolicClustering:
    """Vinokur one-sided tanh stretching, ``clustering.zig:56-95``; needs ``(n-1)*delta_s <= 1``."""

    delta_s: float

    def compute(self, n: int) -> np.ndarray:
        n_1 = float(n - 1)
        b = n_1 * self.delta_s
        y = 1.0 / b
        if y < 1.0:
            raise ValueError("SingleHyperbolicClustering needs (n-1)*delta_s <= 1 (clustering.zig:68-76)")
        if y < 2.7829681:
            y_bar = y - 1.0
            delta = math.sqrt(6.0 * y_bar) * (
                1.0 + y_bar * (-0.15 + y_bar * (0.057321429 + y_bar * (-0.024907295 + y_bar * (0.0077424461 - 0.0010794123 * y_bar))))
            )
        else:
            w = 1.0 / y - 0.028527431
            v = math.log(y)
            delta = v + (1.0 + 1.0 / v) * math.log(2.0 * v) - 0.02041793 + w * (0.24902722 + w * (1.9496443 + w * (-2.6294547 + 8.56795911 * w)))
        out = np.array([i / n_1 for i in range(n)], dtype=np.float64)
        for i in range(1, n):
            out[i] = 1.0 + math.tanh(0.5 * delta * (out[i] - 1.0)) / math.tanh(0.5 * delta)
        assert out[0] == 0.0 and out[-1] == 1.0
        return out


def from_json(obj: dict):
    """Tagged-union form of ``clustering.Function`` (``clustering.zig:97-108``) as std.json writes it."""
    (tag, val), = obj.items()
    if tag == "uniform":
        return Uniform()
    if tag == "roberts":
        return Roberts(float(val["alpha"]), float(val["beta"]))
    if tag == "single_hyperbolic_clustering":
        return SingleHyperbolicClustering(float(val["delta_s"]))
    raise ValueError(f"unknown clustering function {tag!r}")


def create(function, n: int) -> np.ndarray:
    """``clustering.create``, ``clustering.zig:110-116``."""
    return function.compute(n)
