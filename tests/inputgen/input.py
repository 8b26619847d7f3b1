"""JSON input + profile loading -- host-side mirror of ``src/core/input.zig`` and ``src/core/csv.zig``."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from turbomesh_b200 import clustering as cluster

from .geometry import Geometry, Profile
from .templates import O4H, NumCells


def parse_csv_into_vec2d(path: str) -> np.ndarray:
    """``csv.parseCsvIntoVec2d``, ``csv.zig:10-57``: space separated ``x y`` rows, ``#`` comments."""
    rows = []
    with open(path, "r") as f:
        for line in f.read().split("\n"):
            if not line:
                continue
            if line[0] == "#":
                continue
            parts = [p for p in line.split(" ") if p]
            if len(parts) != 2:
                raise ValueError(f"csv parsing error in {path}: {line!r}")
            rows.append((float(parts[0]), float(parts[1])))
    return np.array(rows, dtype=np.float64)


def _read_side(path: str) -> np.ndarray:
    """``input.readSide``, ``input.zig:100-108``: reversed if x is decreasing."""
    side = parse_csv_into_vec2d(path)
    if side[0, 0] > side[-1, 0]:
        side = side[::-1].copy()
    return side


def create_profile(profile_input: dict, scale: float = 1.0, base_dir: str = ".") -> Profile:
    """``input.create_profile``, ``input.zig:43-90``."""
    (tag, val), = profile_input.items()
    if tag == "data":
        down = np.array(val["down"], dtype=np.float64)
        up = np.array(val["up"], dtype=np.float64)
    elif tag == "csv":
        down = _read_side(os.path.join(base_dir, val["down_csv_path"]))
        up = _read_side(os.path.join(base_dir, val["up_csv_path"]))
    else:
        raise ValueError(f"unknown profile input {tag!r}")
    if scale != 1.0:
        down = down * scale
        up = up * scale
    return Profile(down, up)


@dataclass
class SmoothingInput:
    iterations: int
    solver: dict
    wall_control_function: dict


@dataclass
class Input:
    """``input.Input``, ``input.zig:25-41``."""

    template: O4H
    smoothing: SmoothingInput
    scale: float
    pitch: float
    profile: dict
    output: Optional[str] = None
    gui: Optional[bool] = None

    @staticmethod
    def from_json(text: str) -> "Input":
        obj = json.loads(text)
        (ttag, tval), = obj["template"].items()
        if ttag != "O4H":
            raise ValueError(f"unknown template {ttag!r}")
        template = O4H(
            blade_clustering=cluster.from_json(tval["blade_clustering"]),
            num_cells=NumCells(**{k: int(v) for k, v in tval["num_cells"].items()}),
            inlet_distance=tval.get("inlet_distance"),
            outlet_distance=tval.get("outlet_distance"),
        )
        sm = obj["smoothing"]
        smoothing = SmoothingInput(int(sm.get("iterations", 0)), sm["solver"], sm.get("wall_control_function", {"laplace": {}}))
        geo = obj["geometry"]
        return Input(template, smoothing, float(geo.get("scale", 1.0)), float(geo["pitch"]), geo["profile"], obj.get("output"), obj.get("gui"))

    def geometry(self, base_dir: str = ".") -> Geometry:
        """``gui/main.zig:43-45``: the pitch is scaled like the coordinates."""
        return Geometry(self.scale * self.pitch, create_profile(self.profile, self.scale, base_dir))
