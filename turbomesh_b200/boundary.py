"""Block-boundary topology -- host-side mirror of ``src/core/boundary.zig``.

``Side`` / ``Range`` / ``Connection`` / ``Condition`` keep the reference's field names and
semantics (``boundary.zig:8-26, 119-123, 178-187``); they are flattened into the C-ABI structs
``tm_range`` / ``tm_connection`` / ``tm_condition`` of ``include/turbomesh_gpu.h``.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Optional, Tuple


class Side(enum.IntEnum):
    """``boundary.zig:8-13``.  i_min = line j=0 (indexed by i), i_max = j=nj-1, j_min = line i=0, j_max = i=ni-1."""

    i_min = 0
    i_max = 1
    j_min = 2
    j_max = 3


class ConditionTag(enum.IntEnum):
    """``boundary.zig:172-176``."""

    wall = 0
    inlet = 1
    outlet = 2


@dataclass(frozen=True)
class Range:
    """``boundary.zig:15-26``; ``start > end`` means the range is traversed backwards."""

    block: int
    side: Side
    start: int
    end: int

    def len(self) -> int:
        return abs(self.start - self.end) + 1

    def local_ids(self, size: Tuple[int, int]):
        """Block-local node ids along the range (``Range.iterate``, ``boundary.zig:28-62``)."""
        ni, nj = size
        if self.side == Side.i_min:
            base, inc = self.start * nj, nj
        elif self.side == Side.i_max:
            base, inc = self.start * nj + nj - 1, nj
        elif self.side == Side.j_min:
            base, inc = self.start, 1
        else:
            base, inc = (ni - 1) * nj + self.start, 1
        if self.start > self.end:
            inc = -inc
        return [base + k * inc for k in range(self.len())]


@dataclass(frozen=True)
class Connection:
    """``boundary.zig:119-128``; ``periodicity`` maps ``ranges[0]`` onto ``ranges[1]`` (x0 + p == x1)."""

    ranges: Tuple[Range, Range]
    periodicity: Optional[Tuple[float, float]] = None

    def len(self) -> int:
        assert self.ranges[0].len() == self.ranges[1].len()
        return self.ranges[0].len()


@dataclass(frozen=True)
class Condition:
    """``boundary.zig:178-181``."""

    range: Range
    kind: ConditionTag
