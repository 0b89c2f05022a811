"""``FlatTerrain`` (``src/jaxsim/terrain/terrain.py:66-124``): the only terrain the kernel supports."""

from __future__ import annotations

import dataclasses


@dataclasses.dataclass(frozen=True)
class FlatTerrain:
    _height: float = 0.0

    @staticmethod
    def build(height: float = 0.0) -> "FlatTerrain":
        return FlatTerrain(_height=float(height))

    def height(self, x=None, y=None) -> float:
        return self._height
