"""Name-only stand-in for `trimesh` (mesh collisions are disabled by default in the reference)."""


class Trimesh:
    def __init__(self, *a, **k):
        raise NotImplementedError("trimesh is not available in the stand-in")


def load(*a, **k):
    raise NotImplementedError("trimesh is not available in the stand-in")
