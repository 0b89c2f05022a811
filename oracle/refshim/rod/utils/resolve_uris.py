"""`rod.utils.resolve_uris.resolve_local_uri`: only mesh collisions use it (disabled by default in the reference,
JAXSIM_COLLISION_MESH_ENABLED=0)."""
import pathlib


def resolve_local_uri(uri: str) -> pathlib.Path:
    return pathlib.Path(uri.replace("file://", ""))
