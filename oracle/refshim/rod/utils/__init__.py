from . import resolve_uris  # noqa: F401
