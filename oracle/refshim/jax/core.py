class Tracer:
    """Never instantiated: the NumPy stand-in does not trace."""
