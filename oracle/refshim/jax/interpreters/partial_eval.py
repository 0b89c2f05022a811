class DynamicJaxprTracer:
    """Never instantiated: the NumPy stand-in does not trace."""
