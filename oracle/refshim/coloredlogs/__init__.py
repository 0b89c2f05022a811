import logging


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, **_):
        super().__init__(fmt=fmt)


def install(*_a, **_k):
    pass
