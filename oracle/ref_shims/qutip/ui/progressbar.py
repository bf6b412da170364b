"""Shim (test infrastructure only) for
/root/reference/src/krotov/parallelization.py:105."""


class BaseProgressBar:
    def __init__(self, iterations=0, chunk_size=10):
        pass

    def start(self, iterations, chunk_size=10):
        pass

    def update(self, n):
        pass

    def finished(self):
        pass


class TextProgressBar(BaseProgressBar):
    pass
