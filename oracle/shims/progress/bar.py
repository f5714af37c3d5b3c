"""No-op progress bar (pyscqed/numerical_system.py:943,988). TEST INFRASTRUCTURE."""


class Bar:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def next(self, *a, **k):
        pass

    def finish(self):
        pass
