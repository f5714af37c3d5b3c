"""No-op stand-in for `graphviz` (drawing only; pyscqed/circuit_graph.py:359-382). TEST INFRASTRUCTURE."""


class Source:
    def __init__(self, src=None, *a, **k):
        self.source = src
