"""No-op stand-in for `pydot` (drawing only; pyscqed/parameters.py:7-8). TEST INFRASTRUCTURE."""


class _Graph:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        def _noop(*a, **k):
            return b""
        return _noop


Dot = Node = Edge = Cluster = Subgraph = _Graph


def graph_from_dot_data(*a, **k):
    return [_Graph()]
