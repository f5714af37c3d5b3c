"""B200-native sweep engine for PySCQED circuit Hamiltonians (hot path only).

``NumericalSystem`` is the drop-in for ``pyscqed.NumericalSystem``'s sweep / diagonalisation API;
``Engine`` is the thin ctypes layer over ``libscqed.so`` (include/scqed.h).
"""
from .plan import SweepPlan, NodeSpec  # noqa: F401
from .coefprog import CoefProgram  # noqa: F401

__all__ = ["SweepPlan", "NodeSpec", "CoefProgram", "NumericalSystem", "Engine", "lower_system"]
__version__ = "0.1.0"


def __getattr__(name):
    if name == "NumericalSystem":
        from .numerical_system import NumericalSystem
        return NumericalSystem
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "lower_system":
        from .lowering import lower_system
        return lower_system
    raise AttributeError(name)
