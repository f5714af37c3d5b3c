"""No-op stand-in for the `progress` package (pyscqed/numerical_system.py:7). TEST INFRASTRUCTURE."""
