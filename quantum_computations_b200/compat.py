"""Expose this package under the reference's import path.

The reference's other packages import ``simulators.dv_simulator.gates`` etc. and
dispatch on the classes found there (GKP/transpiler.py:1-4, :41-63).  Calling
``install()`` before those imports registers this package's modules under the
same dotted names, so unmodified reference scripts pick up the CUDA-backed
classes:

    import quantum_computations_b200.compat as compat
    compat.install()
    from simulators.dv_simulator.simulator import Simulator   # -> this package
"""
from __future__ import annotations

import sys
import types

from . import gates, numpy_quantum, simulator, states

_NAMES = {
    "gates": gates,
    "numpy_quantum": numpy_quantum,
    "simulator": simulator,
    "states": states,
}


def install(force: bool = False) -> None:
    """Register ``simulators`` and ``simulators.dv_simulator`` aliases in
    ``sys.modules``.  If a real ``simulators`` package is importable it is kept
    (so ``simulators.gkp_simulator`` still resolves) and only its
    ``dv_simulator`` sub-package is replaced."""
    parent = sys.modules.get("simulators")
    if parent is None:
        try:
            import simulators as parent  # type: ignore  # the reference checkout, if on sys.path
        except ImportError:
            parent = types.ModuleType("simulators")
            parent.__path__ = []          # mark as package
            sys.modules["simulators"] = parent
    existing = sys.modules.get("simulators.dv_simulator")
    if existing is not None and not force and getattr(existing, "__qsim_b200__", False):
        return
    pkg = types.ModuleType("simulators.dv_simulator")
    pkg.__path__ = []
    pkg.__qsim_b200__ = True
    for name, module in _NAMES.items():
        setattr(pkg, name, module)
        sys.modules[f"simulators.dv_simulator.{name}"] = module
    sys.modules["simulators.dv_simulator"] = pkg
    parent.dv_simulator = pkg
