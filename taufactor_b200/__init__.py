"""taufactor_b200 -- B200-native (sm_100a) implementation of TauFactor's steady-state diffusion
solve.  Drop-in for ``taufactor.Solver / PeriodicSolver / AnisotropicSolver / MultiPhaseSolver /
PeriodicMultiPhaseSolver`` (reference: tldr-group/taufactor v1.2.1, taufactor/__init__.py:3-11);
plus ``ElectrodeSolver / PeriodicElectrodeSolver`` (taufactor/electrode.py:13-157) on the same kernels and
the benchmark harness (``taufactor_b200.benchmark``, ``taufactor_b200.utils`` structure generators) and
``imread`` (TIFF volumes, what the reference's README / notebooks use ``tifffile.imread`` for);
everything else of the reference package (metrics, the complex-valued ImpedanceSolver, plotting) is out of
scope -- keep importing it from ``taufactor``."""
from .solvers import (AnisotropicSolver, MultiPhaseSolver, PeriodicMultiPhaseSolver, PeriodicSolver, Solver,
                      SORSolver, ThroughTransportSolver)
from .electrode import ElectrodeSolver, PeriodicElectrodeSolver
from .io import imread

__all__ = ["Solver", "PeriodicSolver", "AnisotropicSolver", "MultiPhaseSolver", "PeriodicMultiPhaseSolver",
           "ElectrodeSolver", "PeriodicElectrodeSolver", "SORSolver", "ThroughTransportSolver", "imread"]
__version__ = "0.1.0"
