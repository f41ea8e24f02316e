"""taufactor_b200 -- B200-native (sm_100a) implementation of TauFactor's steady-state diffusion
solve.  Drop-in for ``taufactor.Solver / PeriodicSolver / AnisotropicSolver / MultiPhaseSolver /
PeriodicMultiPhaseSolver`` (reference: tldr-group/taufactor v1.2.1, taufactor/__init__.py:3-11);
everything else of the reference package (metrics, electrode / impedance solvers, plotting) is out
of scope -- keep importing it from ``taufactor``."""
from .solvers import (AnisotropicSolver, MultiPhaseSolver, PeriodicMultiPhaseSolver, PeriodicSolver, Solver,
                      SORSolver, ThroughTransportSolver)

__all__ = ["Solver", "PeriodicSolver", "AnisotropicSolver", "MultiPhaseSolver", "PeriodicMultiPhaseSolver",
           "SORSolver", "ThroughTransportSolver"]
__version__ = "0.1.0"
