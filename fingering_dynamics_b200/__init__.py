"""fingering_dynamics_b200 -- B200-native two-phase D2Q9 lattice-Boltzmann step (sm_100a CUDA behind a C ABI).

Only the reference's hot path is here: the time step of lattice_boltzmann/fingering*.py and validation.py
(collide + interaction force + stream + half-way bounce-back + Zou-He / periodic faces + moments), the
geometry bitfields that feed it, and drop-in mirrors of the reference's module-level entry points
(fingering_dynamics_b200/lattice_boltzmann/).
"""
from .engine import Engine, pinned_empty, pinned_free  # noqa: F401
from . import geometry  # noqa: F401

__all__ = ["Engine", "geometry", "pinned_empty", "pinned_free"]
