"""hyperelasticsolver_b200 -- B200-native finite-volume hot path of
BlackSiberian/HyperelasticSolver behind a C ABI (include/hyperelastic_b200.h).

The Python layer mirrors the reference's Julia module API (same function names, argument
meaning and error behaviour) and is a thin ctypes binding: all arithmetic runs in
libhyperelastic_b200.so on the GPU.  There is no CPU fallback.
"""
from ._lib import (Barton2009, Hank2016, DomainError, HyperelasticError, HLL, LXF, MPH30, SP13, build, lib)
from .hyperelasticity_mph import (cons2prim_mph, flux_mph, get_eigvals, initial_states, noncons_flux, prim2cons_mph)
from .num_fluxes import hll, lxf
from .solver import Solver, Solver2D, initial_condition, update_cell, register_host, unregister_host
from . import hyperelasticity
from . import equations_of_state

__all__ = ["Barton2009", "Hank2016", "equations_of_state", "DomainError", "HyperelasticError", "HLL", "LXF", "MPH30", "SP13", "build", "lib",
           "cons2prim_mph", "flux_mph", "get_eigvals", "initial_states", "noncons_flux", "prim2cons_mph",
           "hll", "lxf", "Solver", "initial_condition", "update_cell", "hyperelasticity"]
