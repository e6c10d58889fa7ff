"""ctypes binding of libhyperelastic_b200.so (include/hyperelastic_b200.h).

The shared library is the product; this module only loads it.  If the library has not been
built, or no CUDA device is visible, calls fail loudly -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("HYPERELASTIC_B200_LIB") or os.path.join(_HERE, "libhyperelastic_b200.so")  # env override: tuning builds
CSRC = os.path.join(_HERE, "csrc")

HS_OK, HS_ERR_ARG, HS_ERR_CUDA, HS_ERR_DOMAIN, HS_ERR_EXCHANGE = 0, 1, 2, 3, 4
SP13, MPH30 = 0, 1
LXF, HLL = 0, 1
NVAR = {SP13: 13, MPH30: 30}
NPHASE = {SP13: 1, MPH30: 2}
NAUX = {SP13: 5, MPH30: 2}   # HS_NAUX(model): cached per-cell rows (SP: c_max, 1/rho, stress row 1; MPh: wave bounds); re-read from the library in lib()
HS_SCAL_SLOTS = 8


class HyperelasticError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hyperelastic_b200 error {code}: {msg}")
        self.code = code


class DomainError(HyperelasticError):
    """Raised where the Julia reference would throw DomainError (sqrt/log of a negative number)."""


class Barton2009(C.Structure):
    """EquationsOfState.jl:71-116 (same field order, same defaults)."""
    _fields_ = [(n, C.c_double) for n in ("rho0", "c0", "cv", "t0", "b0", "alpha", "beta", "gamma", "b0sq", "k0")]

    def __init__(self, _rho0=8.93, _c0=4.6, _cv=3.9e-4, _t0=300, _b0=2.1, _alpha=1, _beta=3, _gamma=2):
        super().__init__(float(_rho0), float(_c0), float(_cv), float(_t0), float(_b0), float(_alpha), float(_beta),
                         float(_gamma), float(_b0) ** 2, float(_c0) ** 2 - (4 / 3) * float(_b0) ** 2)

    def as_tuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class Hank2016(C.Structure):
    """EquationsOfState.jl:305-319 (same field order, same defaults)."""
    _fields_ = [(n, C.c_double) for n in ("rho0", "mu", "gamma", "pres_inf", "a")]

    def __init__(self, rho0=2.7, mu=26e9, gamma=3.4, pres_inf=21.5e9, a=0.5):
        super().__init__(float(rho0), float(mu), float(gamma), float(pres_inf), float(a))


class HsdProblem(C.Structure):
    _fields_ = [("model", C.c_int), ("nphase", C.c_int), ("gen", C.c_int), ("reserved", C.c_int),
                ("ncells", C.c_int64), ("nprob", C.c_int64), ("stride", C.c_int64),
                ("eos_dev", (C.c_double * 20) * 2)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_vp = C.c_void_p
_i64 = C.c_int64
_eosp = C.POINTER(Barton2009)

# name -> (restype, argtypes); every symbol include/hyperelastic_b200.h declares
SIGNATURES = {
    "hs_version": (C.c_char_p, []),
    "hs_last_error": (C.c_char_p, []),
    "hs_device_count": (C.c_int, []),
    "hs_create": (C.c_int, [C.POINTER(_vp), C.c_int, _eosp, C.c_int, _i64, _i64, C.c_int]),
    "hs_create_multi": (C.c_int, [C.POINTER(_vp), C.c_int, _eosp, C.c_int, _i64, _i64, C.POINTER(C.c_int), C.c_int]),
    "hs_destroy": (C.c_int, [_vp]),
    "hs_upload": (C.c_int, [_vp, _vp]),
    "hs_download": (C.c_int, [_vp, _vp]),
    "hs_set_time": (C.c_int, [_vp, C.c_double, _i64]),
    "hs_wave_speeds": (C.c_int, [_vp, _vp, _vp]),
    "hs_step": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _vp]),
    "hs_advance": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, _i64, _vp, _vp, _vp]),
    "hs_step_host": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp]),
    "hs_host_register": (C.c_int, [_vp, C.c_size_t]),
    "hs_host_unregister": (C.c_int, [_vp]),
    "hs_step_host_stats": (C.c_int, [_vp, _ip, _ip]),
    "hs2d_create": (C.c_int, [C.POINTER(_vp), C.c_int, _eosp, C.c_int, _i64, _i64, C.c_int]),
    "hs2d_destroy": (C.c_int, [_vp]),
    "hs2d_upload": (C.c_int, [_vp, _vp]),
    "hs2d_download": (C.c_int, [_vp, _vp]),
    "hs2d_step": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, _vp]),
    "hs2d_advance": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _i64, _vp, _vp]),
    "hs_cons2prim": (C.c_int, [C.c_int, _eosp, C.c_int, _vp, _vp, _i64, C.c_int]),
    "hs_prim2cons": (C.c_int, [C.c_int, _eosp, C.c_int, _vp, _vp, _i64, C.c_int]),
    "hs_flux": (C.c_int, [C.c_int, _eosp, C.c_int, _vp, _vp, _i64, C.c_int]),
    "hs_noncons_flux": (C.c_int, [_eosp, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_get_eigvals": (C.c_int, [C.c_int, _eosp, C.c_int, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_hll": (C.c_int, [C.c_int, _eosp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_lxf": (C.c_int, [C.c_int, _eosp, C.c_int, _vp, _vp, C.c_double, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_hank2016_energy": (C.c_int, [C.POINTER(Hank2016), _vp, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_hank2016_pressure": (C.c_int, [C.POINTER(Hank2016), _vp, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_hank2016_stress": (C.c_int, [C.POINTER(Hank2016), _vp, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_selftest_math": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.c_int]),
    "hs_selftest_eig": (C.c_int, [_vp, _vp, _i64, C.c_int]),
    "hsd_problem_init": (C.c_int, [C.POINTER(HsdProblem), C.c_int, _eosp, C.c_int, _i64, _i64]),
    "hsd_aos_to_soa": (C.c_int, [C.POINTER(HsdProblem), _vp, _vp, _vp]),
    "hsd_soa_to_aos": (C.c_int, [C.POINTER(HsdProblem), _vp, _vp, _vp]),
    "hsd_wave_bounds": (C.c_int, [C.POINTER(HsdProblem), _vp, _vp, _vp, C.c_int, _vp]),
    "hsd_wave_bounds_acc": (C.c_int, [C.POINTER(HsdProblem), _vp, _vp, _vp, C.c_int, _vp]),
    "hsd_step": (C.c_int, [C.POINTER(HsdProblem), C.c_int, C.c_double, C.c_double, C.c_double, _i64,
                           _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, C.c_int, _vp]),
    "hsd_halo": (C.c_int, [C.POINTER(HsdProblem), _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "hsd_mailbox_doubles": (C.c_int, []),
    "hsd_naux": (C.c_int, [C.c_int]),
    "hsd_exchange_p2p": (C.c_int, [C.POINTER(HsdProblem), _vp, _vp, _vp, C.POINTER(_vp), C.c_int, C.c_int, C.c_uint64, _vp, _vp]),
    "hsd_scal_lambda_next": (_vp, [_vp, _i64, _i64]),
    "hsd_scal_lambda_cur": (_vp, [_vp, _i64, _i64]),
    "hsd_scal_time": (_vp, [_vp, _i64, _i64]),
    "hsd_scal_steps": (_vp, [_vp, _i64]),
    "hsd_scal_status": (_vp, [_vp, _i64]),
    "hs_kernel_launch_count": (_i64, []),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libhyperelastic_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or `make -C hyperelasticsolver_b200/csrc`).  There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        NAUX[SP13], NAUX[MPH30] = L.hsd_naux(SP13), L.hsd_naux(MPH30)   # the layout the library was built with
        _lib = L
    return _lib


def check(rc: int, allow_domain: bool = False) -> int:
    if rc == HS_OK:
        return rc
    msg = lib().hs_last_error().decode()
    if rc == HS_ERR_DOMAIN:
        if allow_domain:
            return rc
        raise DomainError(rc, msg)
    raise HyperelasticError(rc, msg)


def eos_array(eos, model):
    """tuple/list of Barton2009 -> contiguous ctypes array of the right length."""
    if isinstance(eos, Barton2009):
        eos = (eos,)
    eos = tuple(eos)
    n = NPHASE[model]
    if len(eos) != n:
        raise ValueError(f"model needs {n} equation(s) of state, got {len(eos)}")
    arr = (Barton2009 * n)()
    for i, e in enumerate(eos):
        C.memmove(C.byref(arr[i]), C.byref(e), C.sizeof(Barton2009))
    return arr
