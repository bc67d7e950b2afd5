"""The reference code AS WRITTEN, behind the same Python wrappers as the C oracle.

TEST INFRASTRUCTURE ONLY (same rule as oracle/oracle.py).  `oracle/_ref/liblna_ref.so` is the reference's own
sources (LNA_functional.cpp, basic_util.cpp, SIR_LNA.cpp, SIR_log_LNA.cpp, SIRS_period.cpp, SIR_phylodyn.cpp)
compiled verbatim from /root/reference/src against the stand-in headers of oracle/shim/ (recipe: `make -C oracle
ref`).  Every `ref_X` it exports has the C signature of `orc_X`, so this module simply re-executes oracle.py
with the library swapped: `reference.volz_loglik_nh2(...)`, `reference.ESlice_par_General(...)` etc. take
the same arguments as their oracle twins and run the reference's code.

`/root/reference` exists only in the build container.  There `build()` compiles the library; on the GPU box the
prebuilt .so travels with the repo snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored) and `available()`
tells whether it is there.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liblna_ref.so")
_SO_DROPIN = os.path.join(_HERE, "_ref", "liblna_dropin.so")
REF_SRC = os.environ.get("LNA_REFERENCE_SRC", "/root/reference/src")


def build(force=False):
    """Compile oracle/_ref/liblna_ref.so when the reference tree is present; returns the path or None."""
    if not os.path.isdir(REF_SRC):
        return _SO if os.path.exists(_SO) else None
    cmd = ["make", "-C", _HERE, "-s", "ref", f"REF_SRC={REF_SRC}"]
    if force:
        cmd.insert(1, "-B")
    subprocess.check_call(cmd)
    return _SO


def available():
    return os.path.exists(_SO) or (os.path.isdir(REF_SRC) and build() is not None)


def build_dropin(force=False):
    """Compile oracle/_ref/liblna_dropin.so: the SAME harness (ref_harness.cpp) over integration/rcpp_shim.cpp -- the
    reference-side binding that forwards every hot-path export to the C ABI -- instead of the reference's own bodies.
    Needs the reference's header (LNA_functional.h) and the built CUDA library to link against."""
    if not os.path.isdir(REF_SRC):
        return _SO_DROPIN if os.path.exists(_SO_DROPIN) else None
    cmd = ["make", "-C", _HERE, "-s", "dropin", f"REF_SRC={REF_SRC}"]
    if force:
        cmd.insert(1, "-B")
    subprocess.check_call(cmd)
    return _SO_DROPIN


def dropin_available():
    return os.path.exists(_SO_DROPIN) or (os.path.isdir(REF_SRC) and build_dropin() is not None)


class _RefLib:
    """Maps orc_X -> ref_X of liblna_ref.so and raises on errors the reference would have thrown as R errors."""

    def __init__(self, so=_SO, builder=build):
        if not os.path.exists(so) and builder() is None:
            raise RuntimeError(f"{so} is missing and /root/reference is not here to build it")
        self._dll = C.CDLL(so)
        self._dll.ref_last_error.restype = C.c_char_p
        for name in ("volz_loglik_nh2", "volz_loglik_nh3", "coal_loglik3", "coal_loglik", "log_like_traj_general2",
                     "ode_rk45_stop", "volz_loglik_nh"):
            getattr(self._dll, "ref_" + name).restype = C.c_double
        self._cache = {}

    def __getattr__(self, name):
        if not name.startswith("orc_"):
            raise AttributeError(name)
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._dll, "ref_" + name[4:])

            def fn(*a, _raw=raw):
                self._dll.ref_clear_error()
                out = _raw(*a)
                err = self._dll.ref_last_error()
                if err and not (isinstance(out, int) and out == 1):  # status 1 = the reference's chol exception
                    raise RuntimeError("reference threw: " + err.decode())
                return out

            self._cache[name] = fn
        return fn


_lib = None


def _get_lib():
    global _lib
    if _lib is None:
        _lib = _RefLib()
    return _lib


def _load_wrappers(name="oracle._reference_wrappers", getter=None, builder=build):
    spec = importlib.util.spec_from_file_location(name, os.path.join(_HERE, "oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.lib = getter or _get_lib
    mod.build = builder
    return mod


_dropin = None


def dropin():
    """The wrappers of oracle.py over liblna_dropin.so: `dropin().ESlice_par_General(...)` runs the reference-side
    binding integration/rcpp_shim.cpp -> C ABI -> GPU, with the arguments of its reference and oracle twins."""
    global _dropin
    if _dropin is None:
        holder = {}

        def getter():
            if "lib" not in holder:
                holder["lib"] = _RefLib(_SO_DROPIN, build_dropin)
            return holder["lib"]

        _dropin = _load_wrappers("oracle._dropin_wrappers", getter, build_dropin)
    return _dropin


_w = _load_wrappers()
_this = sys.modules[__name__]
for _name in dir(_w):
    if not _name.startswith("_") and not hasattr(_this, _name):
        setattr(_this, _name, getattr(_w, _name))
Cfg, Init = _w.Cfg, _w.Init


def driver():
    """oracle/driver.py (the R-level iteration bodies) composed from the REFERENCE's functions instead of the oracle's."""
    spec = importlib.util.spec_from_file_location("oracle._reference_driver", os.path.join(_HERE, "driver.py"))
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "oracle"
    spec.loader.exec_module(mod)
    mod.orc = _this
    return mod
