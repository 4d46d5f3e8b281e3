"""jetmontecarlo_b200 -- B200 (sm_100a) implementation of JetMonteCarlo's Monte
Carlo integration hot path behind the reference's own Python API.

    from jetmontecarlo_b200.montecarlo.integrator import *
    from jetmontecarlo_b200.numerics.radiators.samplers import *
    from jetmontecarlo_b200.numerics.observables import *
    from jetmontecarlo_b200.numerics.weights import *

mirror ``jetmontecarlo.*`` module for module.  All arithmetic on the path runs
in hand-written CUDA kernels (libjmc_b200.so, C ABI in include/jmc_b200.h);
there is no CPU fallback -- importing the package without the built library
raises ImportError and calling into it without a CUDA device raises
RuntimeError.
"""
import importlib
import importlib.util
import pkgutil
import sys

from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__version__ = "0.2.0"


def install_as(name="jetmontecarlo"):
    """Make ``import <name>.<module>`` resolve to this package's mirror of that module, so that a script written
    against the reference (``from jetmontecarlo.montecarlo.integrator import *`` ...) runs on the GPU path unchanged.
    Modules of the reference that are not on the hot path (plotting, ...) are not provided: importing them fails as
    it would without the reference installed.  The reference's second top-level package, ``file_management`` (the
    catalog of stored samples / integrals), is registered too unless a package of that name is importable.
    Returns the list of aliased module names."""
    me = sys.modules[__name__]
    aliased = []
    sys.modules[name] = me
    for info in pkgutil.walk_packages(me.__path__, prefix=__name__ + "."):
        short = info.name[len(__name__) + 1:]
        if short.startswith("_") or short.startswith("csrc") or short in ("fused", "distributed"):
            continue
        sys.modules[name + "." + short] = importlib.import_module(info.name)
        aliased.append(name + "." + short)
    if name == "jetmontecarlo" and "file_management" not in sys.modules and importlib.util.find_spec("file_management") is None:
        for short in ("file_management", "file_management.catalog_utils", "file_management.data_management",
                      "file_management.load_data"):
            sys.modules[short] = importlib.import_module(__name__ + "." + short)
            aliased.append(short)
    return aliased
