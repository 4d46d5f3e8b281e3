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
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__version__ = "0.1.0"
