"""Import shim for the UNMODIFIED reference (test infrastructure only).

The reference lives read-only at /root/reference and exists only in the build
container (never on the GPU box).  It is pure Python but two of its imports are
broken against the SciPy in this image:

* ``from scipy.misc import derivative``  (jetmontecarlo/utils/interpolation.py:5,
  jetmontecarlo/utils/montecarlo_utils.py:5) -- removed from SciPy;
* ``from pynverse import inversefunc``   (jetmontecarlo/utils/montecarlo_utils.py:6)
  -- not installed.

Neither symbol is touched by the Monte Carlo hot path, so we register inert
stand-ins *before* importing and leave the reference sources untouched.  Only
``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py`` use this
module; nothing in the product package may import it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("JMC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "jetmontecarlo"))


def _central_difference(func, x0, dx=1.0, n=1, args=(), order=3):
    # first derivative only; the hot path never calls it
    return (func(x0 + dx, *args) - func(x0 - dx, *args)) / (2.0 * dx)


def install():
    """Make ``import jetmontecarlo`` resolve to the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    import scipy
    try:
        import scipy.misc as misc
    except Exception:  # scipy.misc itself is gone in new SciPy
        misc = types.ModuleType("scipy.misc")
        sys.modules["scipy.misc"] = misc
        scipy.misc = misc
    if not hasattr(misc, "derivative"):
        misc.derivative = _central_difference
    if "pynverse" not in sys.modules:
        stub = types.ModuleType("pynverse")

        def inversefunc(*a, **k):
            raise NotImplementedError("pynverse is not installed in this image")
        stub.inversefunc = inversefunc
        sys.modules["pynverse"] = stub
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            # montecarlo/ewoc.py imports pyplot for its plot method only: an empty stand-in lets the module import
            mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
