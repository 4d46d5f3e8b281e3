"""Host-side O(bins) helpers around the hot path: grids, monotonicity checks and interpolants of finished
integrals -- counterparts of jetmontecarlo/utils/interpolation.py with the same results on the same inputs
(tests/test_interpolation_cpu.py checks them against vectors produced by the reference).  Nothing here touches
per-sample data.

Two properties of the reference are worth knowing when reading results:

* ``get_1d_interpolation(monotonic=True)`` without ``bounds`` restricts the interpolant to the monotonic domain
  found inside the FIRST bin, ``[xs[0], xs[1]]`` (interpolation.py:215), and returns the boundary values outside
  it; at a bound itself the boundary value and the interpolant add up (the three masks at :238-241 overlap).  Both
  are reproduced; pass ``bounds=(xs[0], xs[-1])`` for an interpolant over the whole table.
* its default ``bound_values=[None, None]`` is one mutable list that the reference fills in on first use, so every
  later default call in the same process inherits the first table's boundary values (:188,229-232).  That
  process-history dependence is NOT reproduced: each call here starts from ``[None, None]``.
"""
import numpy as np
from scipy import interpolate


def lin_log_mixed_list(lower_bound, upper_bound, num_vals):
    """Half log-spaced, half lin-spaced grid, sorted, end points of the log half
    dropped.  ref: utils/interpolation.py:18-41"""
    half = int(num_vals / 2)
    log_part = np.logspace(np.log10(lower_bound), np.log10(upper_bound), half + 2)
    lin_part = np.linspace(lower_bound, upper_bound, half)
    merged = np.sort(np.append(log_part, lin_part)[1:-1])
    return merged[~np.isnan(merged)]


# ---- monotonicity -----------------------------------------------------------------------------------------------
def where_monotonic_arr(arr, check_only):
    """Indices i with arr[i] -> arr[i+1] going the requested way.  ref: utils/interpolation.py:43-60"""
    assert check_only in ['increasing', 'decreasing']
    arr = np.asarray(arr)
    steps_up = arr[1:] >= arr[:-1]
    steps_down = arr[:-1] >= arr[1:]
    return np.where(steps_up if check_only == 'increasing' else steps_down)[0]


def is_monotonic_arr(arr, check_only=None):
    """ref: utils/interpolation.py:63-83"""
    assert check_only in ['increasing', 'decreasing', None]
    arr = np.asarray(arr)
    up, down = (arr[1:] >= arr[:-1]).all(), (arr[:-1] >= arr[1:]).all()
    if check_only == 'increasing':
        return up
    if check_only == 'decreasing':
        return down
    return up or down


def is_monotonic_func(func, domain, check_only=None):
    """Monotonicity of ``func`` on 500 mixed lin/log points of ``domain``.  ref: utils/interpolation.py:86-101"""
    return is_monotonic_arr(func(lin_log_mixed_list(domain[0], domain[1], 500)), check_only)


def _slope(func, x, dx=1e-8):
    # scipy.misc.derivative(func, x, dx) of the reference: three-point central difference
    return (func(x + dx) - func(x - dx)) / (2.0 * dx)


def monotonic_domain(func, domain, include_point=None, check_only=None):
    """Largest interval around ``include_point`` (default: the middle of ``domain``) on which the central-difference
    slope of ``func`` keeps its sign, located on a 10000-point grid and refined recursively inside the grid cell
    that holds the point; None if the slope at the point goes against ``check_only``.
    ref: utils/interpolation.py:104-180"""
    if include_point is None:
        include_point = (domain[0] + domain[1]) / 2
    assert domain[0] < include_point < domain[1]
    at_point = _slope(func, include_point)
    if (check_only in ['increasing'] and at_point < 0) or (check_only in ['decreasing'] and at_point > 0):
        return None
    if domain[0] > 0:
        grid = lin_log_mixed_list(domain[0] + 1e-20, domain[1], 10000)
    else:
        grid = np.linspace(domain[0], domain[1], 10000)
    slopes = _slope(func, grid)
    turns = [i for i in range(len(slopes) - 1) if slopes[i] * slopes[i + 1] < 0]
    if len(turns) == 0:
        return domain
    lower, upper = domain[0], domain[1]
    for i in turns:
        if include_point >= grid[i + 1]:
            lower = grid[i + 1]
        if include_point <= grid[i]:
            upper = grid[i]
            return (lower, upper)
        if grid[i] <= include_point <= grid[i + 1]:
            inner = monotonic_domain(func, (grid[i], grid[i + 1]), include_point, check_only)
            assert inner is not None
            assert inner != domain
            inner_lower, inner_upper = inner
            if include_point > inner_lower:
                lower = inner_lower
            if include_point < inner_upper:
                upper = inner_upper
                return (lower, upper)
    return (lower, upper)


# ---- interpolants -----------------------------------------------------------------------------------------------
def get_1d_interpolation(xs, fs, monotonic=True, bounds=None, bound_values=(None, None), verbose=0):
    """Piecewise-linear interpolant of the table (xs, fs): ``np.interp`` (flat beyond the table) when
    ``monotonic``, after asserting that the table is monotone; else ``interp1d`` with linear extrapolation.
    Outside ``bounds`` the function takes ``bound_values`` (None: the interpolant's value at xs[0] / xs[-1]); see the
    module docstring for what the reference does when ``monotonic`` and no bounds are given, and at the bounds.
    ref: utils/interpolation.py:186-245"""
    if monotonic:
        def core(x):
            return np.interp(x, xs, fs)

        at_nodes = core(xs)
        is_monotone = ((at_nodes[1:] <= at_nodes[:-1]).all() or (at_nodes[1:] >= at_nodes[:-1]).all())
        assert is_monotone, "User asked for a monotone integral,"\
            " but the integral's interpolating function is not monotone!"
    else:
        core = interpolate.interp1d(x=xs, y=fs, fill_value="extrapolate")

    if monotonic and bounds is None:
        if verbose >= 1:
            print("Setting the bounds of the interpolating function to its monotonic domain.")
        bounds = monotonic_domain(core, [xs[0], xs[1]])
        if verbose >= 2:
            print(f"monotonic domain: {bounds}")

    if bounds is None:
        assert bound_values is None, "Cannot assign boundary values if"\
            " no bounds for the interpolation function are given."
        # (the reference has no function to return on this path and fails with UnboundLocalError at :245)
        raise ValueError("get_1d_interpolation: without bounds there is nothing to return (pass bounds, or "
                         "monotonic=True)")

    values = [None, None] if bound_values is None else list(bound_values)
    assert len(values) == 2, "bound_values must be a tuple of length 2"
    if values[0] is None:
        values[0] = core(xs[0])
    if values[1] is None:
        values[1] = core(xs[-1])
    lower, upper = bounds[0], bounds[1]

    def bounded(x):
        return values[0] * (x <= lower) + (lower <= x) * (x <= upper) * core(x) + values[1] * (x >= upper)
    return bounded


def get_2d_interpolation(xs, ys, zs, interpolation_method="RectangularGrid", **kwargs):
    """Interpolant of z over (x, y): "RectangularGrid" (linear, extrapolating: the default), or "Nearest" for
    scattered points.  The reference's "Linear"/"Cubic" methods are scipy's ``interp2d``, which SciPy 1.14 removed.
    ref: utils/interpolation.py:248-273"""
    if interpolation_method == "RectangularGrid":
        opts = {'method': 'linear', 'bounds_error': False, 'fill_value': None}
        opts.update(kwargs)
        return interpolate.RegularGridInterpolator((xs, ys), zs, **opts)
    if interpolation_method in ("Linear", "linear", "Cubic", "cubic"):
        raise NotImplementedError("scipy.interpolate.interp2d was removed in SciPy 1.14; use 'RectangularGrid'")
    if interpolation_method in ("Nearest", "nearest"):
        return interpolate.NearestNDInterpolator(list(zip(xs, ys)), zs)
    raise ValueError(f"Unknown interpolation method {interpolation_method}")
