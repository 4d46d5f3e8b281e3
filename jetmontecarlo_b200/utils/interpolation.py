"""Host-side O(bins) helpers around the hot path (grids and interpolants of
finished integrals).  Counterparts of jetmontecarlo/utils/interpolation.py:
``lin_log_mixed_list`` (:18-41), ``get_1d_interpolation`` (:186-245),
``get_2d_interpolation`` (:248-273).  Nothing here touches per-sample data.
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


def is_monotonic_arr(arr, check_only=None):
    arr = np.asarray(arr)
    inc, dec = (arr[1:] >= arr[:-1]).all(), (arr[1:] <= arr[:-1]).all()
    if check_only == 'increasing':
        return bool(inc)
    if check_only == 'decreasing':
        return bool(dec)
    return bool(inc or dec)


def get_1d_interpolation(xs, fs, monotonic=True, bounds=None, bound_values=None, verbose=0):
    """Piecewise-linear interpolant of (xs, fs), clamped to ``bound_values``
    outside ``bounds`` (default: the data range, continuous at the ends)."""
    xs, fs = np.asarray(xs, dtype=float), np.asarray(fs, dtype=float)
    if monotonic:
        assert is_monotonic_arr(fs), "User asked for a monotone integral,"\
            " but the integral's interpolating function is not monotone!"

        def core(x):
            return np.interp(x, xs, fs)
    else:
        core = interpolate.interp1d(x=xs, y=fs, fill_value="extrapolate")
    if bounds is None:
        bounds = (xs[0], xs[-1])
    lo_val, hi_val = (None, None) if bound_values is None else bound_values
    lo_val = core(xs[0]) if lo_val is None else lo_val
    hi_val = core(xs[-1]) if hi_val is None else hi_val

    def bounded(x):
        x = np.asarray(x, dtype=float)
        inside = (bounds[0] <= x) & (x <= bounds[1])
        return np.where(x < bounds[0], lo_val, np.where(inside, core(np.clip(x, xs.min(), xs.max())
                                                                    if monotonic else x), hi_val))
    return bounded


def get_2d_interpolation(xs, ys, zs, interpolation_method="RectangularGrid", **kwargs):
    """ref: utils/interpolation.py:248-273 (grid and nearest-neighbour forms)."""
    if interpolation_method == "RectangularGrid":
        opts = {'method': 'linear', 'bounds_error': False, 'fill_value': None}
        opts.update(kwargs)
        return interpolate.RegularGridInterpolator((xs, ys), zs, **opts)
    if interpolation_method in ("Nearest", "nearest"):
        return interpolate.NearestNDInterpolator(list(zip(xs, ys)), zs)
    raise ValueError(f"Unknown interpolation method {interpolation_method}")
