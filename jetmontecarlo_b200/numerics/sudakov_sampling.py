"""Inverse-transform "Sudakov samples", one sample per event from a cdf that depends on the event -- the stage between
the numerical radiators and the final pdf in the reference's production workflow
(examples/sudakov_comparisons/sudakov_utils.py:100-238: ``generate_critical_samples``,
``generate_precritical_samples``, ``generate_subsequent_samples``).

The reference loops over events in Python; every iteration calls ``samples_from_cdf(cdf_i, 1, domain_i,
force_monotone=True)`` (utils/montecarlo_utils.py:93-375), i.e. evaluates the conditional cdf on ~1000 probe points
and walks a list of known cdf shapes.  Here all events go through that same list at once: the probe grids, the
conditional cdf table [events, 1001], the shape tests and the table inversion are tensor operations on the device
the conditioning values live on (PyTorch element-wise / sort / searchsorted kernels; no per-event host work).
Events whose domain has to be cut (negligible head, turning point) are re-tabulated together in further rounds,
exactly as the reference recurses.  ``tests/test_sudakov_sampling_cpu.py`` holds this to the per-event loop of the
oracle (which is pinned to the reference's own tabulated branches) on CPU tensors; the same code runs on CUDA
tensors in the product.
"""
import numpy as np
import torch

_THRESHOLD = 1e-20          # "negligible cdf" of the reference, utils/montecarlo_utils.py:237


# ---- probe grids (utils/interpolation.py:18-41 and utils/montecarlo_utils.py:147-154, one row per event) ---------
def _linspace_rows(start, stop, num):
    """np.linspace(start, stop, num) for every row: arange * step + start, last column set to stop."""
    k = torch.arange(num, dtype=start.dtype, device=start.device)[None, :]
    y = k * ((stop - start) / (num - 1)) + start
    y[:, -1] = stop[:, 0]
    return y


def _lin_log_mixed_rows(lower, upper, num_vals):
    half = int(num_vals / 2)
    log_part = torch.pow(10., _linspace_rows(torch.log10(lower), torch.log10(upper), half + 2))
    lin_part = _linspace_rows(lower, upper, half)
    merged = torch.cat([log_part, lin_part], dim=1)[:, 1:-1]
    return torch.sort(merged, dim=1).values


def probe_points(lower, upper):
    """[n, M] probe points of the domains [lower_i, upper_i] (tensors [n]); all rows must belong to one family:
    lower == 0 (M = 1001: the point 0 and the mixed lin/log grid from 1e-20), lower > 0 (M = 1000: mixed grid) or
    lower < 0 (M = 1000: linear grid)."""
    lower, upper = lower[:, None], upper[:, None]
    if bool((lower < 0).all()):
        return _linspace_rows(lower, upper, 1000)
    if bool((lower == 0).all()):
        grid = _lin_log_mixed_rows(lower + 1e-20, upper, 1000)
        return torch.cat([torch.zeros_like(lower), grid], dim=1)
    assert bool((lower > 0).all()), "probe_points: mixed domain families"
    return _lin_log_mixed_rows(lower, upper, 1000)


# ---- table inversion (scipy interp1d(cdf, x, fill_value='extrapolate') per row) ---------------------------------
def invert_tables(cdf_vals, x_vals, keep, r):
    """x(r_i) from row i of the table (cdf_vals, x_vals) restricted to ``keep``: entries sorted by cdf value (stable),
    linear interpolation between the neighbours of r_i, linear extrapolation beyond the table.
    ref: utils/montecarlo_utils.py:377-399"""
    inf = torch.full_like(cdf_vals, float('inf'))
    c = torch.where(keep, cdf_vals, inf)
    c, order = torch.sort(c, dim=1, stable=True)
    x = torch.gather(x_vals, 1, order)
    n_keep = keep.sum(dim=1)
    assert bool((n_keep >= 2).all()), "a cdf table needs at least two points"
    idx = torch.searchsorted(c, r[:, None].contiguous()).squeeze(1)
    idx = torch.minimum(torch.clamp(idx, min=1), n_keep - 1)
    lo, hi = (idx - 1)[:, None], idx[:, None]
    c_lo, c_hi = torch.gather(c, 1, lo).squeeze(1), torch.gather(c, 1, hi).squeeze(1)
    x_lo, x_hi = torch.gather(x, 1, lo).squeeze(1), torch.gather(x, 1, hi).squeeze(1)
    # the two-weight form of current SciPy's interp1d (older releases wrote slope * (r - c_lo) + x_lo: same to rounding)
    return ((r - c_lo) / (c_hi - c_lo)) * x_hi + ((c_hi - r) / (c_hi - c_lo)) * x_lo


# ---- the reference's list of cdf shapes, all events at once ----------------------------------------------------------
def conditional_samples_from_cdf(cdf, conds, domain_lo, domain_hi, uniforms, catch_turnpoint=False,
                                 force_monotone=False, max_rounds=64):
    """One sample per event i from the cdf ``x -> cdf(x, conds[i])`` on [domain_lo[i], domain_hi[i]].

    ``cdf(x, cond)`` takes tensors x [n, M] and cond [n, 1] and returns [n, M]; ``conds``, ``domain_lo``,
    ``domain_hi`` and ``uniforms`` are tensors [N] on one device.  Per event this is
    ``samples_from_cdf(cdf_i, 1, domain_i, catch_turnpoint, force_monotone)`` on its tabulated route with the
    uniform ``uniforms[i]``; returns (samples [N], weights [N] = 1).  ValueError if some event's cdf is not monotone
    and none of the options resolves it.  ref: utils/montecarlo_utils.py:142-318"""
    n_all = conds.numel()
    dev, f64 = conds.device, torch.float64
    samples = torch.full((n_all,), float('nan'), dtype=f64, device=dev)
    lo, hi = domain_lo.to(f64).clone(), domain_hi.to(f64).clone()
    pending = torch.arange(n_all, device=dev)
    for _ in range(max_rounds):
        if pending.numel() == 0:
            break
        still = []
        families = ((lo[pending] < 0), (lo[pending] == 0), (lo[pending] > 0))
        for family in families:
            rows = pending[family]
            if rows.numel() == 0:
                continue
            pnts = probe_points(lo[rows], hi[rows])
            raw = cdf(pnts, conds[rows][:, None])
            vals = torch.nan_to_num(raw)                      # np.nan_to_num: NaN -> 0, +-inf -> largest finite
            mono = vals[:, :-1] <= vals[:, 1:]
            ones = torch.ones_like(vals[:, :1], dtype=torch.bool)
            r = uniforms[rows]
            all_mono = mono.all(dim=1)
            # first place where the cdf decreases, and the (raw) cdf there
            first_bad = torch.argmax((~mono).to(torch.int32), dim=1)
            raw_bad = torch.gather(raw, 1, first_bad[:, None]).squeeze(1)
            always_one = (vals == 1.).all(dim=1)
            starts_at_one = (first_bad == 0) & (raw_bad == 1.)
            negligible = (vals < _THRESHOLD).all(dim=1)
            # first point with a non-negligible cdf; is the table monotone from there on?
            above = vals > _THRESHOLD
            first_above = torch.argmax(above.to(torch.int32), dim=1)
            tail_mono = torch.flip(torch.cumprod(torch.flip(torch.cat([mono, ones], dim=1).to(torch.int32), [1]), 1), [1])
            monotone_after = torch.gather(tail_mono, 1, first_above[:, None]).squeeze(1).bool() & above.any(dim=1)
            turning = (raw_bad == 1.) & bool(catch_turnpoint)

            use_table = all_mono
            zero_bin = ~all_mono & (always_one | starts_at_one)
            overflow = ~all_mono & ~zero_bin & negligible
            open_ = ~all_mono & ~zero_bin & ~overflow
            cut_head = open_ & monotone_after
            cut_tail = open_ & ~cut_head & turning
            forced = open_ & ~cut_head & ~cut_tail & bool(force_monotone)
            failed = open_ & ~cut_head & ~cut_tail & ~forced
            if bool(failed.any()):
                bad = rows[failed][0].item()
                raise ValueError("conditional_samples_from_cdf: the cdf of event %d (condition %g) is not monotone on "
                                 "[%g, %g] and neither catch_turnpoint nor force_monotone resolves it"
                                 % (bad, conds[bad].item(), lo[bad].item(), hi[bad].item()))
            draw = use_table | forced
            if bool(draw.any()):
                keep = torch.cat([mono, ones], dim=1) | use_table[:, None]      # forced: drop entries before a decrease
                keep = keep & torch.cat([ones.expand(-1, vals.shape[1] - 1), use_table[:, None]], dim=1)
                samples[rows[draw]] = invert_tables(vals[draw], pnts[draw], keep[draw], r[draw])
            samples[rows[zero_bin]] = 0.
            samples[rows[overflow]] = hi[rows[overflow]]
            if bool(cut_head.any()):
                lo[rows[cut_head]] = torch.gather(pnts, 1, first_above[:, None]).squeeze(1)[cut_head]
                still.append(rows[cut_head])
            if bool(cut_tail.any()):
                hi[rows[cut_tail]] = torch.gather(pnts, 1, first_bad[:, None]).squeeze(1)[cut_tail]
                still.append(rows[cut_tail])
        pending = torch.cat(still) if still else pending[:0]
    assert pending.numel() == 0, "conditional_samples_from_cdf: domains still shrinking after %d rounds" % max_rounds
    return samples, torch.ones_like(samples)


# ---- interpolant of a radiator table on the device (scipy RegularGridInterpolator, linear, extrapolating) --------
def grid_interpolant(xs, ys, zs, device=None):
    """z(x, y) by bilinear interpolation of the table zs[len(xs), len(ys)], linearly extrapolated outside the grid
    -- what ``get_2d_interpolation(xs, ys, zs)`` (utils/interpolation.py:248-273) returns, as a function of
    broadcastable tensors on ``device``."""
    to = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=device)
    gx, gy, table = to(xs), to(ys), to(zs)

    def locate(grid, v):
        i = torch.clamp(torch.searchsorted(grid, v.contiguous(), right=True) - 1, 0, grid.numel() - 2)
        t = (v - grid[i]) / (grid[i + 1] - grid[i])
        return i, t

    def interpolant(x, y):
        x, y = torch.broadcast_tensors(x, y)
        i, tx = locate(gx, x)
        j, ty = locate(gy, y)
        return (table[i, j] * (1 - tx) * (1 - ty) + table[i, j + 1] * (1 - tx) * ty
                + table[i + 1, j] * tx * (1 - ty) + table[i + 1, j + 1] * tx * ty)
    return interpolant


# ---- the three generators of the production workflow ---------------------------------------------------------------
def _device_uniforms(n, seed, device):
    """n uniforms of the counter RNG (sample index = event index) on ``device``"""
    if device.type != 'cuda':
        raise RuntimeError("jetmontecarlo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback for this "
                           "path.")
    from .. import _device as D
    from .. import _lib
    from ..montecarlo.sampler import next_key
    out = D.empty((n, 1))
    _lib.check(_lib.lib.jmc_uniforms(n, next_key() if seed is None else int(seed), 0, 1, D.ptr(out), D.stream()))
    return out[:, 0]


def _format(samples, weights):
    """inf -> 0 with zero weight.  ref: sudakov_utils.py:117-120"""
    bad = torch.isinf(samples)
    zero = torch.zeros_like(samples)
    return torch.where(bad, zero, samples), torch.where(bad, zero, weights)


def generate_precritical_samples(pre_radiator, theta_crits, z_cut, seed=None, uniforms=None):
    """z_pre for every critical angle, from the cdf exp(-R_pre(z_pre, theta_crit)) on [0, z_cut] forced monotone.
    ``pre_radiator(z_pre, theta)`` works on tensors (e.g. ``grid_interpolant`` of a gen_pre_num_rad result);
    ``theta_crits`` is a tensor on the device to work on.  Returns (samples, weights) tensors.
    ref: examples/sudakov_comparisons/sudakov_utils.py:134-182"""
    theta_crits = theta_crits.to(torch.float64)
    n = theta_crits.numel()
    r = _device_uniforms(n, seed, theta_crits.device) if uniforms is None else uniforms
    zeros = torch.zeros_like(theta_crits)
    samples, weights = conditional_samples_from_cdf(lambda z, th: torch.exp(-1. * pre_radiator(z, th)), theta_crits,
                                                    zeros, zeros + float(z_cut), r, force_monotone=True)
    return _format(samples, weights)


def generate_subsequent_samples(sub_radiator, theta_crits, beta, seed=None, uniforms=None):
    """C_sub for every critical angle, from exp(-R_sub(C, theta_crit)) on [0, theta_crit^beta / 2] forced monotone;
    events with theta_crit^beta / 2 < 1e-10 go to the underflow value 1e-100.
    ref: examples/sudakov_comparisons/sudakov_utils.py:185-238"""
    theta_crits = theta_crits.to(torch.float64)
    n = theta_crits.numel()
    r = _device_uniforms(n, seed, theta_crits.device) if uniforms is None else uniforms
    upper = theta_crits ** beta / 2.
    live = ~(upper < 1e-10)
    samples = torch.full_like(theta_crits, 1e-100)
    weights = torch.ones_like(theta_crits)
    if bool(live.any()):
        s, w = conditional_samples_from_cdf(lambda c, th: torch.exp(-1. * sub_radiator(c, th)), theta_crits[live],
                                            torch.zeros_like(upper[live]), upper[live], r[live], force_monotone=True)
        samples[live], weights[live] = s, w
    return _format(samples, weights)


def generate_critical_samples(crit_radiator, num_samples, seed=None):
    """theta_crit samples from exp(-R_crit(theta)) on [0, 1] (host-callable radiator, e.g. the integrator's
    interpolating function): ``samples_from_cdf`` with the draw on the device.  Returns NumPy arrays.
    ref: examples/sudakov_comparisons/sudakov_utils.py:100-131"""
    from ..utils.montecarlo_utils import samples_from_cdf
    samples, weights = samples_from_cdf(lambda theta: np.exp(-1. * crit_radiator(theta)), int(num_samples),
                                        domain=[0., 1.], backup_cdf=None, verbose=0, seed=seed)
    weights = np.where(np.isinf(samples), 0, weights)
    return np.where(np.isinf(samples), 0, samples), weights
