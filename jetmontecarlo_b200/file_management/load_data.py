"""Loaders of the production workflow's catalogued products (radiator interpolants, splitting functions, Sudakov and
shower samples).  ref: file_management/load_data.py:1-195.  The reference takes its default parameter dictionaries and
z_cut / beta lists from ``examples/params.py``; here they are arguments (with the same defaults when that module is
importable), since the parameter file belongs to the user's scripts, not to this package."""
from .data_management import load_data, load_and_interpolate, load_and_tabulate


def _defaults():
    """(radiator params, splitting-function params, shower params, z_cuts, betas) of examples/params.py, if present"""
    try:
        from examples import params as P
    except Exception:      # no parameter module on the path: every loader needs its arguments
        return None
    rad = dict(P.RADIATOR_PARAMS); rad.pop('z_cut', None); rad.pop('beta', None)
    split = dict(P.SPLITTINGFN_PARAMS); split.pop('z_cut', None)
    shower = dict(P.SHOWER_PARAMS); shower.pop('shower beta', None)
    return rad, split, shower, list(P.Z_CUTS), list(P.BETAS)


def _need(value, index, name):
    if value is not None:
        return value
    d = _defaults()
    if d is None:
        raise ValueError("%s is required (no examples.params module to take the reference's defaults from)" % name)
    return d[index]


def load_sudakov_samples(sudakov_params=None, z_cuts=None, betas=None,
                         emissions=('critical', 'pre-critical', 'subsequent')):
    """{emission: {z_cut: [{beta:}] samples}} of the stored inverse-transform samples.  ref: :26-82"""
    sudakov_params, z_cuts, betas = _need(sudakov_params, 0, 'sudakov_params'), _need(z_cuts, 3, 'z_cuts'), _need(betas, 4, 'betas')
    samples = {emission: {} for emission in emissions}
    for z_cut in z_cuts:
        if 'critical' in emissions:
            samples['critical'][z_cut] = {
                b: load_data('montecarlo samples', 'critical sudakov inverse transform',
                             params=dict(**sudakov_params, **{'z_cut': z_cut})) for b in betas}
        if 'pre-critical' in emissions:
            samples['pre-critical'][z_cut] = load_data('montecarlo samples', 'pre-critical sudakov inverse transform',
                                                       params=dict(**sudakov_params, **{'z_cut': z_cut}))
        if 'subsequent' in emissions:
            samples['subsequent'][z_cut] = {
                b: load_data('montecarlo samples', 'subsequent sudakov inverse transform',
                             params=dict(**sudakov_params, **{'z_cut': z_cut, 'beta': b})) for b in betas}
    return samples


def load_partonshower_samples(groomer, n_emissions, shower_params=None, z_cuts=None, betas=None, **kwargs):
    """{z_cut: {beta: samples}} of stored parton-shower correlations.  ref: :85-107"""
    shower_params, z_cuts, betas = _need(shower_params, 2, 'shower_params'), _need(z_cuts, 3, 'z_cuts'), _need(betas, 4, 'betas')
    return {z_cut: {b: load_data('montecarlo samples', 'parton shower',
                                 params=dict(**shower_params, **{'shower beta': b, 'groomer': groomer,
                                                                 'number of emissions': n_emissions, 'z_cut': z_cut},
                                             **kwargs))
                    for b in betas} for z_cut in z_cuts}


def load_radiators(radiator_params=None, z_cuts=None, betas=None,
                   emissions=('critical', 'pre-critical', 'subsequent'), as_tables=False):
    """Stored numerical radiators as interpolating functions: critical (1-D, monotone on [1e-10, 1]) and pre-critical
    (2-D, rectangular grid) per z_cut, subsequent (2-D, nearest) per beta.  ``as_tables``: the 2-D ones as
    ``TableRadiator`` for the device's conditional sampler instead of host interpolants.  ref: :114-165"""
    radiator_params, z_cuts, betas = _need(radiator_params, 0, 'radiator_params'), _need(z_cuts, 3, 'z_cuts'), _need(betas, 4, 'betas')
    radiators = {emission: {} for emission in emissions}
    two_d = ((lambda source, params, method: load_and_tabulate(source, params, method)) if as_tables else
             (lambda source, params, method: load_and_interpolate(source, params, interpolation_method=method)))
    if 'critical' in emissions:
        for z_cut in z_cuts:
            radiators['critical'][z_cut] = load_and_interpolate(
                'critical radiator', params=dict(**radiator_params, **{'z_cut': z_cut}), monotonic=True, bounds=(1e-10, 1))
    if 'pre-critical' in emissions:
        for z_cut in z_cuts:
            radiators['pre-critical'][z_cut] = two_d('pre-critical radiator', dict(**radiator_params, **{'z_cut': z_cut}),
                                                     "RectangularGrid")
    if 'subsequent' in emissions:
        for b in betas:
            radiators['subsequent'][b] = two_d('subsequent radiator', dict(**radiator_params, **{'beta': b}), 'Nearest')
    return radiators


def load_splittingfns(splittingfn_params=None, z_cuts=None):
    """{z_cut: normalised splitting function}.  ref: :172-195"""
    splittingfn_params, z_cuts = _need(splittingfn_params, 1, 'splittingfn_params'), _need(z_cuts, 3, 'z_cuts')
    return {z_cut: load_data('serialized function', 'splitting function',
                             params=dict(**splittingfn_params, **{'z_cut': z_cut})) for z_cut in z_cuts}
