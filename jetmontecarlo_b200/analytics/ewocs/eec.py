"""Normalisation of the LO e+ e- -> q qbar g matrix element -- host mirror of the one function of
jetmontecarlo/analytics/ewocs/eec.py that the Monte Carlo path uses (the analytic EEC curves of that module are
plot-side truth curves, out of scope: SURVEY.md section 2).  ref: analytics/ewocs/eec.py:9-19"""
from ..qcd_utils import alpha1loop, alpha_em_1loop, sum_square_quark_charges


def ee2hadrons_LO_prefactor(energy):
    """8 E^2 alpha_s(E) alpha_em(E)^2 sum_q e_q^2 (a per-process constant, computed once on the host and handed to
    the kernels in ``jmc_process.prefactor``)."""
    alpha_s = alpha1loop(energy)
    prefactor = 8 * energy**2. * alpha_s * alpha_em_1loop(energy)**2.\
        * sum_square_quark_charges(energy)
    return prefactor
