"""GPU tests of the "next" rows that were written after this round's GPU budget was spent: their
host logic and arithmetic are pinned on the CPU (tests/test_samples_from_cdf_cpu.py, tests/test_oracle_golden.py); these
run the same checks with the device in the loop."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("jet,acc", [('quark', 'MLL'), ('gluon', 'LL')])
def test_shower_rss_on_device_chains(cuda, jet, acc):
    """getECFs_rss (ref: partonshower_utils.py:501-674): chains regenerated on the GPU and groomed there, against the
    oracle (pinned bit for bit to the reference's getECFs_rss) applied to the same chains"""
    from jetmontecarlo_b200.utils import partonshower_utils as PS
    n = 2000
    jets = PS.gen_jets(n, beta=2., jet_type=jet, acc=acc, verbose=0, seed=77)
    lens, zs, ths = jets.chains()
    chains = [list(zip(zs[i, :lens[i]], ths[i, :lens[i]])) for i in range(n)]
    for z_cut, kw in ((.1, dict()), (.1, dict(n_emissions=2, obs_acc='MLL')), (.2, dict(n_emissions='all', f=.75)),
                      (.1, dict(f=.5, few_pres=False, n_emissions=2)), (.1, dict(emission_type='crit')),
                      (.2, dict(emission_type='critsub', n_emissions=3, obs_acc='MLL'))):
        args = dict(beta=2., f=1., obs_acc='LL', emission_type='precritsub', n_emissions=1, few_pres=True)
        args.update(kw)
        got = PS.getECFs_rss(jets, z_cut, **args)
        beta = args.pop('beta')
        ref = [O.ecf_from_chain_rss(c, z_cut, beta, **args) for c in chains]
        assert len(got) == n
        assert_allclose(got, ref, rtol=1e-12, err_msg=str(kw))
    # grooming only removes radiation: the RSS-groomed leading ECF never exceeds the ungroomed one
    ungroomed = np.array(PS.getECFs_ungroomed(jets, beta=2., obs_acc='LL', n_emissions=1))
    groomed = np.array(PS.getECFs_rss(jets, .1, beta=2., obs_acc='LL', n_emissions=1))
    assert np.all(groomed <= ungroomed * (1 + 1e-12)) and (groomed < ungroomed).mean() > .2
    assert PS.getECFs_rss(jets, 0, beta=2., n_emissions=2) == PS.getECFs_ungroomed(jets, beta=2., n_emissions=2)


def test_samples_from_cdf_on_device(cuda):
    """samples_from_cdf (ref: montecarlo_utils.py:93-375): table walk on the host, draw on the GPU; equal to the
    oracle (pinned to the reference's tabulated branches) on the same Philox uniforms"""
    from jetmontecarlo_b200.utils.montecarlo_utils import samples_from_cdf
    from test_oracle_golden import _cdf_cases
    n, seed = 100_000, 31
    rand = lambda m: O.philox_uniforms53(seed, 0, m, 1)[:, 0]
    cases = _cdf_cases()
    cases['sudakov_sub'] = (lambda c: np.exp(-O.rad_sub_fc(c, 2., 'quark')), [0., .5], {})
    for name in ('square', 'sudakov_pos_domain', 'neg_domain', 'negligible_then_monotone', 'turnpoint',
                 'wiggle_forced', 'sudakov_sub'):
        cdf, domain, kw = cases[name]
        got, w = samples_from_cdf(cdf, n, domain=domain, verbose=0, seed=seed, **kw)
        with np.errstate(all='ignore'):
            ref, _ = O.samples_from_cdf_tabulated(cdf, n, domain, rand, **kw)
        assert_allclose(got, ref, rtol=1e-12, atol=1e-15, err_msg=name)
        assert np.all(w == 1.) and len(got) == n
    # the samples follow the cdf they came from (Kolmogorov distance of 1e5 samples)
    got, _ = samples_from_cdf(cases['square'][0], n, domain=[0, 1], verbose=0, seed=seed)
    xs = np.sort(got)
    assert np.abs(xs ** 2 - (np.arange(n) + .5) / n).max() < 1e-2
    zero, _ = samples_from_cdf(cases['starts_at_one'][0], 10, domain=[0, 1], verbose=0, seed=seed)
    assert np.all(zero == 0.)
