"""The ECF pickers of the shower (ungroomed, Soft Drop, recursive safe subtraction) on the REFERENCE's own event
records: the golden (z, theta) chains of gen_jets are uploaded and walked by jmc_chain_ecf -- the accumulator the shower
kernel carries per lane -- and compared with what the reference's getECFs_* returned for those jets.
ref: utils/partonshower_utils.py:445-499, 501-674, 676-790"""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu


def _padded(g, cap=None):
    lens = g['lens']
    L = int(lens.max()) if cap is None else cap
    c = np.full((len(lens), L, 2), np.nan)
    ends = np.cumsum(lens)
    for i, (e, n) in enumerate(zip(ends, lens)):
        c[i, :n, 0], c[i, :n, 1] = g['z'][e - n:e], g['theta'][e - n:e]
    return torch.from_numpy(lens.astype(np.int32)).cuda(), torch.from_numpy(c).cuda()


@pytest.mark.parametrize("jet,acc", [('quark', 'LL'), ('quark', 'MLL'), ('gluon', 'LL'), ('gluon', 'MLL')])
def test_rss_on_the_reference_chains(cuda, golden, jet, acc):
    from jetmontecarlo_b200.utils.partonshower_utils import ecfs_from_chains
    chains, g = golden(f"shower_{jet}_{acc}"), golden(f"shower_rss_{jet}_{acc}")
    lens, c = _padded(chains)
    keys = [k for k in g if k.startswith('rss:')]
    assert len(keys) == 24
    for k in keys:
        _, z_cut, f, oacc, etype, nem, few = k.split(':')
        nem = nem if nem == 'all' else int(nem)
        got = ecfs_from_chains(lens, c, beta=2., obs_acc=oacc, n_emissions=nem, groomer='rss', z_cut=float(z_cut),
                               f=float(f), emission_type=etype, few_pres=bool(int(few))).cpu().numpy()
        # Same operation order as the reference, so nearly every jet agrees to the last bit.  The exceptions: the
        # reference squares the hard-branch scale with C pow() in one branch, and with few_pres=False it forms
        # z_cut - sum(z_pres) with Python's sum(), which is a compensated sum from Python 3.12 on (the vectors were
        # generated with 3.12) -- a last-place difference that the cancellation in z_cut - sum can amplify a
        # hundredfold.
        assert_allclose(got, g[k], rtol=1e-14 if int(few) else 1e-12, atol=0, err_msg=k)
        assert (got == g[k]).mean() > .95, k


@pytest.mark.parametrize("jet,acc", [('quark', 'MLL'), ('gluon', 'LL')])
def test_ungroomed_and_softdrop_on_the_reference_chains(cuda, golden, jet, acc):
    """the two other pickers through the same replay entry point, against the oracle on the reference's chains (the
    oracle is pinned to the reference's getECFs_ungroomed / getECFs_softdrop in tests/test_oracle_golden.py)"""
    from jetmontecarlo_b200.utils.partonshower_utils import ecfs_from_chains
    chains = golden(f"shower_{jet}_{acc}")
    lens, c = _padded(chains, cap=int(chains['lens'].max()) + 3)          # a record with spare room
    ends = np.cumsum(chains['lens'])
    lists = [list(zip(chains['z'][e - n:e], chains['theta'][e - n:e])) for e, n in zip(ends, chains['lens'])]
    for oacc in ('LL', 'MLL'):
        for nem in (1, 2, 3, 'all'):
            got = ecfs_from_chains(lens, c, beta=2., obs_acc=oacc, n_emissions=nem).cpu().numpy()
            ref = np.array([O.ecf_from_chain_ungroomed(ch, 2., oacc, nem) for ch in lists])
            assert_allclose(got, ref, rtol=1e-14, err_msg=f"ungroomed {oacc} {nem}")
            for z_cut, beta_sd in ((.1, 0.), (.05, 1.)):
                got = ecfs_from_chains(lens, c, beta=2., obs_acc=oacc, n_emissions=nem, groomer='softdrop', z_cut=z_cut,
                                       beta_sd=beta_sd).cpu().numpy()
                ref = np.array([O.ecf_from_chain_softdrop(ch, z_cut, 2., beta_sd, oacc, nem) for ch in lists])
                assert_allclose(got, ref, rtol=1e-14, err_msg=f"softdrop {oacc} {nem} {z_cut}")


def test_rss_edge_cases(cuda):
    from jetmontecarlo_b200.utils.partonshower_utils import ecfs_from_chains
    nan = float('nan')
    lens = torch.tensor([0, 1, 3, 2], dtype=torch.int32)
    z = torch.tensor([[nan, nan, nan], [.3, nan, nan], [.02, .05, .2], [.01, .02, nan]], dtype=torch.float64)
    th = torch.tensor([[nan, nan, nan], [.5, nan, nan], [.9, .6, .3], [.8, .4, nan]], dtype=torch.float64)
    c = torch.stack([z, th], dim=2)
    for kw in (dict(n_emissions=1), dict(n_emissions=3), dict(n_emissions='all'), dict(f=.6, few_pres=False),
               dict(obs_acc='MLL', emission_type='crit'), dict(beta=1.5, n_emissions=2)):
        got = ecfs_from_chains(lens, c, groomer='rss', z_cut=.1, **kw).cpu().numpy()
        want = []
        for i in range(4):
            chain = [(float(z[i, j]), float(th[i, j])) for j in range(int(lens[i]))]
            args = dict(beta=2., f=1., obs_acc='LL', emission_type='precritsub', n_emissions=1, few_pres=True)
            args.update(kw)
            want.append(O.ecf_from_chain_rss(chain, .1, args.pop('beta'), **args))
        assert_allclose(got, np.array(want), rtol=1e-14, err_msg=str(kw))
    # a jet without emissions, and one whose emissions are all groomed away, keep the reference's 1e-50 floors
    got = ecfs_from_chains(lens, c, groomer='rss', z_cut=.1, n_emissions='all').cpu().numpy()
    assert got[0] == 2e-50 and got[3] == 2e-50
    with pytest.raises(AssertionError, match="less than 1/3"):
        ecfs_from_chains(lens, c, groomer='rss', z_cut=.1, f=.2)
    with pytest.raises(AssertionError, match="truncated"):
        ecfs_from_chains(torch.tensor([5], dtype=torch.int32), c[:1], groomer='rss', z_cut=.1)
    # many jets: every warp-tile path (jets beyond a multiple of 32, chains longer than one 8-emission tile)
    rng = np.random.RandomState(3)
    n, L = 1000, 21
    lens_np = rng.randint(0, L + 1, n).astype(np.int32)
    zz, tt = rng.uniform(1e-3, .5, (n, L)), rng.uniform(1e-3, 1., (n, L))
    cc = torch.from_numpy(np.stack([zz, tt], axis=2))
    got = ecfs_from_chains(torch.from_numpy(lens_np), cc, groomer='rss', z_cut=.1, f=.75, obs_acc='MLL',
                           n_emissions=2).cpu().numpy()
    ref = [O.ecf_from_chain_rss(list(zip(zz[i, :lens_np[i]], tt[i, :lens_np[i]])), .1, 2., f=.75, obs_acc='MLL',
                                n_emissions=2) for i in range(n)]
    ref = np.array([np.nan if r is None else r for r in ref])
    keep = ~np.isnan(ref)
    assert_allclose(got[keep], ref[keep], rtol=1e-13)
    assert np.all(got[~keep] < 0)


def test_getECFs_rss_is_the_chain_picker_on_the_generated_jets(cuda):
    """getECFs_rss (grooming inside the shower kernel, nothing stored) == the replay picker on the event record of the
    same jets; the ecf >= 0 filter; the z_cut = 0 shortcut"""
    from jetmontecarlo_b200.utils import partonshower_utils as PS
    jets = PS.gen_jets(3000, beta=2., jet_type='gluon', acc='MLL', verbose=0, seed=5)
    lens, chains = jets.chains_device(64)
    for kw in (dict(f=.75, obs_acc='MLL', emission_type='precritsub', n_emissions='all'),
               dict(f=1., obs_acc='LL', emission_type='crit', n_emissions=1),
               dict(f=.75, obs_acc='LL', emission_type='precritsub', n_emissions=2, few_pres=False)):
        got = np.array(PS.getECFs_rss(jets, .1, beta=2., **kw))
        ref = PS.ecfs_from_chains(lens, chains, beta=2., groomer='rss', z_cut=.1, **kw).cpu().numpy()
        assert_allclose(got, ref[ref >= 0], rtol=0, atol=0)
    assert PS.getECFs_rss(jets, 0, beta=2., obs_acc='MLL', n_emissions=2) == \
        PS.getECFs_ungroomed(jets, beta=2., obs_acc='MLL', n_emissions=2)
