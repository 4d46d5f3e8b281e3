"""RSS grooming of shower chains (rss_from_chains: device-agnostic tensor code that runs on the GPU in the
product) against the reference's getECFs_rss, on the reference's own chains -- CPU tensors here.
ref: utils/partonshower_utils.py:501-674"""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from oracle import jmc_oracle as O


def _padded(g):
    lens = g['lens']
    L = int(lens.max())
    z = np.full((len(lens), L), np.nan)
    th = np.full((len(lens), L), np.nan)
    ends = np.cumsum(lens)
    for i, (e, n) in enumerate(zip(ends, lens)):
        z[i, :n], th[i, :n] = g['z'][e - n:e], g['theta'][e - n:e]
    return torch.from_numpy(lens.astype(np.int32)), torch.from_numpy(z), torch.from_numpy(th)


@pytest.mark.parametrize("jet,acc", [('quark', 'LL'), ('quark', 'MLL'), ('gluon', 'LL'), ('gluon', 'MLL')])
def test_rss_from_chains_is_the_reference(golden, jet, acc):
    from jetmontecarlo_b200.utils.partonshower_utils import rss_from_chains
    chains, g = golden(f"shower_{jet}_{acc}"), golden(f"shower_rss_{jet}_{acc}")
    lens, z, th = _padded(chains)
    keys = [k for k in g if k.startswith('rss:')]
    assert len(keys) == 24
    for k in keys:
        _, z_cut, f, oacc, etype, nem, few = k.split(':')
        nem = nem if nem == 'all' else int(nem)
        got = rss_from_chains(lens, z, th, float(z_cut), beta=2., f=float(f), obs_acc=oacc, emission_type=etype,
                              n_emissions=nem, few_pres=bool(int(few))).numpy()
        # Same operation order as the reference, so nearly every jet agrees to the last bit.  The exceptions: the
        # reference squares the hard-branch scale with C pow() in one branch, and with few_pres=False it forms
        # z_cut - sum(z_pres) with Python's sum(), which is a compensated sum from Python 3.12 on (the vectors were
        # generated with 3.12) -- a last-place difference that the cancellation in z_cut - sum can amplify a
        # hundredfold.
        assert_allclose(got, g[k], rtol=1e-14 if int(few) else 1e-12, atol=0, err_msg=k)
        assert (got == g[k]).mean() > .95, k


def test_rss_edge_cases():
    from jetmontecarlo_b200.utils.partonshower_utils import rss_from_chains
    nan = float('nan')
    lens = torch.tensor([0, 1, 3, 2], dtype=torch.int32)
    z = torch.tensor([[nan, nan, nan], [.3, nan, nan], [.02, .05, .2], [.01, .02, nan]], dtype=torch.float64)
    th = torch.tensor([[nan, nan, nan], [.5, nan, nan], [.9, .6, .3], [.8, .4, nan]], dtype=torch.float64)
    for kw in (dict(n_emissions=1), dict(n_emissions=3), dict(n_emissions='all'), dict(f=.6, few_pres=False),
               dict(obs_acc='MLL', emission_type='crit'), dict(beta=1.5, n_emissions=2)):
        got = rss_from_chains(lens, z, th, .1, **kw).numpy()
        want = []
        for i in range(4):
            chain = [(float(z[i, j]), float(th[i, j])) for j in range(int(lens[i]))]
            args = dict(beta=2., f=1., obs_acc='LL', emission_type='precritsub', n_emissions=1, few_pres=True)
            args.update(kw)
            want.append(O.ecf_from_chain_rss(chain, .1, args.pop('beta'), **args))
        assert_allclose(got, np.array(want), rtol=1e-14, err_msg=str(kw))
    # a jet without emissions, and one whose emissions are all groomed away, keep the reference's 1e-50 floors
    got = rss_from_chains(lens, z, th, .1, n_emissions='all').numpy()
    assert got[0] == 2e-50 and got[3] == 2e-50
    with pytest.raises(AssertionError, match="less than 1/3"):
        rss_from_chains(lens, z, th, .1, f=.2)
    with pytest.raises(AssertionError, match="truncated"):
        rss_from_chains(torch.tensor([5], dtype=torch.int32), z[:1], th[:1], .1)


def test_getECFs_rss_glue(golden, monkeypatch):
    """getECFs_rss end to end with the device call replaced by the reference's chains: argument plumbing, the
    regenerate-with-a-longer-record loop, the ecf >= 0 filter, and the z_cut = 0 shortcut"""
    from jetmontecarlo_b200 import _device as D
    from jetmontecarlo_b200.utils import partonshower_utils as PS
    chains, g = golden("shower_gluon_MLL"), golden("shower_rss_gluon_MLL")
    lens, z, th = _padded(chains)
    calls = []

    def fake_chains_device(self, max_emissions=64, precision='f64'):
        calls.append(max_emissions)
        L = min(int(max_emissions), z.shape[1])
        c = torch.full((len(lens), int(max_emissions), 2), float('nan'), dtype=torch.float64)
        c[:, :L, 0], c[:, :L, 1] = z[:, :L], th[:, :L]
        return lens, c

    monkeypatch.setattr(PS.JetList, "chains_device", fake_chains_device)
    monkeypatch.setattr(D, "to_host", lambda t: t.detach().cpu().numpy())
    jets = PS.JetList(len(lens), 2., 'gluon', 1., 'MLL', O.MU_NP, seed=1)
    got = PS.getECFs_rss(jets, .1, beta=2., f=.75, obs_acc='MLL', emission_type='precritsub', n_emissions='all')
    assert isinstance(got, list) and len(got) == len(lens)
    assert_allclose(np.array(got), g['rss:0.1:0.75:MLL:precritsub:all:1'], rtol=1e-14)
    assert calls == [64]
    got = PS.getECFs_rss(jets, .2, n_emissions=2, few_pres=False, f=.75)
    assert_allclose(np.array(got), g['rss:0.2:0.75:LL:precritsub:2:0'], rtol=1e-12)
    # a record shorter than the longest chain is regenerated with more room
    calls.clear()
    monkeypatch.setattr(PS, "CHAIN_RECORD", 8)
    got = PS.getECFs_rss(jets, .1)
    assert calls == [8, 16, 32, 64][:len(calls)] and calls[-1] >= int(lens.max()) > calls[-2]
    assert_allclose(np.array(got), g['rss:0.1:1.0:LL:precritsub:1:1'], rtol=1e-14)
    # z_cut = 0 is the ungroomed observable
    seen = {}
    monkeypatch.setattr(PS, "getECFs_ungroomed", lambda jl, **kw: seen.update(kw) or ['ungroomed'])
    assert PS.getECFs_rss(jets, 0, beta=2., obs_acc='MLL', n_emissions=2) == ['ungroomed']
    assert seen == dict(beta=2., obs_acc='MLL', n_emissions=2, verbose=0)


def test_save_shower_correlations_file_format(golden, monkeypatch, tmp_path):
    """keys, shapes and values of the production script's correlations file (partonshower_utils.py:802-884), with
    the three device-backed extractors replaced by the oracle on the reference's chains"""
    from jetmontecarlo_b200.utils import partonshower_utils as PS
    chains_file, g = golden("shower_quark_MLL"), golden("shower_corr_quark_MLL")
    ends = np.cumsum(chains_file['lens'])
    chains = [list(zip(chains_file['z'][e - n:e], chains_file['theta'][e - n:e])) for e, n in zip(ends, chains_file['lens'])]
    monkeypatch.setattr(PS, "getECFs_ungroomed", lambda jet_list, beta=2., obs_acc='LL', n_emissions=1, verbose=0:
                        [O.ecf_from_chain_ungroomed(c, beta, obs_acc, n_emissions) for c in chains])
    monkeypatch.setattr(PS, "getECFs_softdrop", lambda jet_list, z_cut, beta=2., beta_sd=0., obs_acc='LL', n_emissions=1,
                        verbose=0: [O.ecf_from_chain_softdrop(c, z_cut, beta, beta_sd, obs_acc, n_emissions) for c in chains])
    monkeypatch.setattr(PS, "getECFs_rss", lambda jet_list, z_cut, beta=2., f=1., obs_acc='LL', emission_type='precritsub',
                        n_emissions=1, few_pres=True, verbose=0:
                        [O.ecf_from_chain_rss(c, z_cut, beta, f=f, obs_acc=obs_acc, emission_type=emission_type,
                                              n_emissions=n_emissions, few_pres=few_pres) for c in chains])
    path = str(tmp_path / "corr.npz")
    PS.save_shower_correlations(None, path, z_cuts=[.05, .1, .2], beta=2., f_soft=1., obs_acc='MLL', few_pres=True,
                                verbose=0)
    z = np.load(path)
    assert set(z.files) == set(g)
    for k in g:
        assert z[k].shape == g[k].shape, k
        if k == 'softdrop_c1s_two':
            continue
        # chains recomputed from rotated 3-vectors in the reference (theta to ~1e-5, see test_shower_replay) do not
        # enter here: both sides use the reference's own (z, theta)
        assert_allclose(z[k], g[k], rtol=1e-13, err_msg=k)
    # Soft Drop with two emissions: the reference decrements n_emissions in place (:765), so only its FIRST jet uses
    # two emissions and the rest fall back to one; the per-jet rule of its docstring is what is implemented here
    assert_allclose(z['softdrop_c1s_two'][:, 0, 0], g['softdrop_c1s_two'][:, 0, 0], rtol=1e-13)
    assert_allclose(g['softdrop_c1s_two'][:, 0, 1:], g['softdrop_c1s_crit'][:, 0, 1:], rtol=0)
    assert np.all(z['softdrop_c1s_two'] >= z['softdrop_c1s_crit'])
    # dispatcher
    assert PS.getECFs(None, dict(beta=2., obs_acc='LL', n_emissions=1)) == PS.getECFs_ungroomed(None)
    assert PS.getECFs(None, dict(z_cut=.1), groomer='RSS') == PS.getECFs_rss(None, .1)
    assert PS.getECFs(None, dict(z_cut=.1), groomer='SD') == PS.getECFs_softdrop(None, .1)
    with pytest.raises(ValueError, match="Invalid groomer type mMDT."):
        PS.getECFs(None, {}, groomer='mMDT')
    assert PS.getangs(None, beta=2., acc='MLL', verbose=0) == PS.getECFs_ungroomed(None, obs_acc='LL')
