"""Host glue of the shower observables: the correlations file of the production script and the getECFs dispatcher.
(The pickers themselves -- ungroomed, Soft Drop, RSS -- run in the shower kernel's per-jet accumulator; they are held to
the reference's own chains on the GPU, tests/test_gpu_chain_ecf.py.)  ref: utils/partonshower_utils.py:792-884"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import jmc_oracle as O


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
