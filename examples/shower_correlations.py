"""Veto-algorithm parton shower and its correlations file.
ref: examples/event_generation/parton_shower_gen.py (gen_jets + save_shower_correlations)"""
import numpy as np

from jetmontecarlo_b200.utils.partonshower_utils import gen_jets, getECFs_softdrop, save_shower_correlations

NUM_EVENTS, BETA = int(1e6), 2.
for jet_type in ('quark', 'gluon'):
    jets = gen_jets(NUM_EVENTS, beta=BETA, jet_type=jet_type, acc='MLL', verbose=0, seed=1)
    n_em, n_trials = jets.emission_counts()
    print("%-5s %d jets: %.2f emissions and %.1f veto trials per jet" % (jet_type, len(jets), n_em.mean(), n_trials.mean()))
    path = "shower_correlations_%s.npz" % jet_type
    save_shower_correlations(jets, path, z_cuts=[.05, .1, .2], beta=BETA, f_soft=1., obs_acc='MLL', verbose=0)
    saved = np.load(path)
    print("      wrote", path, {k: saved[k].shape for k in saved.files})
    c1 = np.array(getECFs_softdrop(jets, .1, beta=BETA, obs_acc='MLL', n_emissions=1))
    print("      Soft Drop (z_cut = .1) leading ECF: median %.3e, zero-bin fraction %.3f"
          % (np.median(c1), (c1 < 1e-40).mean()))
