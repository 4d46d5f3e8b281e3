import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gpu_helpers as H
import test_gpu_fast as T
np.set_printoptions(precision=9, linewidth=200)
for kind, jet, fc, acc, eps in [('crit_sub','gluon',True,'LL',1e-10), ('crit_sub','quark',True,'MLL',1e-10), ('pre_crit_sub','quark',True,'LL',1e-10), ('crit_sub','quark',True,'MLL',1e-3)]:
    ig = T._integrand(kind, jet, fc, acc, eps)
    n = 400000
    samples, obs, w = H.dump(ig, n, 31337, 'f32')
    obs_ref, w_ref = T._oracle_on_dump(kind, samples, jet, fc, acc, eps=eps)
    rel = np.abs(obs/obs_ref - 1)
    bad = np.argsort(-np.nan_to_num(rel))[:6]
    print("==", kind, jet, fc, acc, eps, "obs rel>2e-5:", (rel > 2e-5).sum())
    for i in bad:
        print(" obs", obs[i], obs_ref[i], "samples", samples[i])
    good = ~np.isnan(w_ref) & (w != 0) & (np.abs(w_ref) > 1e-30)
    relw = np.where(good, np.abs(w/np.where(good, w_ref, 1) - 1), 0)
    bad = np.argsort(-relw)[:6]
    print(" w rel>2e-4:", (relw > 2e-4).sum())
    for i in bad:
        print(" w", w[i], w_ref[i], relw[i], "samples", samples[i])
