"""Generate tests/golden/*.npz by EXECUTING the unmodified reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The vectors are small on purpose (a few thousand samples per case): they pin
the oracle (oracle/jmc_oracle.py) bit for bit and serve as fixed-sample replay
inputs for the CUDA parity tests on the GPU box, where the reference is absent.
Seed convention: ``random.seed(20260101 + case)`` right before each sampler run.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

ref_shim.install()

from jetmontecarlo.analytics import qcd_utils as Q  # noqa: E402
from jetmontecarlo.analytics.radiators import fixedcoupling as FC  # noqa: E402
from jetmontecarlo.analytics.radiators import running_coupling as RC  # noqa: E402
from jetmontecarlo.analytics.sudakov_factors.fixedcoupling import critSudakov_fc_LL  # noqa: E402
from jetmontecarlo.montecarlo.integrator import integrator, integrate_1d  # noqa: E402
from jetmontecarlo.montecarlo.sampler import simpleSampler  # noqa: E402
from jetmontecarlo.numerics import observables as OBS  # noqa: E402
from jetmontecarlo.numerics import weights as W  # noqa: E402
from jetmontecarlo.numerics.radiators.samplers import (  # noqa: E402
    criticalSampler, precriticalSampler, ungroomedSampler)
from jetmontecarlo.utils import partonshower_utils as PS  # noqa: E402
from jetmontecarlo.utils.interpolation import lin_log_mixed_list  # noqa: E402
from jetmontecarlo.utils.hist_utils import vals_to_pdf  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 20260101
np.seterr(all='ignore')


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def integ_arrays(it):
    return dict(edges=np.asarray(it.bins), density=np.asarray(it.density),
                density_err=np.asarray(it.densityErr),
                integral=np.asarray(it.integral, dtype=float),
                integral_err=np.asarray(it.integralErr))


# ---------------------------------------------------------------- KATs ----
def kats():
    k = {}
    k['alpha_fixed'] = Q.alpha_fixed
    k['beta_0'] = Q.beta_0
    k['MU_NP'] = Q.MU_NP
    k['alpha_s(.2,.3)'] = Q.alpha_s(.2, .3)
    k['alpha_s(1e-3,1e-3)'] = Q.alpha_s(1e-3, 1e-3)
    for jet in ('quark', 'gluon'):
        j = jet[0]
        k[j + ':splittingFn(.2,LL)'] = Q.splittingFn(.2, jet, 'LL')
        k[j + ':splittingFn(.2,LL_nonsing)'] = Q.splittingFn(.2, jet, 'LL_nonsing')
        k[j + ':splittingFn(.2,MLL)'] = Q.splittingFn(.2, jet, 'MLL')
        k[j + ':radiatorWeight(.2,.3,fc,LL)'] = W.radiatorWeight(.2, .3, jet, True, 'LL')
        k[j + ':radiatorWeight(.2,.3,rc,MLL)'] = W.radiatorWeight(.2, .3, jet, False, 'MLL')
        k[j + ':critPDFAnalytic_fc_LL(.2,.3,.1)'] = FC.critPDFAnalytic_fc_LL(.2, .3, z_c=.1, jet_type=jet)
        k[j + ':critRadAnalytic(.3,.1)'] = RC.critRadAnalytic(.3, .1, jet)
        k[j + ':critRadAnalytic(2e-3,.1)'] = RC.critRadAnalytic(2e-3, .1, jet)
        k[j + ':critRadAnalytic(1e-4,.1)'] = RC.critRadAnalytic(1e-4, .1, jet)
        k[j + ':critPDFAnalytic(.2,.3,.1)'] = RC.critPDFAnalytic(.2, .3, .1, jet)
        k[j + ':critRadPrimeAnalytic(.3,.1)'] = RC.critRadPrimeAnalytic(.3, .1, jet)
        k[j + ':subRadAnalytic_fc_LL(.01,2)'] = FC.subRadAnalytic_fc_LL(.01, 2, jet)
        k[j + ':subPDFAnalytic_fc_LL(.01,2)'] = FC.subPDFAnalytic_fc_LL(.01, 2, jet)
        k[j + ':subPDFAnalytic_fc_LL(.01,2,R=.3)'] = FC.subPDFAnalytic_fc_LL(.01, 2, jet, maxRadius=.3)
        k[j + ':subRadAnalytic(.01,2)'] = RC.subRadAnalytic(.01, 2, jet)
        k[j + ':subRadAnalytic(1e-5,2)'] = RC.subRadAnalytic(1e-5, 2, jet)
        k[j + ':subRadPrimeAnalytic(.01,2)'] = RC.subRadPrimeAnalytic(.01, 2, jet)
        k[j + ':subPDFAnalytic(.01,2,R=.3)'] = RC.subPDFAnalytic(.01, 2, jet, maxRadius=.3)
        k[j + ':critSudakov_fc_LL(.01,.1,2)'] = float(critSudakov_fc_LL(.01, .1, 2, jet))
        k[j + ':precriticalEmissionWeight(.03,.3,.1)'] = W.precriticalEmissionWeight(.03, .3, .1, jet)
        c = np.array([[.2, .3]])
        s = np.array([[.05, .4]])
        k[j + ':twoEmissionWeight(fc)'] = W.twoEmissionWeight(c, s, .1, 2., jet, True)[0]
        k[j + ':twoEmissionWeight(rc)'] = W.twoEmissionWeight(c, s, .1, 2., jet, False)[0]
    k['C_groomed(.2,.3,.1,2)'] = OBS.C_groomed(.2, .3, .1, 2)
    k['C_groomed(.2,.3,.1,2,zpre=.03,f=1,MLL)'] = OBS.C_groomed(.2, .3, .1, 2, z_pre=.03, f=1, acc='MLL')
    k['C_groomed(.2,.3,.1,2,zpre=.1,f=0,MLL)'] = OBS.C_groomed(.2, .3, .1, 2, z_pre=.1, f=0, acc='MLL')
    k['C_ungroomed(.2,.3,2,MLL)'] = OBS.C_ungroomed(.2, .3, 2, 'MLL')
    k['landau'] = float(np.exp(-1 / (2 * Q.alpha_fixed * Q.beta_0)))
    names = np.array(sorted(k))
    vals = np.array([float(k[n]) for n in names])
    save("kats", names=names, values=vals)


# ------------------------------------------------- function-level vectors --
def function_vectors():
    """Every analytic function on broad log-uniform inputs that reach all
    regions, including the NaN (Landau) ones."""
    rng = np.random.RandomState(SEED)
    n = 1200
    z = np.exp(rng.uniform(np.log(1e-12), np.log(.5), n))
    th = np.exp(rng.uniform(np.log(1e-12), 0., n))
    C = np.exp(rng.uniform(np.log(1e-14), np.log(.7), n))
    zpre = np.exp(rng.uniform(np.log(1e-12), np.log(.12), n))
    maxr = np.exp(rng.uniform(np.log(1e-6), 0., n))
    # put a few values exactly on thresholds / edges
    z[:4] = [.1, .5, .25, 1e-300]
    th[:4] = [1., 2 * Q.MU_NP, Q.MU_NP / .1, 1e-300]
    C[:4] = [.5, Q.MU_NP, Q.MU_NP ** 2 * 2, 0.]
    out = dict(z=z, theta=th, C=C, z_pre=zpre, max_radius=maxr)
    out['alpha_s'] = Q.alpha_s(z, th)
    for jet in ('quark', 'gluon'):
        j = jet[0]
        for acc in ('LL', 'LL_nonsing', 'MLL'):
            out[f'{j}:splittingFn:{acc}'] = Q.splittingFn(z, jet, acc)
            for fc in (True, False):
                out[f'{j}:radiatorWeight:{acc}:{int(fc)}'] = W.radiatorWeight(z, th, jet, fc, acc)
        for zc in (.05, .1, .2):
            out[f'{j}:critPDF_fc:{zc}'] = FC.critPDFAnalytic_fc_LL(z, th, zc, jet)
            out[f'{j}:critPDF_rc:{zc}'] = RC.critPDFAnalytic(z, th, zc, jet)
            out[f'{j}:critRad_rc:{zc}'] = RC.critRadAnalytic(th, zc, jet)
            out[f'{j}:critRadPrime_rc:{zc}'] = RC.critRadPrimeAnalytic(th, zc, jet)
            out[f'{j}:critRad_fc:{zc}'] = FC.critRadAnalytic_fc_LL(th, zc, jet)
            out[f'{j}:preWeight:{zc}'] = W.precriticalEmissionWeight(zpre, th, zc, jet)
            out[f'{j}:preRad_fc:{zc}'] = FC.preRadAnalytic_fc_LL(zpre, th, zc, jet)
        for beta in (2., 3., 1.5, 4.):
            out[f'{j}:subRad_fc:{beta}'] = FC.subRadAnalytic_fc_LL(C, beta, jet)
            out[f'{j}:subRadPrime_fc:{beta}'] = FC.subRadPrimeAnalytic_fc_LL(C, beta, jet)
            out[f'{j}:subPDF_fc:{beta}'] = FC.subPDFAnalytic_fc_LL(C, beta, jet)
            out[f'{j}:subPDF_fc_R:{beta}'] = FC.subPDFAnalytic_fc_LL(C, beta, jet, maxRadius=maxr)
            out[f'{j}:subRad_rc:{beta}'] = RC.subRadAnalytic(C, beta, jet)
            out[f'{j}:subRadPrime_rc:{beta}'] = RC.subRadPrimeAnalytic(C, beta, jet)
            out[f'{j}:subPDF_rc:{beta}'] = RC.subPDFAnalytic(C, beta, jet)
            out[f'{j}:subPDF_rc_R:{beta}'] = RC.subPDFAnalytic(C, beta, jet, maxRadius=maxr)
            out[f'{j}:subRad_rc_R:{beta}'] = RC.subRadAnalytic(C, beta, jet, maxRadius=maxr)
    for beta in (2., 3., 1.5):
        for acc in ('LL', 'MLL'):
            out[f'C_ungroomed:{beta}:{acc}'] = OBS.C_ungroomed(z, th, beta, acc)
            zz = np.maximum(z, .1)
            out[f'C_groomed:{beta}:{acc}'] = OBS.C_groomed(zz, th, .1, beta, acc=acc)
            out[f'C_groomed_pre_f1:{beta}:{acc}'] = OBS.C_groomed(zz, th, .1, beta, z_pre=np.minimum(zpre, .1), f=1., acc=acc)
            out[f'C_groomed_pre_f.3:{beta}:{acc}'] = OBS.C_groomed(zz, th, .1, beta, z_pre=np.minimum(zpre, .1), f=.3, acc=acc)
            out[f'C_groomed_mmdt:{beta}:{acc}'] = OBS.C_groomed(zz, th, .1, beta, z_pre=.1, f=0., acc=acc)
    save("functions", **out)


# ------------------------------------------------------ config replays ----
def run_sampler(s, n, case):
    random.seed(SEED + case)
    s.generateSamples(n)
    return np.asarray(s.samples), np.asarray(s.jacobians, dtype=float), float(s.area)


def config_c1():
    """C1: ungroomed quark LL fixed coupling, log sampling, e_2^(2)."""
    n = 2500
    for case, (method, eps) in enumerate([('log', 1e-10), ('lin', 0.), ('log', 1e-3)]):
        s = ungroomedSampler(method, epsilon=eps)
        samples, jac, area = run_sampler(s, n, case)
        z, th = samples[:, 0], samples[:, 1]
        out = dict(seed=SEED + case, samples=samples, jac=jac, area=area, epsilon=eps)
        for jet in ('quark', 'gluon'):
            for fc, acc in ((True, 'LL'), (False, 'MLL')):
                obs = OBS.C_ungroomed(z, th, 2., acc)
                w = W.radiatorWeight(z, th, jet, fixedcoupling=fc, acc=acc)
                it = integrator()
                it.setLastBinBndCondition([0., 'plus'])
                it.setBins(100, obs, 'log' if method == 'log' else 'lin')
                it.setDensity(obs, w * jac, area)
                it.integrate()
                tag = f"{jet}_{'fc' if fc else 'rc'}_{acc}:"
                out[tag + 'obs'] = obs
                out[tag + 'weights'] = w
                out[tag + 'midpoints'] = it.bin_midpoints
                for k, v in integ_arrays(it).items():
                    out[tag + k] = v
        save(f"c1_{method}{case}", **out)


def config_crit():
    """Critical emission Sudakov (test_fixedcoupcritsudakov / running variant)."""
    n = 2500
    case = 10
    for method, eps in (('log', 1e-10), ('lin', 0.)):
        for zc in (.1, .05):
            s = criticalSampler(method, zc=zc, epsilon=eps)
            samples, jac, area = run_sampler(s, n, case)
            z, th = samples[:, 0], samples[:, 1]
            out = dict(seed=SEED + case, samples=samples, jac=jac, area=area,
                       z_cut=zc, epsilon=eps)
            for jet in ('quark', 'gluon'):
                for fc in (True, False):
                    for beta in (2., 3.):
                        obs = OBS.C_groomed(z, th, zc, beta)
                        w = W.criticalEmissionWeight(z, th, zc, jet, fixedcoupling=fc)
                        wj = w * jac
                        if not fc:
                            wj = np.nan_to_num(wj)
                        it = integrator()
                        it.setLastBinBndCondition([1., 'minus'])
                        it.setBins(100, obs, method)
                        it.setDensity(obs, wj, area)
                        it.integrate()
                        tag = f"{jet}_{'fc' if fc else 'rc'}_b{beta:g}:"
                        out[tag + 'obs'] = obs
                        out[tag + 'weights'] = w
                        for k, v in integ_arrays(it).items():
                            out[tag + k] = v
            save(f"crit_{method}_{zc}", **out)
            case += 1


def config_c3_c4():
    """C3: Soft Drop (mMDT form) crit+sub, verbatim twoEmissionWeight; C4 adds
    the pre-critical emission with the zero-bin logic."""
    n = 4000
    zc, beta, eps = .1, 2., 1e-10
    case = 30
    random.seed(SEED + case)
    cs = criticalSampler('log', zc=zc, epsilon=eps)
    cs.generateSamples(n)
    ss = ungroomedSampler('log', epsilon=eps)
    ss.generateSamples(n)
    ps = precriticalSampler('log', zc=zc, epsilon=eps)
    ps.generateSamples(n)
    zero_bin_u = np.array([random.random() for _ in range(n)])
    crit, sub, zpre = np.asarray(cs.samples), np.asarray(ss.samples), np.asarray(ps.samples)
    jc, js, jp = (np.asarray(x.jacobians, dtype=float) for x in (cs, ss, ps))
    edges = np.logspace(np.log10(eps) - 1, np.log10(.5), 100)
    base = dict(seed=SEED + case, crit=crit, sub=sub, z_pre=zpre, jac_crit=jc,
                jac_sub=js, jac_pre=jp, area_crit=cs.area, area_sub=ss.area,
                area_pre=ps.area, zero_bin_u=zero_bin_u, edges_fixed=edges)
    out3 = dict(base)
    for jet in ('quark', 'gluon'):
        for fc in (True, False):
            tag = f"{jet}_{'fc' if fc else 'rc'}"
            w = W.twoEmissionWeight(crit, sub, zc, beta, jet, fixedcoupling=fc)
            csub = sub[:, 0] * sub[:, 1] ** beta * sub[:, 1] ** beta
            for acc in ('LL', 'MLL'):
                ccrit = OBS.C_groomed(crit[:, 0], crit[:, 1], zc, beta,
                                      z_pre=zc, f=0., acc=acc)
                obs = np.maximum(ccrit, csub)
                wj = np.nan_to_num(w) * jc * js
                it = integrator()
                it.setLastBinBndCondition([1., 'minus'])
                it.bins = edges
                it.hasBins = True
                it.setDensity(obs, wj, cs.area * ss.area)
                it.integrate()
                t = f"{tag}_{acc}:"
                out3[t + 'obs'] = obs
                out3[t + 'weights'] = w
                out3['csub'] = csub
                out3[f'c_crit:{acc}'] = ccrit
                for k, v in integ_arrays(it).items():
                    out3[t + k] = v
    save("c3", **out3)
    # C4 (fixed coupling only in the reference), all-emission integrand of
    # examples/tests/comparison_tests/comparison_utils_probabilistic.py:384-474
    for jet in ('quark', 'gluon'):
        CRj = Q.CR(jet)
        th_c = crit[:, 1]
        zp = zpre.copy()
        cdf_eps = np.exp(-2. * CRj * Q.alpha_fixed / np.pi * np.log(Q.R0 / th_c)
                         * np.log(zc / eps))
        zp[zero_bin_u < cdf_eps] = 0.
        c_sub = sub[:, 0] * th_c ** beta
        obs = np.maximum(OBS.C_groomed(crit[:, 0], th_c, zc, beta, z_pre=zp, f=1., acc='LL'), c_sub)
        zero_w = np.exp(-2. * CRj * Q.alpha_fixed / np.pi * np.log(Q.R0 / th_c)
                        * np.log(zc / eps)) * (zp == 0)
        pw = np.nan_to_num(W.precriticalEmissionWeight(zp, th_c, zc, jet, True) * (zp > 0))
        w = (W.criticalEmissionWeight(crit[:, 0], th_c, zc, jet, fixedcoupling=True)
             * FC.subPDFAnalytic_fc_LL(c_sub / th_c ** beta, beta, jet_type=jet)
             * (zero_w + pw))
        jacs = np.where(zp == 0, jc * js, jc * jp * js)
        area = cs.area * ps.area * ss.area
        it = integrator()
        it.setLastBinBndCondition([1., 'minus'])
        it.setBins(100, obs, 'log')
        it.setDensity(obs, w * jacs, area)
        it.integrate()
        save(f"c4_{jet}_fc", seed=SEED + case, z_pre_zb=zp, obs=obs, weights=w, jacs=jacs,
             area=area, **integ_arrays(it))


def config_c5():
    """C5: sub radiator on a theta_crit grid (gen_crit_sub_num_rad loop body)."""
    n = 3000
    eps = 1e-10
    case = 50
    s = ungroomedSampler('log', epsilon=eps)
    samples, jac, area = run_sampler(s, n, case)
    grid = lin_log_mixed_list(eps, 1., 12)
    for beta in (2., 3.):
        for jet, fc, oacc, sacc in (('quark', True, 'LL', 'LL'), ('gluon', False, 'MLL', 'MLL')):
            E, D, DE, I, IE = [], [], [], [], []
            for thc in grid:
                th_em = samples[:, 1] * thc
                obs = OBS.C_ungroomed(samples[:, 0], th_em, beta, acc=oacc)
                it = integrator()
                it.setLastBinBndCondition([0., 'plus'])
                it.setBins(40, obs, 'log', min_log_bin=thc ** beta * 1e-15)
                w = W.radiatorWeight(samples[:, 0], th_em, jet, fixedcoupling=fc, acc=sacc)
                it.setDensity(obs, w * (jac * thc), area)
                it.integrate()
                a = integ_arrays(it)
                E.append(a['edges']); D.append(a['density']); DE.append(a['density_err'])
                I.append(a['integral']); IE.append(a['integral_err'])
            save(f"c5_{jet}_{'fc' if fc else 'rc'}_b{beta:g}", seed=SEED + case,
                 samples=samples, jac=jac, area=area, grid=grid, epsilon=eps,
                 edges=np.array(E), density=np.array(D), density_err=np.array(DE),
                 integral=np.array(I), integral_err=np.array(IE))


def simple_integrals():
    """simpleSampler + integrate_1d known answers (test_simpleIntegrator.py)."""
    case = 60
    for method, eps in (('lin', 0.), ('log', 1e-8)):
        s = simpleSampler(method, bounds=[0, 1], epsilon=eps)
        samples, jac, area = run_sampler(s, 4000, case)
        it = integrator()
        it.setFirstBinBndCondition(0.)
        it.setBins(50, samples, method)
        w = 3. * samples ** 2.
        it.setDensity(samples, w * jac, area)
        it.integrate()
        save(f"simple_{method}", seed=SEED + case, samples=samples, jac=jac, area=area,
             weights=w, epsilon=eps, **integ_arrays(it))
        case += 1
    random.seed(SEED + case)
    val, err, x = integrate_1d(lambda x: 2. * x, [.1, .5], bin_space='lin', num_samples=2000)
    save("integrate_1d_lin", seed=SEED + case, value=val, error=err, x=x)
    random.seed(SEED + case + 1)
    integ, err, xs = integrate_1d(lambda x: 1. / x, [.1, .5], bin_space='log', epsilon=1e-6,
                                  num_samples=2000, num_bins=20, bnd_cond=0, bnd_cond_bin='first')
    save("integrate_1d_log", seed=SEED + case + 1, integral=np.asarray(integ, float),
         error=np.asarray(err), xs=np.asarray(xs))


def shower():
    """Veto-algorithm shower chains (z, theta) and their ECFs."""
    case = 70
    for jet in ('quark', 'gluon'):
        for acc in ('LL', 'MLL'):
            random.seed(SEED + case)
            njets = 400
            jets = PS.gen_jets(njets, beta=2., jet_type=jet, radius=1., acc=acc,
                               cutoff=Q.MU_NP, verbose=0)
            n_uniform = None
            chains_z, chains_t, lens = [], [], []
            for jt in jets:
                zs, ts = [], []
                for m in jt.partons:
                    if m.daughter1 is None or m.daughter2 is None:
                        continue
                    d1, d2 = m.daughter1, m.daughter2
                    ts.append(PS.angle(d1.momentum, d2.momentum))
                    zs.append(min(d1.momentum.mag() / m.momentum.mag(),
                                  d2.momentum.mag() / m.momentum.mag()))
                chains_z += zs
                chains_t += ts
                lens.append(len(zs))
            out = dict(seed=SEED + case, lens=np.array(lens), z=np.array(chains_z),
                       theta=np.array(chains_t))
            for oacc in ('LL', 'MLL'):
                for nem in (1, 2, 'all'):
                    out[f'ungroomed:{oacc}:{nem}'] = np.array(
                        PS.getECFs_ungroomed(jets, beta=2., obs_acc=oacc, n_emissions=nem))
                for nem in (1, 'all'):
                    out[f'softdrop:{oacc}:{nem}'] = np.array(
                        PS.getECFs_softdrop(jets, .1, beta=2., beta_sd=0., obs_acc=oacc,
                                            n_emissions=nem))
                # n_emissions=2 evaluated one jet at a time so the in-place
                # decrement at partonshower_utils.py:765 cannot leak across jets
                out[f'softdrop:{oacc}:2'] = np.array(
                    [PS.getECFs_softdrop([jt] * 5, .1, beta=2., beta_sd=0., obs_acc=oacc,
                                         n_emissions=2)[0] for jt in jets])
                out[f'softdrop_bsd1:{oacc}:1'] = np.array(
                    PS.getECFs_softdrop(jets, .1, beta=2., beta_sd=1., obs_acc=oacc,
                                        n_emissions=1))
            save(f"shower_{jet}_{acc}", **out)
            case += 1


def misc():
    save("lin_log_mixed", a=lin_log_mixed_list(1e-10, 1., 12), b=lin_log_mixed_list(1e-15, 1., 501))
    rng = np.random.RandomState(7)
    vals = np.exp(rng.uniform(np.log(1e-14), 0, 3000))
    wts = rng.uniform(0, 2, 3000)
    xs, pdf = vals_to_pdf(vals, 50, weights=wts, log_cutoff=1e-10)
    save("vals_to_pdf", vals=vals, weights=wts, xs=xs, pdf=pdf)


if __name__ == "__main__" and len(sys.argv) == 1:
    os.makedirs(OUT, exist_ok=True)
    kats()
    function_vectors()
    config_c1()
    config_crit()
    config_c3_c4()
    config_c5()
    simple_integrals()
    shower()
    misc()


# ------------------------------------------------------------------ round 1b: 2-D integrator and grid callers ----
def callers():
    """integrator_2d and the numerical-radiator generators (the batched callers of the hot path).
    Run alone with ``python oracle/make_golden.py callers`` (keeps the other fixtures untouched)."""
    from jetmontecarlo.montecarlo.integrator import integrator_2d
    from jetmontecarlo.numerics.radiators import generation as GEN
    n, zc, eps = 3000, .1, 1e-8
    for method in ('lin', 'log'):
        random.seed(SEED + 80)
        e = eps if method == 'log' else 0.
        cs = criticalSampler(method, zc=zc, epsilon=e)
        cs.generateSamples(n)
        ps = precriticalSampler(method, zc=zc, epsilon=e)
        ps.generateSamples(n)
        us = ungroomedSampler(method, epsilon=e)
        us.generateSamples(n)
        out = dict(seed=SEED + 80, crit=np.asarray(cs.samples), pre=np.asarray(ps.samples), sub=np.asarray(us.samples),
                   jac_crit=np.asarray(cs.jacobians, float), jac_pre=np.asarray(ps.jacobians, float),
                   jac_sub=np.asarray(us.jacobians, float), area_crit=cs.area, area_pre=ps.area, area_sub=us.area,
                   epsilon=e, z_cut=zc)
        # plain integrator_2d
        x, y = np.asarray(ps.samples), np.asarray(cs.samples)[:, 1]
        w = W.radiatorWeight(x, y, 'quark', fixedcoupling=True, acc='LL') * (x * y if method == 'log' else 1.)
        for bc_name, setter in (('last', lambda it: it.setLastBinBndCondition([0., 'plus'])),
                                ('lastminus', lambda it: it.setLastBinBndCondition([1., 'minus'])),
                                ('first', lambda it: it.setFirstBinBndCondition(0.))):
            it = integrator_2d()
            setter(it)
            it.setBins(12, [x, y], method)
            it.setDensity([x, y], w, 2.5)
            it.integrate()
            t = f'i2d_{bc_name}:'
            out[t + 'bins0'], out[t + 'bins1'] = it.bins[0], it.bins[1]
            out[t + 'density'], out[t + 'density_err'] = it.density, it.densityErr
            out[t + 'integral'], out[t + 'integral_err'] = it.integral, it.integralErr
            out[t + 'ext'], out[t + 'ext_err'] = it.extended_integral, it.extended_integralErr
        out['i2d_weights'] = w
        # gen_numerical_radiator
        for jet, fc, acc in (('quark', True, 'LL'), ('gluon', False, 'MLL')):
            tag = f"{jet}_{'fc' if fc else 'rc'}:"
            _, d = GEN.gen_numerical_radiator(cs, 'crit', jet_type=jet, obs_accuracy=acc, splitfn_accuracy=acc,
                                              bin_space=method, fixed_coupling=fc, num_bins=30)
            for k, v in d.items():
                out[tag + 'crit:' + k] = np.asarray(v, float)
            _, d = GEN.gen_numerical_radiator(us, 'sub', jet_type=jet, obs_accuracy=acc, splitfn_accuracy=acc,
                                              beta=2., bin_space=method, fixed_coupling=fc, num_bins=30)
            for k, v in d.items():
                out[tag + 'sub:' + k] = np.asarray(v, float)
            _, d = GEN.gen_pre_num_rad(ps, cs, jet_type=jet, splitfn_accuracy=acc, bin_space=method,
                                       fixed_coupling=fc, num_bins=10)
            for k, v in d.items():
                out[tag + 'pre:' + k] = np.asarray(v, float)
            _, d = GEN.gen_crit_sub_num_rad(us, jet_type=jet, obs_accuracy=acc, splitfn_accuracy=acc, beta=2.,
                                            epsilon=eps, bin_space=method, fixed_coupling=fc, num_bins=10)
            for k, v in d.items():
                out[tag + 'critsub:' + k] = np.asarray(v, float)
        save(f"callers_{method}", **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "callers":
    callers()


# ------------------------------------------------------------------ round 1c: RSS grooming of shower chains ----
RSS_CASES = [   # (f, obs_acc, emission_type, n_emissions, few_pres)
    (1., 'LL', 'precritsub', 1, True), (1., 'LL', 'precritsub', 2, True), (1., 'MLL', 'precritsub', 'all', True),
    (1., 'MLL', 'crit', 1, True), (1., 'LL', 'critsub', 2, True), (1., 'LL', 'sub', 1, True),
    (1., 'MLL', 'precrit', 1, True), (.75, 'LL', 'precritsub', 1, True), (.75, 'MLL', 'precritsub', 'all', True),
    (.75, 'LL', 'precritsub', 2, False), (.5, 'MLL', 'precritsub', 1, False), (1., 'LL', 'precritsub', 'all', False),
]


def shower_rss():
    """getECFs_rss on the same jets as shower() (same seeds, so the chains must coincide with shower_*.npz)."""
    case = 70
    for jet in ('quark', 'gluon'):
        for acc in ('LL', 'MLL'):
            random.seed(SEED + case)
            jets = PS.gen_jets(400, beta=2., jet_type=jet, radius=1., acc=acc, cutoff=Q.MU_NP, verbose=0)
            lens = []
            for jt in jets:
                lens.append(sum(1 for m in jt.partons if m.daughter1 is not None and m.daughter2 is not None))
            out = dict(seed=SEED + case, lens=np.array(lens))
            for z_cut in (.1, .2):
                for (f, oacc, etype, nem, few) in RSS_CASES:
                    vals = PS.getECFs_rss(jets, z_cut, beta=2., f=f, obs_acc=oacc, emission_type=etype,
                                          n_emissions=nem, few_pres=few)
                    assert len(vals) == len(jets)        # no jet is dropped for these settings
                    out[f'rss:{z_cut}:{f}:{oacc}:{etype}:{nem}:{int(few)}'] = np.array(vals)
            out['rss_zcut0:LL:2'] = np.array(PS.getECFs_rss(jets, 0, beta=2., obs_acc='LL', n_emissions=2))
            save(f"shower_rss_{jet}_{acc}", **out)
            case += 1


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "rss":
    shower_rss()


def shower_correlations():
    """The npz written by save_shower_correlations (the shower production script's output format,
    partonshower_utils.py:802-884) for the quark MLL jets of shower()."""
    random.seed(SEED + 71)
    jets = PS.gen_jets(400, beta=2., jet_type='quark', radius=1., acc='MLL', cutoff=Q.MU_NP, verbose=0)
    path = os.path.join(OUT, "shower_corr_quark_MLL.npz")
    PS.save_shower_correlations(jets, path, z_cuts=[.05, .1, .2], beta=2., f_soft=1., obs_acc='MLL', few_pres=True,
                                verbose=0)
    z = dict(np.load(path))
    np.savez_compressed(path, **z)
    print("wrote", path, os.path.getsize(path), "bytes", {k: v.shape for k, v in z.items()})


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "corr":
    shower_correlations()


# ------------------------------------------------------------------ round 1d: samples_from_cdf without pynverse ----
def _refuse(*a, **k):
    raise ValueError("inversefunc: refused (pynverse is not installed here; this drives samples_from_cdf into "
                     "its own tabulated fall-backs)")


CDF_CASES = {
    # name: (cdf, domain, kwargs)
    'square': (lambda x: np.clip(x, 0, 1) ** 2, [0, 1], {}),
    'sudakov_pos_domain': (lambda x: np.exp(-np.log(1. / x) ** 2 * .3), [1e-8, 1.], {}),
    'neg_domain': (lambda x: .5 * (1 + np.tanh(x)), [-4., 4.], {}),
    'starts_at_one': (lambda x: np.where(x == 0, 1., .5 + .5 * x), [0, 1], {}),
    'negligible_then_monotone': (lambda x: np.where(x < .01, 1e-30 * (1 + np.sin(1e4 * x)), (x - .01) / .99),
                                 [0, 1], {}),
    'turnpoint': (lambda x: np.where(x < .6, np.minimum(x / .4, 1.), 1. - (x - .6)), [1e-6, 1.],
                  dict(catch_turnpoint=True)),
    'wiggle_forced': (lambda x: np.clip(x + .02 * np.sin(60 * x), 0, 1), [0, 1], dict(force_monotone=True)),
    'nan_head_forced': (lambda x: np.where(x < 1e-4, np.nan, np.clip(x + .05 * np.sin(25 * x), 0, 1)), [0, 1],
                        dict(force_monotone=True)),
}


def cdf_sampling():
    """samples_from_cdf with inversefunc refusing every cdf: the reference's own tabulated branches."""
    import jetmontecarlo.utils.montecarlo_utils as MU
    MU.inversefunc = _refuse
    out = {}
    for k, (name, (cdf, domain, kw)) in enumerate(CDF_CASES.items()):
        np.random.seed(SEED + 300 + k)
        samples, weights = MU.samples_from_cdf(cdf, 1500, domain=domain, verbose=0, **kw)
        out[name + ':samples'], out[name + ':weights'] = np.asarray(samples, float), np.asarray(weights, float)
        out[name + ':seed'] = SEED + 300 + k
    for name in ('wiggle_forced', 'turnpoint'):        # without its flag the reference gives up
        cdf, domain, _ = CDF_CASES[name]
        try:
            MU.samples_from_cdf(cdf, 10, domain=domain, verbose=0)
            raised = False
        except ValueError:
            raised = True
        out[name + ':raises_without_flag'] = raised
    save("samples_from_cdf", **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "cdf":
    cdf_sampling()


# ------------------------------------------------------------------ round 1e: interpolation helpers ----
def interpolation_helpers():
    """get_1d_interpolation / monotonic_domain / get_2d_interpolation as the reference evaluates them
    (utils/interpolation.py:43-273), on tables taken from the golden integrals and on a few analytic functions."""
    from jetmontecarlo.utils import interpolation as I
    out = {}
    g = np.load(os.path.join(OUT, "c1_log0.npz"))
    tables = {
        'rad_log': (g['quark_fc_LL:edges'], g['quark_fc_LL:integral']),                  # decreasing, log-spaced
        'lin_inc': (np.linspace(0., 2., 21), np.linspace(0., 2., 21) ** 2),               # increasing, lin-spaced
        'wiggle': (np.linspace(0., 1., 41), np.linspace(0., 1., 41) + .05 * np.sin(20 * np.linspace(0., 1., 41))),
    }
    for name, (xs, fs) in tables.items():
        probe = np.concatenate([[xs[0] - 1., xs[0], xs[1], xs[-1], xs[-1] + 1.], np.sqrt(np.abs(xs[1:] * xs[:-1])),
                                (xs[1:] + xs[:-1]) / 2., xs])
        out[f'{name}:xs'], out[f'{name}:fs'], out[f'{name}:probe'] = xs, fs, probe
        if name != 'wiggle':
            # (a fresh [None, None] every time: the reference's default is one mutable list that keeps the boundary
            # values of the first table it ever saw, interpolation.py:188,229-232)
            out[f'{name}:mono_default'] = I.get_1d_interpolation(xs, fs, monotonic=True,
                                                                 bound_values=[None, None])(probe)
            out[f'{name}:mono_bounds'] = I.get_1d_interpolation(xs, fs, monotonic=True, bounds=(xs[2], xs[-3]),
                                                               bound_values=[None, None])(probe)
        out[f'{name}:free_bounds'] = I.get_1d_interpolation(xs, fs, monotonic=False, bounds=(xs[0], xs[-1]),
                                                           bound_values=[0., 0.])(probe)
        out[f'{name}:free_cont'] = I.get_1d_interpolation(xs, fs, monotonic=False, bounds=(xs[1], xs[-2]),
                                                         bound_values=None)(probe)
        out[f'{name}:is_mono'] = np.array([I.is_monotonic_arr(fs), I.is_monotonic_arr(fs, 'increasing'),
                                           I.is_monotonic_arr(fs, 'decreasing')])
        out[f'{name}:where_inc'] = I.where_monotonic_arr(fs, 'increasing')
        out[f'{name}:where_dec'] = I.where_monotonic_arr(fs, 'decreasing')
    funcs = {'square': lambda x: x ** 2, 'sin': lambda x: np.sin(3. * x), 'cubic': lambda x: x ** 3 - x}
    md = []
    for fname, domain, point, check in [('square', [-10, 10], 1, 'increasing'), ('square', [0, 10], 1, 'increasing'),
                                        ('square', [-10, 10], 1, 'decreasing'), ('square', [-10, 10], -1, 'decreasing'),
                                        ('square', [-10, 10], -1, 'increasing'), ('sin', [0.1, 3.], 1.5, None),
                                        ('sin', [-2., 2.], 0.1, 'increasing'), ('cubic', [-2., 2.], 0., None),
                                        ('cubic', [.1, 2.], None, None), ('square', [1e-3, 4.], None, 'increasing')]:
        res = I.monotonic_domain(funcs[fname], domain, point, check)
        md.append([np.nan, np.nan] if res is None else [float(res[0]), float(res[1])])
    out['monotonic_domain'] = np.array(md)
    out['is_monotonic_func'] = np.array([I.is_monotonic_func(funcs['square'], [.1, 3.]),
                                         I.is_monotonic_func(funcs['sin'], [.1, 3.]),
                                         I.is_monotonic_func(funcs['sin'], [.1, .4], 'increasing'),
                                         I.is_monotonic_func(funcs['square'], [.1, 3.], 'decreasing')])
    # 2-D: rectangular grid (linear, extrapolating) and nearest neighbour
    x, y = np.linspace(0., 1., 6), np.logspace(-3, 0, 5)
    z = np.outer(x ** 2, np.log(1. / y) + 1.)
    pts = np.array([[.1, .01], [.5, .5], [0., 1e-3], [1., 1.], [1.2, .5], [.33, 2.], [-.1, 1e-4]])
    out['grid:x'], out['grid:y'], out['grid:z'], out['grid:pts'] = x, y, z, pts
    out['grid:linear'] = I.get_2d_interpolation(x, y, z)(pts)
    out['grid:nearest_method'] = I.get_2d_interpolation(x, y, z, method='nearest')(pts)
    xx, yy = np.meshgrid(x, y, indexing='ij')
    out['scatter:nearest'] = I.get_2d_interpolation(xx.flatten(), yy.flatten(), z.flatten(),
                                                    interpolation_method='Nearest')(pts[:, 0], pts[:, 1])
    save("interpolation", **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "interp":
    interpolation_helpers()


# ------------------------------------------------------------------ round 1f: histogram post-processing ----
def hist_postprocessing():
    """histDerivative / smooth_data (utils/hist_utils.py:7-91, 158-173) on a golden cumulative integral."""
    from jetmontecarlo.utils.hist_utils import histDerivative, smooth_data
    out = {}
    for name in ('c1_log0', 'c1_lin1'):
        g = np.load(os.path.join(OUT, name + ".npz"))
        edges, integral = g['quark_fc_LL:edges'], g['quark_fc_LL:integral']
        spacing = 'log' if 'log' in name else 'lin'
        hist = integral[:-1]
        interp, deriv = histDerivative(hist, edges, giveHist=True, binInput=spacing)
        mids = np.exp((np.log(edges[1:]) + np.log(edges[:-1])) / 2.) if spacing == 'log' else (edges[1:] + edges[:-1]) / 2.
        probe = np.concatenate([mids, edges[1:-1], [edges[0] * .5 if spacing == 'log' else -1., edges[-1] * 2.]])
        out[f'{name}:deriv'], out[f'{name}:probe'], out[f'{name}:interp'] = deriv, probe, interp(probe)
        out[f'{name}:smooth'] = smooth_data(mids, hist)
    save("hist_postprocessing", **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "histpost":
    hist_postprocessing()


# ------------------------------------------------------------------ round 1g: closed-form truth curves ----
def closed_forms():
    """critSudakov_fc_LL (analytics/sudakov_factors/fixedcoupling.py:11-51) and the Soft Drop closed forms
    (analytics/radiators/soft_drop.py:15-194) on observable grids -- the curves the reference's tests plot their
    Monte Carlo results against."""
    import types
    for name in ("jetmontecarlo.utils.plot_utils", "jetmontecarlo.utils.color_utils"):
        if name not in sys.modules:                      # plotting helpers (matplotlib): not needed for the formulas
            sys.modules[name] = types.ModuleType(name)
    from jetmontecarlo.analytics.radiators import soft_drop as SD
    out = {}
    cs = np.logspace(-9, np.log10(.6), 60)
    out['c'] = cs
    for jet in ('quark', 'gluon'):
        for z_c in (.05, .1, .2):
            for beta in (1.5, 2., 3.):
                for f in (1., .75):
                    out[f'critSudakov:{jet}:{z_c}:{beta}:{f}'] = np.array(
                        critSudakov_fc_LL(cs, z_c, beta, jet, f=f), dtype=float)
                out[f'sd_rad_fc:{jet}:{z_c}:{beta}'] = SD.softdrop_rad_fc(cs, z_c, beta, jet_type=jet)
                out[f'sd_radprime_fc:{jet}:{z_c}:{beta}'] = SD.softdrop_radprime_fc(cs, z_c, beta, jet_type=jet)
                out[f'sd_sudakov_fc_LL:{jet}:{z_c}:{beta}'] = SD.softdrop_sudakov_fc(cs, z_c, beta, jet_type=jet)
                out[f'sd_sudakov_fc_ME:{jet}:{z_c}:{beta}'] = SD.softdrop_sudakov_fc(cs, z_c, beta, jet_type=jet,
                                                                                     acc='MLL')
                for beta_sd in (0, 1.):
                    out[f'sd_rad_rc:{jet}:{z_c}:{beta}:{beta_sd}'] = SD.softdrop_rad_rc(cs, z_c, beta, beta_sd=beta_sd,
                                                                                        jet_type=jet)
    save("closed_forms", **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "closed":
    closed_forms()


def closed_forms_fc():
    """The fixed-coupling radiators beyond leading log (analytics/radiators/fixedcoupling.py:20-57, 110-160) and the
    pre-critical radiator without frozen coupling (running_coupling.py:291-312)."""
    out = {}
    thetas = np.concatenate([[0.], np.logspace(-8, 0, 40)])
    cs = np.concatenate([[0.], np.logspace(-9, np.log10(.499), 40)])
    zps = np.logspace(-6, np.log10(.099), 30)
    out['theta'], out['c'], out['z_pre'] = thetas, cs, zps
    for jet in ('quark', 'gluon'):
        for z_c in (.05, .1, .2):
            out[f'critRadPrime_fc_LL:{jet}:{z_c}'] = FC.critRadPrimeAnalytic_fc_LL(thetas, z_c, jet)
            out[f'critRad_fc:{jet}:{z_c}'] = FC.critRadAnalytic_fc(thetas, z_c, jet)
            out[f'critRadPrime_fc:{jet}:{z_c}'] = FC.critRadPrimeAnalytic_fc(thetas, z_c, jet)
            for th in (.3, 1e-2):
                out[f'preRad_nofreeze:{jet}:{z_c}:{th}'] = RC.preRadAnalytic_nofreeze(zps * z_c / .1, th, z_c, jet_type=jet)
        for beta in (1.5, 2., 3.):
            out[f'subRad_fc:{jet}:{beta}'] = np.array(FC.subRadAnalytic_fc(cs, beta, jet), dtype=float)
            out[f'subRadPrime_fc:{jet}:{beta}'] = FC.subRadPrimeAnalytic_fc(cs, beta, jet)
    save("closed_forms_fc", **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "closedfc":
    closed_forms_fc()


def process_ewoc():
    """The generic MonteCarloProcess (hit-or-miss box sampling with an efficiency-scaled area, montecarlo/process.py:
    109-191) through its one client, e+e- -> hadrons at LO (examples/processes/ee2hadrons_LO.py), the pairwise
    observables and the energy-weighted correlator built on it (montecarlo/ewoc.py:11-75;
    examples/ewocs/ee2hadrons_LO/eec.py, mass-squared_ewoc.py)."""
    import types
    if "matplotlib" not in sys.modules:                  # ewoc.py imports pyplot for plot_ewoc only
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    from jetmontecarlo.montecarlo.ewoc import Npoint_MonteCarloEWOC
    from jetmontecarlo.analytics.ewocs.eec import ee2hadrons_LO_prefactor
    from examples.processes import ee2hadrons_LO as E
    for case, method in enumerate(('log', 'lin')):
        out = {}
        random.seed(SEED + 900 + case)
        # the uniforms the sampling loop will consume (2 per trial), for the oracle's replay
        state = random.getstate()
        n = 3000
        process = E.ee2hadrons_LO(energy=1000., num_samples=n, sampleMethod=method)
        n_trials = n + process.num_rejected_samples
        random.setstate(state)
        out['uniforms'] = np.array([random.random() for _ in range(2 * n_trials)])
        out['samples'] = np.array(process.samples)
        out['jacobians'] = np.array(process.jacobians)
        out['num_rejected'] = np.array(process.num_rejected_samples)
        out['area'] = np.array(process.area)
        out['prefactor'] = np.array(ee2hadrons_LO_prefactor(1000.))
        named = process.named_samples()
        out['x_1'], out['x_2'], out['x_3'] = (np.array(named[k]) for k in ('x_1', 'x_2', 'x_3'))
        for name, fn in (('costheta', E.costheta_ij), ('m2', E.m2_ij), ('theta', E.theta_ij), ('z_i', E.z_i), ('z_j', E.z_j)):
            ob = E.ee2hLO_PairwiseObservable(process, fn, name)
            out[f'obs:{name}'] = np.array(ob.observables)
            out[f'obs:{name}:weights'] = np.array(ob.weights)
        out['obs:gecf1.5'] = np.array(E.ee2hLO_PairwiseObservable(process, lambda a, b: E.gecf_ij(a, b, 1.5), 'g').observables)
        zi = E.ee2hLO_PairwiseObservable(process, E.z_i, 'z_i')
        zj = E.ee2hLO_PairwiseObservable(process, E.z_j, 'z_j')
        for name, fn, spacing, ew in (('costheta', E.costheta_ij, 'lin', (1, 1)), ('m2', E.m2_ij, 'lin', (1, 1)),
                                      ('m2', E.m2_ij, 'log', (2, 1)), ('theta', E.theta_ij, 'log', (1, 1))):
            ob = E.ee2hLO_PairwiseObservable(process, fn, name)
            ewoc = Npoint_MonteCarloEWOC(process, ob, energy_weights=ew)
            ewoc.set_weights(zi, zj)
            ewoc.generate_ewoc(numbins=40, binspacing=spacing)
            tag = f'ewoc:{name}:{spacing}:{ew[0]}{ew[1]}'
            out[tag + ':weights'] = np.array(ewoc.weights)
            out[tag + ':bins'] = np.array(ewoc.bins)
            out[tag + ':density'] = np.array(ewoc.density)
            out[tag + ':density_err'] = np.array(ewoc.densityErr)
        # explicit range end points
        ob = E.ee2hLO_PairwiseObservable(process, E.costheta_ij, 'c')
        ewoc = Npoint_MonteCarloEWOC(process, ob)
        ewoc.set_weights(zi, zj)
        ewoc.generate_ewoc(numbins=25, binspacing='lin', minval=-1., maxval=1.)
        out['ewoc:range:bins'], out['ewoc:range:density'] = np.array(ewoc.bins), np.array(ewoc.density)
        save(f"process_ee2h_{method}", **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "process":
    process_ewoc()


def catalog_files():
    """A file catalog written by the reference's own file_management package (catalog_utils.py:115-185,
    data_management.py:43-69): one 1-D numerical integral and one 2-D one (.npz) plus a pickled sample dictionary.
    The catalog's YAML and the data files are committed as they were written, paths relative to the catalog's root."""
    import os, shutil, tempfile
    sys.path.insert(0, "/root/reference")
    here = os.path.dirname(os.path.abspath(__file__))
    dest = os.path.join(here, "..", "tests", "golden", "catalog")
    work = tempfile.mkdtemp()
    cwd = os.getcwd()
    try:
        os.chdir(work)                                  # the reference's folders are relative to the working directory
        os.makedirs("output/examples/current")
        os.makedirs("output/examples/overwritten")
        open("output/examples/file_catalog.yaml", "w").close()
        from file_management.data_management import save_new_data
        rng = np.random.RandomState(11)
        edges = np.logspace(-6, 0, 21)
        save_new_data({'bins': edges, 'integral': np.sort(rng.rand(21))[::-1].copy()}, 'numerical integral', 'critical radiator',
                      {'z_cut': 0.1, 'jet type': 'quark', 'fixed coupling': True, 'number of samples': 1000}, '.npz')
        xs, ys = np.logspace(-8, -1, 9), np.logspace(-8, 0, 7)
        save_new_data({'bins': np.array([xs[:7], ys]), 'integral': rng.rand(7, 7)}, 'numerical integral', 'pre-critical radiator',
                      {'z_cut': 0.1, 'jet type': 'gluon', 'fixed coupling': False, 'number of samples': 1000}, '.npz')
        save_new_data({'samples': rng.rand(5), 'weights': np.ones(5)}, 'montecarlo samples', 'critical sudakov inverse transform',
                      {'z_cut': 0.05, 'beta': 2, 'epsilon': 1e-10}, '.pkl')
        if os.path.exists(dest):
            shutil.rmtree(dest)
        shutil.copytree("output", os.path.join(dest, "output"))
    finally:
        os.chdir(cwd)
        shutil.rmtree(work)
    print("wrote", dest)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "catalog":
    catalog_files()
