"""Which CUDA kernels run under the mirrored API that are NOT this library's?  Each flow runs under torch.profiler
(CUDA activities); kernels whose name does not contain 'jmc::' are listed per flow (memcpy / memset are not kernels).
Expectation: none on the SURVEY 8 (a) and (f) paths."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import jetmontecarlo_b200 as P
from jetmontecarlo_b200.montecarlo.sampler import set_seed, simpleSampler
from jetmontecarlo_b200.montecarlo.integrator import integrator, integrate_1d
from jetmontecarlo_b200.numerics.radiators.samplers import criticalSampler, precriticalSampler, ungroomedSampler
from jetmontecarlo_b200.numerics.observables import C_groomed, C_ungroomed
from jetmontecarlo_b200.numerics.weights import twoEmissionWeight, radiatorWeight, criticalEmissionWeight
from jetmontecarlo_b200.numerics.radiators.generation import gen_numerical_radiator, gen_pre_num_rad, gen_crit_sub_num_rad
from jetmontecarlo_b200.numerics import sudakov_sampling as S
from jetmontecarlo_b200.utils import partonshower_utils as PS
from jetmontecarlo_b200.utils.hist_utils import vals_to_pdf
from jetmontecarlo_b200.montecarlo.ewoc import fused_ewoc
from jetmontecarlo_b200.processes import ee2hadrons_LO as EE
from jetmontecarlo_b200 import fused
set_seed(1)
flows = {}
def flow(name):
    def reg(fn): flows[name] = fn; return fn
    return reg
st = {}
@flow("a4-a8 samplers.generateSamples")
def _():
    st['c'] = criticalSampler('log', zc=.1, epsilon=1e-10); st['c'].generateSamples(200000)
    st['u'] = ungroomedSampler('log', epsilon=1e-10); st['u'].generateSamples(200000)
    st['p'] = precriticalSampler('log', zc=.1, epsilon=1e-10); st['p'].generateSamples(200000)
@flow("a9-a14 observables + weights on host arrays")
def _():
    c, u = st['c'].getSamples(), st['u'].getSamples()
    st['obs'] = np.maximum(C_groomed(c[:, 0], c[:, 1], .1, 2., z_pre=.1, f=0., acc='MLL'), C_ungroomed(u[:, 0], u[:, 1], 2., 'MLL'))
    st['w'] = np.nan_to_num(twoEmissionWeight(c, u, .1, 2., 'quark', False)) * np.array(st['c'].jacobians) * np.array(st['u'].jacobians)
    radiatorWeight(u[:, 0], u[:, 1], 'quark', fixedcoupling=False, acc='MLL'); criticalEmissionWeight(c[:, 0], c[:, 1], .1, 'quark', fixedcoupling=True)
@flow("a21-a23 integrator.setBins / setDensity / integrate")
def _():
    it = integrator(); it.setLastBinBndCondition([1., 'minus']); it.setBins(100, st['obs'], 'log')
    it.setDensity(st['obs'], st['w'], st['c'].area * st['u'].area); it.integrate()
@flow("fused.integrate (C3 rc, data-dependent bins)")
def _():
    ig = fused.Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False, obs_acc='MLL')
    fused.integrate(ig, 2_000_000, numbins=100, binspacing='log', seed=3, precision='f32', last_bc=[1., 'minus'])
    fused.integrate(ig, 2_000_000, edges=np.logspace(-11, np.log10(.5), 100), seed=3, precision='f64', last_bc=[1., 'minus'])
@flow("a26 radiator generators (1-D, 2-D, theta_crit grid)")
def _():
    _, st['crit_rad'] = gen_numerical_radiator(st['c'], 'crit', 'quark', obs_accuracy='LL', splitfn_accuracy='LL', beta=None, bin_space='log', fixed_coupling=True, num_bins=40)
    _, st['pre_rad'] = gen_pre_num_rad(st['p'], st['c'], 'quark', splitfn_accuracy='LL', bin_space='log', fixed_coupling=True, num_bins=40)
    gen_crit_sub_num_rad(st['u'], 'quark', obs_accuracy='LL', splitfn_accuracy='LL', beta=2., epsilon=1e-10, bin_space='log', fixed_coupling=True, num_bins=20)
@flow("a27-a28 gen_jets + ECF pickers")
def _():
    jets = PS.gen_jets(200000, beta=2., jet_type='quark', acc='MLL', verbose=0, seed=2); st['jets'] = jets
    PS.getECFs_ungroomed(jets, 2., obs_acc='MLL', n_emissions=1, verbose=0)
    PS.getECFs_softdrop(jets, .1, beta=2., beta_sd=0., obs_acc='MLL', n_emissions=2, verbose=0)
    PS.getangs(jets, beta=2., acc='MLL', verbose=0)
@flow("a29 vals_to_pdf")
def _():
    vals_to_pdf(st['obs'], 60, weights=st['w'], log_cutoff=1e-10)
@flow("f1 generate_precritical / subsequent samples")
def _():
    b = np.array(st['pre_rad']['bins']); rad = S.TableRadiator(b[0], b[1], np.array(st['pre_rad']['integral']))
    th = np.random.RandomState(0).uniform(1e-3, .9, 50000)
    S.generate_precritical_samples(rad, th, .1, seed=1); S.generate_subsequent_samples(S.TableRadiator(b[0], b[1], np.array(st['pre_rad']['integral']), "Nearest"), th, 2., seed=2)
@flow("f3 getECFs_rss + event record + replay picker")
def _():
    PS.getECFs_rss(st['jets'], .1, 2., 1., obs_acc='MLL', n_emissions=2, verbose=0)
    lens, chains = st['jets'].chains_device(max_emissions=64)
    PS.ecfs_from_chains(lens, chains, beta=2., obs_acc='MLL', n_emissions=2, groomer='rss', z_cut=.1, f=1.)
@flow("f4 process sampling + pair observables + fused EWOC")
def _():
    p = EE.ee2hadrons_LO(energy=1000., num_samples=100000, sampleMethod='lin', seed=1)
    EE.ee2hLO_PairwiseObservable(p, EE.costheta_ij, 'costheta')
    fused_ewoc(p, 1_000_000, EE.costheta_ij, energy_weights=(1, 1), numbins=50, binspacing='lin', seed=4)
# preflight: can this process see kernels through the profiling interface at all?
try:
    from jetmontecarlo_b200 import _device as _D
    _D.full((1024,), 1.)
    with profile(activities=[ProfilerActivity.CUDA]) as _p:
        _D.full((1024,), 2.); torch.cuda.synchronize()
    _ok = any(e.device_type == torch.autograd.DeviceType.CUDA and 'jmc::' in e.name for e in _p.events())
except Exception as exc:
    _ok = False
    print("profiler error:", repr(exc))
if not _ok:
    print("profiler unavailable: no kernel of a profiled launch was reported (another profiler attached?)")
    sys.exit(0)
bad_total = 0
for name, fn in flows.items():
    try:
        fn()                                   # warm (lazy initialisation outside the profile)
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            fn(); torch.cuda.synchronize()
        names = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and 'memcpy' not in e.name.lower() and 'memset' not in e.name.lower():
                names[e.name] = names.get(e.name, 0) + 1
        mine = sum(v for k, v in names.items() if 'jmc::' in k)
        foreign = {k: v for k, v in names.items() if 'jmc::' not in k}
        bad_total += sum(foreign.values())
        print("%-58s jmc kernels %4d   foreign %d %s" % (name, mine, sum(foreign.values()), {k[:60]: v for k, v in foreign.items()} if foreign else ""))
    except Exception as exc:
        print("%-58s ERROR %r" % (name, exc))
print("total foreign kernel launches:", bad_total)
