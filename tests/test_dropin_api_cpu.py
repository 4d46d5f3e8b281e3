"""Drop-in surface against the reference's OWN scripts (runs where /root/reference exists: the build container).

Every ``examples/**.py`` of the reference is parsed; for each module of ``jetmontecarlo.*`` that this package mirrors,
the names those scripts take from it -- explicitly, or through their star-imports -- must exist in the mirror, and the
callables on the hot path must accept the reference's arguments (same parameter names, order and defaults; the mirror
may only ADD keyword parameters with defaults)."""
import ast
import builtins
import glob
import importlib
import inspect
import os

import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")

MIRRORED = ["montecarlo.sampler", "montecarlo.integrator", "montecarlo.process", "montecarlo.ewoc",
            "numerics.radiators.samplers", "numerics.radiators.generation", "numerics.observables", "numerics.weights",
            "numerics.splitting", "analytics.qcd_utils", "analytics.radiators.fixedcoupling",
            "analytics.radiators.running_coupling", "analytics.sudakov_factors.fixedcoupling", "analytics.ewocs.eec",
            "utils.montecarlo_utils", "utils.hist_utils", "utils.partonshower_utils", "utils.interpolation"]

# names the reference's modules re-export by accident of their own imports (numpy, matplotlib, ...) or that belong to
# rows SURVEY.md section 8 leaves out (plots, the Parton/Jet trees, pynverse); a script that takes these through a
# star-import gets them from their home module
NOT_ON_THE_PATH = {
    "np", "plt", "random", "math", "warnings", "interpolate", "scipy", "derivative", "inversefunc", "dill", "os", "sys",
    "ABC", "abstractmethod", "dataclass", "Callable", "Path", "time", "cm", "mpl", "matplotlib", "polylog", "hyp2f1",
    "gamma", "gaussian_filter1d", "mpmath", "fsolve", "reduce", "TwoParticleJetAlgorithm", "spence", "beta",
    "eec",       # the analytic EEC curve: a plot-side truth curve (out of scope)
}


def _reference_uses():
    """{module: set of names the reference's scripts use from it}"""
    root = ref_shim.REFERENCE_ROOT
    uses = {m: set() for m in MIRRORED}
    for path in glob.glob(os.path.join(root, "examples", "**", "*.py"), recursive=True):
        try:
            tree = ast.parse(open(path).read())
        except SyntaxError:
            continue
        loaded = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name)} | \
                 {n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute)}
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("jetmontecarlo."):
                mod = node.module[len("jetmontecarlo."):]
                if mod not in uses:
                    continue
                for alias in node.names:
                    if alias.name == "*":
                        uses[mod] |= {("*", name) for name in loaded}
                    else:
                        uses[mod].add(("explicit", alias.name))
    return uses


def test_names_the_reference_scripts_import_exist_in_the_mirror():
    ref_shim.install()
    uses = _reference_uses()
    missing = []
    for mod, names in uses.items():
        ref = importlib.import_module("jetmontecarlo." + mod)
        ours = importlib.import_module("jetmontecarlo_b200." + mod)
        public = {n for n in vars(ref) if not n.startswith("_")}
        for kind, name in sorted(names):
            if name in NOT_ON_THE_PATH or name not in public:
                continue                         # not a name of that module (a star-import supplies only module names)
            owner = str(getattr(getattr(ref, name), "__module__", "") or "")
            if inspect.ismodule(getattr(ref, name)):
                continue
            if owner and not owner.startswith("jetmontecarlo."):
                continue                         # numpy / scipy / stdlib objects the module happens to hold
            if owner.startswith("jetmontecarlo.") and owner[len("jetmontecarlo."):] not in MIRRORED:
                continue                         # re-exported from a module that is not on the path (plots, vectors)
            if not hasattr(ours, name):
                missing.append("%s.%s" % (mod, name))
    assert not missing, "names used by the reference's scripts but absent from the mirror: %s" % sorted(set(missing))


HOT_PATH_CALLABLES = {
    "montecarlo.sampler": ["simpleSampler", "sampler.generateSamples", "sampler.saveSamples", "sampler.loadSamples"],
    "numerics.radiators.samplers": ["criticalSampler", "precriticalSampler", "ungroomedSampler"],
    "montecarlo.integrator": ["integrator.setBins", "integrator.setDensity", "integrator.integrate",
                              "integrator.setLastBinBndCondition", "integrator.setFirstBinBndCondition",
                              "integrator.setMonotone", "integrator.makeInterpolatingFn", "integrate_1d",
                              "integrator_2d.setBins", "integrator_2d.setDensity", "integrator_2d.integrate"],
    "numerics.observables": ["C_groomed", "C_ungroomed", "C_ungroomed_max"],
    "numerics.weights": ["radiatorWeight", "criticalEmissionWeight", "precriticalEmissionWeight", "twoEmissionWeight"],
    "analytics.qcd_utils": ["alpha1loop", "alpha_s", "splittingFn", "CR"],
    "analytics.radiators.fixedcoupling": ["critPDFAnalytic_fc_LL", "subRadAnalytic_fc_LL", "subRadPrimeAnalytic_fc_LL",
                                          "subPDFAnalytic_fc_LL", "critRadAnalytic_fc_LL"],
    "analytics.radiators.running_coupling": ["critRadAnalytic", "critPDFAnalytic", "subRadAnalytic",
                                             "subRadPrimeAnalytic", "subPDFAnalytic"],
    "numerics.radiators.generation": ["gen_numerical_radiator", "gen_pre_num_rad", "gen_crit_sub_num_rad"],
    "numerics.splitting": ["gen_normalized_splitting"],
    "utils.partonshower_utils": ["gen_jets", "getECFs_ungroomed", "getECFs_softdrop", "getECFs_rss", "getECFs"],
    "utils.hist_utils": ["vals_to_pdf"],
    "utils.montecarlo_utils": ["getLinSample", "getLogSample", "getLogSample_zerobin", "samples_from_cdf",
                               "inverse_transform_samples"],
    "montecarlo.process": ["BasicProcess", "KinematicProcess", "MonteCarloProcess", "MonteCarloObservable"],
    "montecarlo.ewoc": ["Npoint_MonteCarloEWOC", "Npoint_MonteCarloEWOC.set_weights",
                        "Npoint_MonteCarloEWOC.generate_ewoc"],
}


def _resolve(module, dotted):
    obj = module
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj.__init__ if inspect.isclass(obj) else obj


def test_hot_path_signatures_accept_the_reference_arguments():
    ref_shim.install()
    problems = []
    for mod, names in HOT_PATH_CALLABLES.items():
        ref_mod = importlib.import_module("jetmontecarlo." + mod)
        our_mod = importlib.import_module("jetmontecarlo_b200." + mod)
        for dotted in names:
            ref_params = list(inspect.signature(_resolve(ref_mod, dotted)).parameters.values())
            our_params = list(inspect.signature(_resolve(our_mod, dotted)).parameters.values())
            ours_by_name = {p.name: p for p in our_params}
            for i, rp in enumerate(ref_params):
                if rp.kind in (rp.VAR_KEYWORD, rp.VAR_POSITIONAL):
                    continue
                op = ours_by_name.get(rp.name)
                if op is None:
                    if any(p.kind == p.VAR_KEYWORD for p in our_params):
                        continue                 # swallowed by **kwargs, as the reference's own subclasses do
                    problems.append("%s.%s: parameter %r missing" % (mod, dotted, rp.name))
                    continue
                if i < len(our_params) and our_params[i].name != rp.name and rp.default is rp.empty:
                    problems.append("%s.%s: positional parameter %d is %r, reference has %r"
                                    % (mod, dotted, i, our_params[i].name, rp.name))
                if rp.default is not rp.empty and op.default is not op.empty and repr(rp.default) != repr(op.default) \
                        and not callable(rp.default):
                    problems.append("%s.%s: default of %r is %r, reference has %r"
                                    % (mod, dotted, rp.name, op.default, rp.default))
            # anything the mirror adds must be optional
            ref_names = {p.name for p in ref_params}
            for op in our_params:
                if op.name not in ref_names and op.default is op.empty and op.kind not in (op.VAR_KEYWORD, op.VAR_POSITIONAL):
                    problems.append("%s.%s: extra required parameter %r" % (mod, dotted, op.name))
    assert not problems, "\n".join(problems)


def test_install_as_aliases_the_reference_names(monkeypatch):
    import sys
    import jetmontecarlo_b200
    saved = {k: v for k, v in sys.modules.items() if k == "jetmontecarlo" or k.startswith("jetmontecarlo.")}
    for k in saved:
        monkeypatch.delitem(sys.modules, k, raising=False)
    try:
        aliased = jetmontecarlo_b200.install_as("jetmontecarlo")
        for mod in MIRRORED:
            assert "jetmontecarlo." + mod in aliased
            assert sys.modules["jetmontecarlo." + mod] is importlib.import_module("jetmontecarlo_b200." + mod)
        ns = {}
        exec("from jetmontecarlo.montecarlo.integrator import *\n"
             "from jetmontecarlo.numerics.radiators.samplers import *\n"
             "from jetmontecarlo.numerics.observables import *\n"
             "from jetmontecarlo.numerics.weights import *\n"
             "from jetmontecarlo.analytics.sudakov_factors.fixedcoupling import *\n", ns)
        for name in ("integrator", "integrator_2d", "integrate_1d", "criticalSampler", "ungroomedSampler",
                     "precriticalSampler", "C_groomed", "C_ungroomed", "criticalEmissionWeight", "twoEmissionWeight",
                     "critSudakov_fc_LL"):
            assert name in ns, name
        assert ns["integrator"].__module__ == "jetmontecarlo_b200.montecarlo.integrator"
    finally:
        for k in [k for k in sys.modules if k == "jetmontecarlo" or k.startswith("jetmontecarlo.")]:
            del sys.modules[k]
        sys.modules.update(saved)
