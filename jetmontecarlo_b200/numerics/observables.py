"""Energy-correlation observables per sample -- host mirror of
jetmontecarlo/numerics/observables.py:15-114."""
from .. import _lib
from .._arrayfn import eval_fn


def C_groomed(z_crit, theta_crit, z_cut, beta, z_pre=0., f=1., acc='LL', verbose=0):
    """Groomed ECF of a critical emission; Soft Drop / mMDT is
    ``z_pre=z_cut, f=0``.  ref: numerics/observables.py:15-79"""
    assert acc in ['LL', 'MLL'], "Accuracy must be LL or MLL!"
    return eval_fn(_lib.FN_C_GROOMED, [z_crit, theta_crit, z_pre, f], acc=_lib.ACCS[acc], z_cut=z_cut,
                   beta=beta, scalar_fields={2: "z_pre", 3: "f_soft"})


def C_ungroomed(z, theta, beta, acc='LL'):
    """ref: numerics/observables.py:82-109"""
    assert acc in ['LL', 'MLL'], "Accuracy must be LL or MLL"
    return eval_fn(_lib.FN_C_UNGROOMED, [z, theta], acc=_lib.ACCS[acc], beta=beta)


def C_ungroomed_max(beta, radius, acc):
    """ref: numerics/observables.py:111-114"""
    assert acc in ['LL', 'MLL'], "Accuracy must be LL or MLL"
    if acc == 'LL':
        return radius ** beta / 2.
    return radius ** beta / 4.
