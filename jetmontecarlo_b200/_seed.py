"""Global sequence of Philox keys: the analogue of Python's global MT19937 state the reference relies on."""
import itertools
import os

# ---- seeding -----------------------------------------------------------------
# The reference relies on Python's global MT19937 state.  Here every
# generateSamples call takes the next 64-bit Philox key of a global sequence
# seeded by set_seed (or by the OS), so separate samplers are independent.
_base_seed = int.from_bytes(os.urandom(8), "little")
_stream = itertools.count()


def set_seed(seed):
    """Seed the global stream sequence (the analogue of ``random.seed``)."""
    global _base_seed, _stream
    _base_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    _stream = itertools.count()


def _mix64(x):
    """splitmix64 finaliser: decorrelates consecutive stream ids."""
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def next_key():
    """Philox key of the next sampler run."""
    return _mix64(_base_seed + next(_stream))
