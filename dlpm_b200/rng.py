"""Host-side state of the counter-based Philox generator used by every kernel.

The reference draws A with numpy's *global* RandomState (scipy) and G with torch's global generator
(``bem/Experiments.py:58-63`` seeds both).  Here one (seed, offset, sample_base) triple plays that
role: ``offset`` advances by the number of "calls" a routine consumes (one per ``generate``, ``T`` per
Sigma chain, ...) and ``sample_base`` is the global index of this rank's first sample, so a batch
sharded over N GPUs draws exactly the variates the single-GPU run would.
"""
import threading


class PhiloxState:
    def __init__(self, seed=0x5EED_D1F5, offset=0, sample_base=0):
        self.seed = int(seed) & (2 ** 64 - 1)
        self.offset = int(offset)
        self.sample_base = int(sample_base)
        self._lock = threading.Lock()

    def reserve(self, n=1):
        """Return the current offset and advance it by n."""
        with self._lock:
            o = self.offset
            self.offset += int(n)
            return o


_default = PhiloxState()
_follow_torch = True     # until dlpm_b200.manual_seed() is called, the default stream is keyed by torch's seed
_torch_seed_seen = None


def _mix(seed):
    """splitmix64 finaliser: decorrelates our Philox key from the raw user seed (torch's own generator uses it as is)."""
    z = (int(seed) + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
    return z ^ (z >> 31)


def default_state():
    """The process-wide state.  Unless ``dlpm_b200.manual_seed`` has been called it FOLLOWS ``torch.manual_seed``: the
    reference seeds its noise through ``torch.manual_seed`` / ``np.random.seed`` (bem/Experiments.py:58-63), so a run that
    only does that must not silently reuse one fixed stream -- whenever ``torch.initial_seed()`` changes, the Philox key
    is re-derived from it and the call offset restarts at 0."""
    global _torch_seed_seen
    if _follow_torch:
        import torch
        ts = int(torch.initial_seed())
        if ts != _torch_seed_seen:
            _torch_seed_seen = ts
            _default.seed = _mix(ts)
            _default.offset = 0
    return _default


def manual_seed(seed, offset=0):
    """Seed the noise generator explicitly (from here on ``torch.manual_seed`` no longer re-keys it)."""
    global _follow_torch
    _follow_torch = False
    _default.seed = int(seed) & (2 ** 64 - 1)
    _default.offset = int(offset)


def follow_torch_seed(on=True):
    """Return to (or leave) the default behaviour of keying the noise stream by ``torch.initial_seed()``."""
    global _follow_torch, _torch_seed_seen
    _follow_torch = bool(on)
    _torch_seed_seen = None


def set_sample_base(base):
    """Global index of this process's first sample (batch sharding over GPUs / over generation processes: two processes
    with the same seed and the same sample base draw the SAME samples)."""
    _default.sample_base = int(base)
