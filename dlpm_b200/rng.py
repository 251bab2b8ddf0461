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


def default_state():
    return _default


def manual_seed(seed, offset=0):
    """Seed the noise generator (mirrors ``torch.manual_seed`` / ``np.random.seed`` in the reference)."""
    _default.seed = int(seed) & (2 ** 64 - 1)
    _default.offset = int(offset)


def set_sample_base(base):
    """Global index of this process's first sample (batch sharding over GPUs)."""
    _default.sample_base = int(base)
