"""CPU (gloo, world_size 2): the host logic of batch-sharded sampling -- shard bounds, global sample indexing of the
Philox state, and the gather that reassembles the slices in global order (the only collective on the path)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_batch_exactly():
    from dlpm_b200.distributed import shard_bounds
    for total in (0, 1, 7, 512, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert shard_bounds(4096, 8, 3) == (1536, 512)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


class FakeMethod:
    """Stands in for GenerativeLevyProcess.sample on CPU: a sample's value is a function of its GLOBAL index, exactly
    what the Philox keying guarantees on the GPU."""

    def sample(self, models, shape, get_sample_history=False, **kw):
        from dlpm_b200 import rng
        base = rng.default_state().sample_base
        idx = torch.arange(base, base + shape[0], dtype=torch.float32)
        x = idx.view(-1, *([1] * (len(shape) - 1))).expand(*shape).contiguous() * 10.0
        if get_sample_history:
            return x, torch.stack([x + k for k in range(3)])
        return x


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dlpm_b200 import rng
        from dlpm_b200.distributed import gather_samples, init_shard, sample_sharded
        start, count = init_shard(total)
        assert rng.default_state().sample_base == start
        rng.set_sample_base(0)
        out = sample_sharded(FakeMethod(), {}, [total, 2, 3])
        want = (torch.arange(total, dtype=torch.float32) * 10.0).view(-1, 1, 1).expand(total, 2, 3)
        ok = torch.equal(out, want) and rng.default_state().sample_base == 0
        out2, hist = sample_sharded(FakeMethod(), {}, [total, 2, 3], get_sample_history=True)
        ok = ok and torch.equal(out2, want) and hist.shape == (3, total, 2, 3) and torch.equal(hist[2], want + 2)
        # plain even gather
        x = torch.full((4, 2), float(rank))
        g = gather_samples(x, 4 * world)
        ok = ok and torch.equal(g[:4], torch.zeros(4, 2)) and torch.equal(g[4:], torch.ones(4, 2))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_default_noise_stream_follows_torch_seed():
    """the reference seeds its noise through torch.manual_seed / np.random.seed (bem/Experiments.py:58-63); the
    default Philox state is re-keyed whenever torch's seed changes, until dlpm_b200.manual_seed pins it."""
    import dlpm_b200
    from dlpm_b200 import rng
    rng.follow_torch_seed(True)
    try:
        torch.manual_seed(123)
        s1 = rng.default_state()
        k1, o1 = s1.seed, s1.reserve(5)
        assert o1 == 0 and rng.default_state().offset == 5
        torch.manual_seed(124)
        assert rng.default_state().seed != k1 and rng.default_state().offset == 0   # new key, call offset restarts
        torch.manual_seed(123)
        assert rng.default_state().seed == k1                                        # same seed -> same stream
        dlpm_b200.manual_seed(99)
        torch.manual_seed(5)
        assert rng.default_state().seed == 99                                        # pinned explicitly: torch no longer re-keys
    finally:
        rng.follow_torch_seed(True)


@pytest.mark.parametrize("total", [8, 7, 1])
def test_sharded_sampling_gathers_in_global_order(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}
