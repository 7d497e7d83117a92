"""Host logic of the sample-sharded chain on CPU: world_size 2 and 3 over gloo.

The chain itself needs a B200; here a stand-in model with the DenoisingModel call shape returns a
deterministic function of (global sample index, inputs), which is exactly the property the Philox
keying gives the real chain.  What is checked: shard arithmetic (ragged batches, more ranks than
samples), ``sample_offset`` handed to the model, one collective, gathered result == unsharded result.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ccdm_b200.sharded import gather_shards, padded_shard, sample_sharded, shard_range


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 2, 7, 16, 64, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == padded_shard(n, world) or n == 0
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class _FakeDiffusion:
    num_classes = 3


class _FakeModel:
    """Call shape of DenoisingModel; output depends on the GLOBAL sample index like the Philox-keyed chain."""

    def __init__(self, mode):
        self.diffusion = _FakeDiffusion()
        self.step_T_sample = mode
        self.sample_offset = 0
        self.calls = []

    def __call__(self, x, condition, feature_condition=None, t=None):
        B, K, H, W = x.shape
        self.calls.append((self.sample_offset, B))
        g = torch.arange(self.sample_offset, self.sample_offset + B).view(B, 1, 1)
        yy = torch.arange(H).view(1, H, 1)
        xx = torch.arange(W).view(1, 1, W)
        labels = (g * 7 + yy * 3 + xx + condition[:, 0].long()) % K
        if self.step_T_sample == "confidence":
            p = torch.nn.functional.one_hot(labels, K).float() * 0.5 + 0.5 / K
            return {"diffusion_out": p.permute(0, 3, 1, 2)}
        return {"diffusion_out": torch.nn.functional.one_hot(labels, K).permute(0, 3, 1, 2)}


def _worker(rank, world, port, n, mode, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        K, H, W = 3, 5, 6
        x = torch.nn.functional.one_hot(torch.randint(0, K, (n, H, W), generator=g), K).permute(0, 3, 1, 2).float()
        cond = torch.randint(0, 5, (n, 1, H, W), generator=g).float()
        ref = _FakeModel(mode)(x, cond)["diffusion_out"]
        m = _FakeModel(mode)
        m.sample_offset = 100  # a caller-provided base offset must be preserved
        ref100 = _FakeModel(mode)
        ref100.sample_offset = 100
        want = ref100(x, cond)["diffusion_out"]
        out = sample_sharded(m, x, cond, None, None)["diffusion_out"]
        b, e = shard_range(n, rank, world)
        ok = torch.equal(out, want) and out.dtype == want.dtype and m.sample_offset == 100
        ok = ok and (m.calls == ([(100 + b, e - b)] if e > b else []))
        ok = ok and not torch.equal(ref, want) if n > 0 else ok
        # the collective alone, ragged
        local = torch.arange(b, e, dtype=torch.uint8).view(-1, 1).repeat(1, 4)
        full = gather_shards(local, n)
        ok = ok and torch.equal(full[:, 0], torch.arange(n, dtype=torch.uint8))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,mode", [(2, 16, "majority"), (2, 7, "confidence"), (3, 2, "majority")])
def test_sharded_equals_unsharded_over_gloo(world, n, mode):
    port = 29500 + (os.getpid() + world * 131 + n) % 2000
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n, mode, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert dict(ret) == {r: True for r in range(world)}


def test_single_process_passthrough():
    m = _FakeModel("majority")
    x = torch.nn.functional.one_hot(torch.zeros(2, 4, 4, dtype=torch.long), 3).permute(0, 3, 1, 2).float()
    out = sample_sharded(m, x, torch.zeros(2, 1, 4, 4))["diffusion_out"]
    assert out.shape == (2, 3, 4, 4) and m.calls == [(0, 2)]
