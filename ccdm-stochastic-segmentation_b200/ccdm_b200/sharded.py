"""Sample-sharded reverse chains: one process per GPU, one collective at the end.

The reference evaluates single-process: ``Tester.test_step`` fabricates the batch with
``image.repeat_interleave(N, dim=0)`` and calls the sampler once
(/root/reference/evaluation/evaluate_lidc_uncertainty.py:96-103).  The samples of that batch are
independent given (image, noise): GroupNorm, attention, posterior and draw are all per-sample, so
the batch splits over ranks with no data-path exchange.  This module keeps the call shape
(every rank passes the same full batch) and makes the result independent of the number of ranks:

* rank ``r`` of ``R`` runs the contiguous block ``shard_range(B, r, R)`` of the batch;
* the in-kernel Philox noise is keyed by the GLOBAL sample index (``model.sample_offset``), so a
  sample's trajectory is the same whether it runs on 1 or 8 GPUs;
* one ``all_gather`` of the local results (uint8 label maps, or fp32 probabilities in
  ``confidence`` mode) over NCCL / NVLink ends the chain.  Shards are padded to equal length so the
  collective is a single fixed-size ``all_gather_into_tensor``.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "padded_shard", "gather_shards", "sample_sharded", "CANONICAL_TILE_BATCH"]

CANONICAL_TILE_BATCH = 64  # DenoisingModel.tile_batch that sample_sharded uses unless the caller set one


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous block: the first ``n % world`` ranks take one extra sample."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} of world {world}")
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def padded_shard(n: int, world: int) -> int:
    """Rows every rank contributes to the gather (the largest shard)."""
    return (n + world - 1) // world


def gather_shards(local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """All-gather per-rank blocks (rank r holds rows ``shard_range(n, r, R)``) into the full [n, ...] tensor.

    The single collective of the path.  ``local`` may be shorter than the padded shard; it is padded with zeros.
    """
    world = dist.get_world_size(group)
    rows = padded_shard(n, world)
    if local.shape[0] > rows:
        raise ValueError("local shard longer than the padded shard")
    if local.shape[0] < rows:
        pad = torch.zeros((rows - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if n == world * rows:
        return out
    pieces = []
    for r in range(world):
        b, e = shard_range(n, r, world)
        pieces.append(out[r * rows:r * rows + (e - b)])
    return torch.cat(pieces, 0)


@torch.no_grad()
def sample_sharded(model, x: torch.Tensor, condition: torch.Tensor, feature_condition: Optional[torch.Tensor] = None,
                   t: Optional[torch.Tensor] = None, group=None) -> dict:
    """``model(x, condition, feature_condition, t)`` with the batch sharded over the process group.

    Every rank passes the same full-batch tensors (as the reference evaluators build them); each rank
    runs its block and the label maps / probabilities are gathered.  Returns what the single-process
    call returns: ``{"diffusion_out": [B, K, H, W]}`` (int64 one-hot view or fp32 probabilities).
    """
    if not dist.is_initialized():
        return model(x, condition, feature_condition, t)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0]
    b, e = shard_range(n, rank, world)
    old_offset = getattr(model, "sample_offset", 0)
    # Sharded draws must come from the counter RNG keyed by the GLOBAL sample index.  The reference's noise source
    # (noise='torch': each device's global generator) would make the result depend on the number of ranks and, with the
    # usual identical seeding of every rank, hand all shards the SAME noise -- correlated samples, biased GED/diversity.
    old_noise = getattr(model, "noise", "philox")
    if isinstance(old_noise, str) and old_noise != "philox":
        model.noise = "philox"
    elif not isinstance(old_noise, str):
        raise ValueError("sample_sharded: explicit noise tensors cannot be sharded; use noise='philox'")
    model.sample_offset = old_offset + b
    # batch-independent tiling: a sample's result must not depend on how many ranks share the batch (bit-identical in the
    # 'exact' mode; see DenoisingModel.tile_batch)
    old_tile = getattr(model, "tile_batch", 0)
    if not old_tile:
        model.tile_batch = CANONICAL_TILE_BATCH
    try:
        if e > b:
            out = model(x[b:e], condition[b:e], feature_condition[b:e] if feature_condition is not None else None, t)["diffusion_out"]
        else:  # more ranks than samples: contribute an empty block
            K = model.diffusion.num_classes
            conf = model.step_T_sample == "confidence"
            out = torch.zeros((0, K) + tuple(x.shape[-2:]), dtype=torch.float32 if conf else torch.int64, device=condition.device)
    finally:
        model.sample_offset = old_offset
        model.noise = old_noise
        model.tile_batch = old_tile
    if out.dtype == torch.int64:  # majority: ship 1 byte per pixel, rebuild the one-hot view after the gather
        K = out.shape[1]
        labels = gather_shards(out.argmax(dim=1).to(torch.uint8), n, group)
        full = torch.nn.functional.one_hot(labels.long(), K).permute(0, 3, 1, 2)
    else:
        full = gather_shards(out.permute(0, 2, 3, 1).contiguous(), n, group).permute(0, 3, 1, 2)
    return {"diffusion_out": full}
