"""Multi-GPU data parallelism over the image batch.

One process per GPU (torchrun).  The path has no cross-sample reduction in steady state, so ranks never
talk to each other per step: the only collective is the one-time replication of the (calibrated,
prequantised) checkpoint from rank 0 over NCCL/NVLink; after that every rank denoises its own slice of
the batch with its own replica.  On CPU test runs the same code uses the gloo backend.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process => (0, 0, 1))."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend: str = "nccl") -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` samples owned by `rank`; the first total % world ranks get one
    extra sample.  Ranks beyond `total` get an empty slice."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_request(request: Dict[str, Tensor], rank: int, world: int) -> Dict[str, Tensor]:
    """Rows [lo, hi) of every per-sample tensor of a request (img, img_ids, txt, txt_ids, y, guidance)."""
    total = request["img"].shape[0]
    lo, hi = shard_range(total, rank, world)
    return {k: (v[lo:hi] if isinstance(v, Tensor) and v.dim() > 0 and v.shape[0] == total else v)
            for k, v in request.items()}


@torch.no_grad()
def broadcast_state(model: nn.Module, src: int = 0, bucket_bytes: int = 256 << 20) -> int:
    """Replicate rank `src`'s checkpoint state (parameters and buffers, including F8Linear's fp8 bytes and
    scale buffers) to every rank.  Tensors are viewed as bytes and coalesced into buckets so the ~12 GB of
    Flux-dev state goes out in a few dozen large NCCL broadcasts.  Returns the number of bytes sent.
    Shapes/dtypes must already agree on all ranks (same spec, same quantisation flow)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    tensors: List[Tensor] = [t for _, t in sorted(model.state_dict(keep_vars=True).items()) if t is not None]
    total = 0
    bucket: List[Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.detach().reshape(-1).view(torch.uint8) for t in bucket])
        dist.broadcast(flat, src=src)
        off = 0
        for t in bucket:
            n = t.numel() * t.element_size()
            t.detach().reshape(-1).view(torch.uint8).copy_(flat[off:off + n])
            off += n
        bucket, size = [], 0

    for t in tensors:
        if not t.is_contiguous():
            raise ValueError("broadcast_state needs contiguous state tensors")
        n = t.numel() * t.element_size()
        if size + n > bucket_bytes:
            flush()
        bucket.append(t)
        size += n
        total += n
    flush()
    # the buffers were rewritten IN PLACE: values derived from them (F8Linear quantising scales, RMSNorm fp32 weights,
    # the batched-modulation table) must be recomputed from the received state, not from what this rank held before
    from .blocks import invalidate_derived

    invalidate_derived(model)
    return total


def frozen_flags_sync(model: nn.Module) -> None:
    """After broadcast_state the scale buffers are identical everywhere; mark every F8Linear frozen on the
    receiving ranks too (calibration ran on rank 0 only -- amax is a whole-batch statistic,
    float8_quantize.py:227, so it must not be recomputed per shard)."""
    from .blocks import invalidate_derived
    from .f8linear import F8Linear

    for m in model.modules():
        if isinstance(m, F8Linear) and m.input_scale is not None:
            m.input_scale_initialized = True
            m.trial_index = m.num_scale_trials
    invalidate_derived(model)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
