"""Batch sharding of the hot path across the GPUs of one box (one process per GPU).

Rows of the MMA tensors ([bsz*heads, tgt, src]) and of CIF ([bsz, src, C]) never interact, so
the path shards by UTTERANCE with no data-path collective: rank r owns a contiguous block of
utterances and, for MMA, all heads of those utterances (so the bmm(beta, v) that follows stays
local).  The only collective of a training step is the all-reduce of parameter gradients, which
belongs to the surrounding trainer (fairseq DDP in the reference); `allreduce_gradients` is the
stand-in used by bench.py."""
from typing import Tuple

import torch
import torch.distributed as dist


def utterance_shard(n_utterances: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) utterance range of `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_utterances, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def row_shard(n_utterances: int, heads: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Row range in the flattened [bsz*heads, ...] layout (utterance-major, as fairseq's
    view(bsz * num_heads, ...) produces)."""
    b0, b1 = utterance_shard(n_utterances, world_size, rank)
    return b0 * heads, b1 * heads


def allreduce_gradients(buf: torch.Tensor, async_op: bool = True):
    """Sum-all-reduce of a flat gradient buffer (NCCL on GPUs, gloo in the CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    return dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=async_op)


def max_over_ranks(value: float, device=None) -> float:
    """Step time of the job = slowest rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
