"""Data parallelism over clouds: one process per GPU, one gradient all-reduce per step.

The reference has no distributed code at all (SURVEY §2 "Parallelism strategies": none); the hot path
shards naturally over clouds: each rank owns a contiguous, even-sized chunk of the
interleaved Siamese batch so that pairs (rows 2b, 2b+1) stay on one GPU.  The only exchange is the SUM
all-reduce of the flat fp32 gradient buffer (3.94 MB for the S3DIS net), issued on the compute stream
right after the last weight-gradient kernel; the 1/world_size scaling is folded into the Adam kernel
(`gscale`), so there is no extra pass over the buffer.  BatchNorm statistics stay per-rank (standard
DDP semantics; BASELINE north-star: "allreduce ... on gradients only").
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, backend: str | None = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        self.enabled = self.world_size > 1
        if self.enabled and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if self.backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", rank=self.rank, world_size=self.world_size,
                                        device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(self.backend, rank=self.rank, world_size=self.world_size)

    # -- the one collective on the data path ------------------------------------------------------
    def all_reduce(self, flat_grad: torch.Tensor) -> None:
        if self.enabled:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)

    def all_reduce_async(self, t: torch.Tensor):
        """start the SUM all-reduce of a slice of the gradient buffer: NCCL runs it on its own stream once the kernels
        enqueued so far on the current stream have finished; the returned handle's wait() orders the current stream after it"""
        if self.enabled:
            return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
        return None

    def barrier(self) -> None:
        if self.enabled:
            dist.barrier()

    def max_over_ranks(self, value: float, device=None) -> float:
        if not self.enabled:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device if self.backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value: float, device=None) -> float:
        if not self.enabled:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device if self.backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def broadcast_params(self, flat: torch.Tensor) -> None:
        """all ranks start from rank 0's variables (tf.global_variables_initializer ran once in the reference)"""
        if self.enabled:
            dist.broadcast(flat, src=0)

    def shutdown(self) -> None:
        if self.enabled and dist.is_initialized():
            dist.destroy_process_group()


def shard_pairs(n_samples_global: int, rank: int, world_size: int):
    """[lo, hi) Siamese-sample range owned by `rank`: contiguous, equal-sized, pairs never split."""
    if n_samples_global % world_size:
        raise ValueError(f"{n_samples_global} samples do not split evenly over {world_size} ranks")
    per = n_samples_global // world_size
    return rank * per, (rank + 1) * per


def attach(trainer, dp: DataParallel) -> None:
    """make `trainer.train_batch` all-reduce its gradients; synchronise the initial variables."""
    trainer.dist = dp
    # Overlap (63 % of the gradient bytes belong to seg/conv1): the head and adj_conv7 gradients are complete long
    # before the EdgeConv blocks', and they are the contiguous tail of the flat buffer.  The engine calls the hook as soon as
    # they are enqueued; the all-reduce of that slice then runs on NCCL's stream under the rest of the backward pass.
    eng = trainer.engine
    tail = eng.vs.tail_offset(("adj_conv7/", "seg/")) if hasattr(eng.vs, "tail_offset") else None
    trainer._tail_work, trainer._tail_off = None, tail
    if dp.enabled and tail is not None and hasattr(eng, "on_head_grads_ready"):
        def start_tail():
            trainer._tail_work = dp.all_reduce_async(eng.vs.grad[tail:])
        eng.on_head_grads_ready = start_tail
    # every rank draws its own dropout masks (one Philox stream per rank), as independent tf.nn.dropout ops would
    trainer.engine.seed = int(trainer.engine.seed) + 104729 * dp.rank
    if dp.enabled:
        dp.broadcast_params(trainer.engine.vs.theta)
        dp.broadcast_params(trainer.engine.vs.state)
