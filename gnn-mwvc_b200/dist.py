"""Vertex-range sharding of the forward over several GPUs, one process per GPU.

SURVEY.md 8(e): contiguous vertex ranges with balanced work, every shard keeps the
CSR rows of its range with GLOBAL neighbour ids; ``x`` is replicated, so stage 0
needs no exchange; after stages 0 and 1 every rank publishes its 16-float rows
and receives everybody else's (the graphs of interest are expander-like, nearly
every remote row is referenced, so the exchange is an all-gather of row slices).
Per-vertex arithmetic does not depend on the sharding, hence P-GPU scores equal
1-GPU scores bit for bit.

``torch.distributed`` is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU
tests); the compute is ``Context.stage_device`` (libgvc).  The stage function is
a parameter so that the host logic here can be tested on CPU ranks.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class Shard:
    """What one rank owns: rows [v_begin, v_end) of the global graph."""
    n_global: int
    v_begin: int
    v_end: int
    bounds: list            # parts + 1 boundaries, the same on every rank
    row_ptr: torch.Tensor   # int32/int64 [n_local + 1], starts at 0
    col: torch.Tensor       # int32 [nnz_local], global ids
    weights: torch.Tensor   # int32 [n_local]
    nw: torch.Tensor        # int32 [n_local]
    live: list = None       # per rank: leading rows that other ranks may read (None: all of them)

    @property
    def n_local(self) -> int:
        return self.v_end - self.v_begin

    @property
    def nnz(self) -> int:
        return int(self.col.numel())


def make_shard(g, bounds, rank: int, skip_isolated: bool = False) -> Shard:
    """Cut rank's vertex range out of a whole graph (graphs.Graph) that lives on any device.
    skip_isolated (symmetric graphs only): exchange just the leading rows of every shard that
    hold non-isolated vertices (graphs.live_rows)."""
    from . import graphs
    a, b = bounds[rank], bounds[rank + 1]
    lo, hi = int(g.row_ptr[a].item()), int(g.row_ptr[b].item())
    return Shard(n_global=g.n, v_begin=a, v_end=b, bounds=list(bounds),
                 row_ptr=(g.row_ptr[a:b + 1] - lo).contiguous(),
                 col=g.col[lo:hi].contiguous(), weights=g.weights[a:b].contiguous(),
                 nw=g.nw[a:b].contiguous(),
                 live=graphs.live_rows(g.row_ptr, bounds) if skip_isolated else None)


def exchange_rows(full: torch.Tensor, bounds, group=None, live=None) -> None:
    """All-gather of row slices, in place: on entry ``full[bounds[r]:bounds[r+1]]`` is valid on
    rank r; on return every rank holds all rows.  Slices may differ in length.  With ``live``
    only the first live[r] rows of every slice travel (the rest is never read remotely)."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    rank = dist.get_rank(group)
    if live is None:
        views = [full[bounds[r]:bounds[r + 1]] for r in range(world)]
    else:
        # equal prefixes keep the single collective; a row more than needed does no harm
        same = max(live) if max(live) <= min(bounds[r + 1] - bounds[r] for r in range(world)) else None
        views = [full[bounds[r]:bounds[r] + (same if same is not None else live[r])] for r in range(world)]
    sizes = {v.shape[0] for v in views}
    if len(sizes) == 1:
        dist.all_gather(views, views[rank], group=group)          # equal slices: one collective
    else:
        # work-balanced ranges differ in length; one broadcast per owner, in place
        for r in range(world):
            if views[r].numel():
                dist.broadcast(views[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def sharded_forward(stage_fn, shard: Shard, x_full: torch.Tensor, h1: torch.Tensor, h2: torch.Tensor,
                    scores_local: torch.Tensor, weight_scale: float, mode: int, group=None,
                    before_exchange=None) -> None:
    """The three stages with the two row exchanges between them.

    stage_fn(stage, d_in, d_out, weight_scale, mode) enqueues one fused stage for this rank's
    shard (``Context.stage_device``).  ``before_exchange()`` must make the stage's output
    visible to the communication stream (stream sync for CUDA; nothing on CPU)."""
    stage_fn(0, x_full, h1, weight_scale, mode)
    if before_exchange:
        before_exchange()
    exchange_rows(h1, shard.bounds, group, shard.live)
    stage_fn(1, h1, h2, weight_scale, mode)
    if before_exchange:
        before_exchange()
    exchange_rows(h2, shard.bounds, group, shard.live)
    stage_fn(2, h2, scores_local, weight_scale, mode)


def gather_scores(scores_local: torch.Tensor, bounds, group=None) -> torch.Tensor:
    """Concatenate the per-rank score slices on every rank (used by tests and the e2e read-back)."""
    world = dist.get_world_size(group)
    n = bounds[-1]
    full = torch.empty(n, dtype=scores_local.dtype, device=scores_local.device)
    full[bounds[dist.get_rank(group)]:bounds[dist.get_rank(group) + 1]] = scores_local
    if world > 1:
        exchange_rows(full, bounds, group)
    return full
