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


class PeerRows:
    """``h1``/``h2`` in buffers every rank of the node has mapped (CUDA IPC over NVLink): the stage
    kernels of this rank's context store each row another shard may read straight into the other
    ranks' copies (``gvc_stage_peers``), so the exchange between two stages shrinks to a barrier.
    One process per GPU, at most 8 ranks, NCCL group for the handle exchange and the barrier."""

    def __init__(self, ctx, n_global: int, group=None, bounds=None):
        """bounds (the vertex ranges of all ranks, as in Shard.bounds): when given, a row only
        travels to the ranks that own a neighbour of its vertex instead of to all of them.
        Collective: every rank of the group must call it.  Raises RuntimeError on EVERY rank if
        any rank could not allocate or map the buffers (the ranks agree before they go on, so a
        caller can fall back to exchange_rows without hanging anybody)."""
        self.ctx, self.group = ctx, group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        # (a context may name its own torch device and tensor wrapper: the CPU stand-in of
        # tests/test_dist_cpu.py does, to run this protocol over gloo and shared host memory)
        dev = ctx.torch_device() if hasattr(ctx, "torch_device") else torch.device("cuda", ctx.device)
        self._dev = dev
        wrap = ctx.peer_tensor if hasattr(ctx, "peer_tensor") else (lambda ptr, shape: _as_tensor(ptr, shape, dev))
        nbytes = int(n_global) * 16 * 4
        self.own, self.mapped, handles, err = [], [[], []], b"", None

        def agree(what):
            ok = torch.tensor([0 if err else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self._abandon()
                raise RuntimeError(f"peer memory unavailable ({what}): {err or 'another rank failed'}")

        try:
            if world - 1 > 7:
                raise RuntimeError(f"{world} ranks; the stage kernels mirror into at most 7 peers")
            for _ in range(2):
                ptr, h = ctx.peer_alloc(nbytes)
                self.own.append(ptr)
                handles += h
        except Exception as e:           # noqa: BLE001 -- reported on every rank below
            err = e
            handles = bytes(128)
        agree("allocation")
        mine = torch.tensor(list(handles), dtype=torch.uint8, device=dev)
        everybody = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(everybody, mine, group=group)
        try:
            for r in range(world):
                if r == rank:
                    continue
                hb = bytes(everybody[r].cpu().tolist())
                for k in range(2):
                    self.mapped[k].append(ctx.peer_open(hb[64 * k:64 * (k + 1)]))
        except Exception as e:           # noqa: BLE001
            err = e
        agree("mapping the other ranks' buffers")
        ctx.stage_peers(0, self.mapped[0])
        ctx.stage_peers(1, self.mapped[1])
        if bounds is not None:
            ctx.peer_owners(bounds, [-1 if r == rank else (r if r < rank else r - 1) for r in range(world)])
        self.h1 = wrap(self.own[0], (int(n_global), 16))
        self.h2 = wrap(self.own[1], (int(n_global), 16))
        self._token = torch.zeros(1, device=dev)

    def _abandon(self) -> None:
        """Give back what this rank got when the set-up failed somewhere (no collectives)."""
        for k in range(2):
            for p in self.mapped[k]:
                try:
                    self.ctx.peer_close(p)
                except Exception:        # noqa: BLE001
                    pass
        for p in self.own:
            try:
                self.ctx.peer_free(p)
            except Exception:            # noqa: BLE001
                pass
        self.own, self.mapped = [], [[], []]

    def barrier(self) -> None:
        """Every rank's stage kernel (and with it its stores into our buffers) has finished: a
        one-element all-reduce ordered on the current stream, no host synchronisation."""
        dist.all_reduce(self._token, group=self.group)

    def close(self) -> None:
        self.ctx.stage_peers(0, [])
        self.ctx.stage_peers(1, [])
        self.ctx.peer_owners([0], [])
        if self._dev.type == "cuda":
            torch.cuda.synchronize()
        dist.barrier(self.group)                 # nobody still writes into a buffer that is about to go
        for k in range(2):
            for p in self.mapped[k]:
                self.ctx.peer_close(p)
        dist.barrier(self.group)
        self.h1 = self.h2 = None
        for p in self.own:
            self.ctx.peer_free(p)


class _RawCuda:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2,
                                         "strides": None}


def _as_tensor(ptr: int, shape, dev) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, shape), device=dev)


def sharded_forward(stage_fn, shard: Shard, x_full: torch.Tensor, h1: torch.Tensor, h2: torch.Tensor,
                    scores_local: torch.Tensor, weight_scale: float, mode: int, group=None,
                    before_exchange=None, peer_rows: PeerRows = None) -> None:
    """The three stages with the two row exchanges between them.

    stage_fn(stage, d_in, d_out, weight_scale, mode) enqueues one fused stage for this rank's
    shard (``Context.stage_device``).  ``before_exchange()`` must make the stage's output
    visible to the communication stream (stream sync for CUDA; nothing on CPU).  With
    ``peer_rows`` (h1/h2 must be its buffers) the kernels have already delivered the rows and
    the exchange is a barrier."""
    def exchange(h):
        if peer_rows is not None:
            peer_rows.barrier()
        else:
            exchange_rows(h, shard.bounds, group, shard.live)
    stage_fn(0, x_full, h1, weight_scale, mode)
    if before_exchange:
        before_exchange()
    exchange(h1)
    stage_fn(1, h1, h2, weight_scale, mode)
    if before_exchange:
        before_exchange()
    exchange(h2)
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
