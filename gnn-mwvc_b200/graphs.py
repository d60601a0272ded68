"""Synthetic weighted graphs in the reference's input format.

The reference reads METIS-format vertex-weighted graphs (parse_graph,
reference src/GNN_VC.cpp:34-91; format README.md:45-60): ``N E 10`` then one line
per vertex ``weight nbr nbr ...`` with 1-indexed ascending neighbours.  The
generators below produce the same information as arrays: a sorted, unique list of
undirected edges (u < v), integer vertex weights, and the CSR adjacency in the
order the reference's reduction_graph ctor builds it (ascending per vertex,
include/reduction_graph.hpp:103-128).

Everything is written with torch tensor ops so the big benchmark graphs are
generated on the GPU (seconds instead of minutes); on CPU tensors the same code
serves the small parity cases.  Workloads follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Graph:
    """CSR view of a vertex-weighted undirected graph (what predict() reads)."""
    n: int
    row_ptr: torch.Tensor   # int64 [n+1]
    col: torch.Tensor       # int32 [2E] (values are uint32 ids, ascending per row)
    weights: torch.Tensor   # int32 [n]  (uint32 vertex weights, W(u))
    nw: torch.Tensor        # int32 [n]  (uint32 neighbourhood weights, NW(u))
    eu: torch.Tensor | None = None   # int64 [E] undirected edges, u < v, sorted
    ev: torch.Tensor | None = None
    name: str = "graph"

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    @property
    def n_edges(self) -> int:
        return self.nnz // 2

    def numpy(self):
        """(row_ptr u64, col u32, W u32, NW u32) as numpy arrays on the host."""
        return (self.row_ptr.cpu().numpy().astype(np.uint64),
                self.col.cpu().numpy().view(np.uint32),
                self.weights.cpu().numpy().view(np.uint32),
                self.nw.cpu().numpy().view(np.uint32))

    def edges_numpy(self):
        return (self.eu.cpu().numpy().astype(np.uint32), self.ev.cpu().numpy().astype(np.uint32))


def _canonical_edges(u: torch.Tensor, v: torch.Tensor, n: int):
    """Drop self loops, orient u<v, sort, de-duplicate (parse_graph :62-64,:86-87)."""
    keep = u != v
    u, v = u[keep], v[keep]
    lo, hi = torch.minimum(u, v), torch.maximum(u, v)
    key = torch.unique(lo * n + hi)          # sorted
    return key // n, key % n


def to_u32(t: torch.Tensor) -> torch.Tensor:
    """int64 values in [0, 2^32) -> int32 tensor holding the same 32 bits."""
    return torch.where(t >= 2 ** 31, t - 2 ** 32, t).to(torch.int32)


def graph_from_edges(n: int, eu: torch.Tensor, ev: torch.Tensor, weights: torch.Tensor,
                     name: str = "graph") -> Graph:
    """Build the CSR the reduction_graph ctor would (reduction_graph.hpp:103-128)."""
    dev = eu.device
    src = torch.cat([eu, ev])
    dst = torch.cat([ev, eu])
    order = torch.argsort(src * n + dst)     # ascending neighbours per vertex
    src, dst = src[order], dst[order]
    deg = torch.bincount(src, minlength=n)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(deg, 0)
    w64 = weights.to(torch.int64)
    nw = torch.zeros(n, dtype=torch.int64, device=dev)
    nw.index_add_(0, src, w64[dst])
    assert int(nw.max().item() if n else 0) < 2 ** 32, "NW must fit uint32 (reference Tw)"
    return Graph(n=n, row_ptr=row_ptr, col=to_u32(dst), weights=to_u32(w64), nw=to_u32(nw),
                 eu=eu, ev=ev, name=name)


def random_weights(n: int, seed: int, device="cpu", lo: int = 1, hi: int = 200) -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(lo, hi + 1, (n,), generator=g, device=device, dtype=torch.int64)


def er_graph(n: int, m: int, seed: int = 1, device="cpu") -> Graph:
    """G(n, m): exactly m distinct undirected edges, uniform, weights 1..200."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    eu = torch.empty(0, dtype=torch.int64, device=device)
    ev = torch.empty(0, dtype=torch.int64, device=device)
    max_m = n * (n - 1) // 2
    m = min(m, max_m)
    while eu.numel() < m:
        need = m - eu.numel()
        u = torch.randint(0, n, (need + need // 8 + 16,), generator=g, device=device)
        v = torch.randint(0, n, (need + need // 8 + 16,), generator=g, device=device)
        eu, ev = _canonical_edges(torch.cat([eu, u]), torch.cat([ev, v]), n)
    if eu.numel() > m:   # drop a random surplus, keep sorted order
        keep = torch.randperm(eu.numel(), generator=g, device=device)[:m].sort().values
        eu, ev = eu[keep], ev[keep]
    return graph_from_edges(n, eu, ev, random_weights(n, seed + 1, device), name=f"er_{n}_{m}")


def rmat_graph(scale: int, edge_factor: int = 16, seed: int = 42, device="cpu",
               abcd=(0.57, 0.19, 0.19, 0.05), n_limit: int | None = None,
               chunk: int = 1 << 24) -> Graph:
    """R-MAT (Graph500 parameters), symmetrised, de-duplicated, self loops
    dropped; SURVEY.md 8(d) configs 2 and 4.  ``n_limit`` rejects endpoints >= it."""
    n_ids = 1 << scale
    n = n_limit if n_limit is not None else n_ids
    target = edge_factor * n
    a, b, c, _ = abcd
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    keys = []
    made = 0
    while made < target:
        cnt = min(chunk, target - made)
        u = torch.zeros(cnt, dtype=torch.int64, device=device)
        v = torch.zeros(cnt, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(cnt, generator=g, device=device)
            ubit = (r >= a + b).to(torch.int64)                       # quadrants c, d
            vbit = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)   # b, d
            u = (u << 1) | ubit
            v = (v << 1) | vbit
        if n_limit is not None:
            ok = (u < n) & (v < n)
            u, v = u[ok], v[ok]
        lo, hi = torch.minimum(u, v), torch.maximum(u, v)
        ok = lo != hi
        keys.append(lo[ok] * n + hi[ok])
        made += cnt
    key = torch.unique(torch.cat(keys))
    del keys
    eu, ev = key // n, key % n
    return graph_from_edges(n, eu, ev, random_weights(n, seed + 1, device),
                            name=f"rmat{scale}_ef{edge_factor}")


def grid_graph(rows: int, cols: int, seed: int = 7, device="cpu") -> Graph:
    """2-D 4-neighbour grid, row-major ids; SURVEY.md 8(d) config 3."""
    n = rows * cols
    idx = torch.arange(n, dtype=torch.int64, device=device)
    r, c = idx // cols, idx % cols
    right = idx[c < cols - 1]
    down = idx[r < rows - 1]
    eu = torch.cat([right, down])
    ev = torch.cat([right + 1, down + cols])
    key = torch.unique(eu * n + ev)
    return graph_from_edges(n, key // n, key % n, random_weights(n, seed, device),
                            name=f"grid_{rows}x{cols}")


def er10k_fixture() -> Graph:
    """Config 1 exactly as SURVEY.md App. D generated it (python ``random.seed(1)``)
    so the METIS file, and the reference run on it, are reproducible byte for byte."""
    rnd = random.Random(1)
    n, m = 10000, 50000
    seen = set()
    while len(seen) < m:
        u = rnd.randrange(n)
        v = rnd.randrange(n)
        if u == v:
            continue
        e = (min(u, v), max(u, v))
        if e in seen:
            continue
        seen.add(e)
    weights = [rnd.randint(1, 200) for _ in range(n)]
    e = sorted(seen)
    eu = torch.tensor([p[0] for p in e], dtype=torch.int64)
    ev = torch.tensor([p[1] for p in e], dtype=torch.int64)
    return graph_from_edges(n, eu, ev, torch.tensor(weights, dtype=torch.int64), name="er10k")


def write_metis(g: Graph, path) -> None:
    """METIS text as the reference parses it (src/GNN_VC.cpp:44-67)."""
    row_ptr, col, w, _ = g.numpy()
    lines = [f"{g.n} {g.n_edges} 10"]
    for u in range(g.n):
        nb = col[int(row_ptr[u]):int(row_ptr[u + 1])].astype(np.int64) + 1
        lines.append(f"{int(w[u])} " + " ".join(map(str, nb.tolist())))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def nnz_balanced_ranges(row_ptr: torch.Tensor, parts: int, align: int = 32):
    """Contiguous vertex ranges with ~equal nnz + per-vertex cost.
    Returns a list of ``parts + 1`` boundaries, multiples of ``align`` except the last."""
    n = row_ptr.numel() - 1
    cost = row_ptr[1:].to(torch.float64) + 16.0 * torch.arange(1, n + 1, device=row_ptr.device,
                                                               dtype=torch.float64)
    total = float(cost[-1].item()) if n else 0.0
    bounds = [0]
    for p in range(1, parts):
        t = total * p / parts
        i = int(torch.searchsorted(cost, torch.tensor([t], dtype=torch.float64, device=cost.device)).item())
        i = min(n, (i + align - 1) // align * align)
        bounds.append(max(i, bounds[-1]))
    bounds.append(n)
    return bounds


def live_rows(row_ptr: torch.Tensor, bounds) -> list:
    """Per shard: how many of its leading rows can be a neighbour of anybody, i.e. 1 + the local
    index of its last vertex with a non-empty adjacency list (symmetric graphs: an isolated vertex
    is nobody's neighbour, its 16-float row never has to leave the GPU that computes it).  After
    balanced_relabel every shard is sorted by descending degree, so this prefix is exactly its
    non-isolated vertices."""
    deg = row_ptr[1:] - row_ptr[:-1]
    out = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        nz = torch.nonzero(deg[a:b] > 0)
        out.append(int(nz[-1].item()) + 1 if nz.numel() else 0)
    return out


def relabel(g: Graph, perm: torch.Tensor, n_new: int, name: str) -> Graph:
    """Rename vertex v -> perm[v] (a one-to-one map into [0, n_new)); ids without a preimage become
    isolated vertices of weight 1.  Adjacency lists keep their order, only the names change, so
    per-vertex results are bit-identical: scores_new[perm[v]] == scores_old[v]."""
    dev = g.row_ptr.device
    deg = g.row_ptr[1:] - g.row_ptr[:-1]
    new_deg = torch.zeros(n_new, dtype=torch.int64, device=dev)
    new_deg[perm] = deg
    row_ptr = torch.zeros(n_new + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(new_deg, 0)
    # move every adjacency list to its new place, ids renamed, order kept
    src_new = torch.repeat_interleave(perm, deg)                    # new owner of every entry (old entry order)
    within = torch.arange(g.nnz, dtype=torch.int64, device=dev) - torch.repeat_interleave(g.row_ptr[:-1], deg)
    dst = row_ptr[src_new] + within
    col_old = g.col.to(torch.int64) & 0xFFFFFFFF
    col = torch.empty(g.nnz, dtype=torch.int32, device=dev)
    col[dst] = to_u32(perm[col_old])
    weights = torch.ones(n_new, dtype=torch.int32, device=dev)
    weights[perm] = g.weights
    nw = torch.zeros(n_new, dtype=torch.int32, device=dev)
    nw[perm] = g.nw
    return Graph(n=n_new, row_ptr=row_ptr, col=col, weights=weights, nw=nw, name=name)


def cyclic_relabel(g: Graph, parts: int):
    """Relabel vertex v -> (v % parts) * ceil(n / parts) + v // parts so that contiguous ranges of
    the new ids hold every parts-th original vertex: equal-sized shards (one NCCL all-gather per
    exchange instead of per-owner broadcasts).  Balances the WORK only when the degree does not
    depend on the low bits of the id -- on R-MAT it does (see balanced_relabel).
    Returns (relabelled graph on ceil(n/parts)*parts ids -- padding vertices are isolated, weight 1 --,
    perm [n] int64)."""
    n = g.n
    per = (n + parts - 1) // parts
    v = torch.arange(n, dtype=torch.int64, device=g.row_ptr.device)
    perm = (v % parts) * per + v // parts
    return relabel(g, perm, per * parts, g.name + f"_cyc{parts}"), perm


def balanced_relabel(g: Graph, parts: int):
    """Equal-sized shards with equal work: the vertices are dealt to the shards in order of
    descending degree (the i-th largest goes to shard i % parts, slot i // parts), so every shard
    gets the same number of vertices, the same number of adjacency entries up to one maximum
    degree, and its share of the hubs.  An R-MAT id with zero low bits has several times the
    expected degree, which is why dealing by id (cyclic_relabel) leaves shard 0 with about half of
    all entries.  Same return value and bit-identity property as cyclic_relabel."""
    n = g.n
    per = (n + parts - 1) // parts
    deg = g.row_ptr[1:] - g.row_ptr[:-1]
    by_degree = torch.argsort(deg, descending=True, stable=True)
    i = torch.arange(n, dtype=torch.int64, device=g.row_ptr.device)
    perm = torch.empty(n, dtype=torch.int64, device=g.row_ptr.device)
    perm[by_degree] = (i % parts) * per + i // parts
    return relabel(g, perm, per * parts, g.name + f"_bal{parts}"), perm
