"""Multi-GPU driver of the hot path: one process per GPU, torch.distributed for the plumbing.

SURVEY.md §8(e): the index is replicated in every GPU's HBM; the traversal (phases 2-3) is sharded
by position-contiguous slices of a shallow frontier (e2i_navigate_shard); the shards' bit writes
land all over [0, n), so ONE exchange step is needed -- a bitwise OR of the 3n-bit LCP vectors
(+ n-bit DA in mode -2).  Every bit has a single writer and the vectors start at zero, so an
integer SUM all-reduce of the 32-bit words is exactly that OR (no carries).  Phase 4 is then
sharded by contiguous suffix-array ranges; the per-cluster records are gathered on rank 0, which
numbers clusters in SA order and writes the .snp text (cluster numbering is sequential by nature,
/root/reference/ebwt2InDel.cpp:1250, 1328).

The collective helpers take plain tensors so that the host logic is testable with the gloo
backend on CPU (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


_pinned = None
_comm = None            # (world, handle): shared-memory communicator of the position-range traversal, made once per process


def shm_comm(api, rank: int, world: int, group=None):
    """The e2i_comm of this process group: a POSIX shared-memory segment (barrier + publish slots) that the ranks of
    one box share; rank 0 picks the name and the others learn it through torch.distributed."""
    global _comm
    if _comm is not None and _comm[0] == world:
        return _comm[1]
    import os
    import time
    name = [f"/e2i_{os.getpid()}_{int(time.time() * 1e6) & 0xffffffff:x}" if rank == 0 else None]
    dist.broadcast_object_list(name, src=0, group=group)
    h = api.comm_shm(name[0], rank, world)
    dist.barrier(group=group)
    _comm = (world, h)
    return h


class _DeviceWords:
    """Zero-copy view of `words32` int32 words at a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr: int, words32: int):
        self.__cuda_array_interface__ = {"shape": (int(words32),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def wrap_device_words(ptr: int, words32: int, device) -> torch.Tensor:
    return torch.as_tensor(_DeviceWords(ptr, words32), device=device)


TILE = 16384            # e2i_index_slice_align(): slices of the index start on tile boundaries


def index_slices(n: int, world: int):
    """Slices of [0, n) for the slice-wise index build.  The string is cut into tiles of TILE symbols;
    the last tile may hold no symbol at all (it carries the block that makes rank(n) addressable).  Every
    rank owns `per` consecutive tiles -- fewer, possibly none, at the end.  Returns ([(begin, end,
    n_tiles) per rank], per): begin is tile-aligned, an empty rank gets (0, 0, 0), and every tile,
    the last one included, has exactly one owner."""
    tiles = (n // 64 + 1 + 255) // 256
    per = (tiles + world - 1) // world
    out = []
    for r in range(world):
        t_lo, t_hi = min(tiles, r * per), min(tiles, (r + 1) * per)
        if t_hi <= t_lo:
            out.append((0, 0, 0))
        else:
            out.append((min(n, t_lo * TILE), min(n, t_hi * TILE), t_hi - t_lo))
    return out, per


def all_ranks_ok(ok: bool, device, group=None) -> bool:
    """True iff `ok` on every rank: exchanged BEFORE the next collective so that all ranks fail together."""
    t = torch.tensor([0 if ok else 1], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0]) == 0


def build_index_sharded(ctx, bwt_slice, n: int, term: int, rank: int, world: int, group=None):
    """Index of the whole string from one ASCII slice per rank.

    Every rank counts and packs its own tile-aligned slice (bwt_slice = CUDA uint8 tensor holding
    positions index_slices(n, world)[rank]); symbol totals and the superblock table are exchanged,
    then the block ranges are all-gathered over NVLink.  Bit-identical to ctx.index(whole string).
    A forbidden symbol on any rank raises ValueError on every rank."""
    device = bwt_slice.device
    slices, per = index_slices(n, world)
    begin, end, n_tiles = slices[rank]
    assert bwt_slice.numel() == end - begin
    ix = ctx.index_alloc(n, term, tile_multiple=world)
    err = None
    try:
        counts = ix.slice_count(bwt_slice, begin, n_tiles)
    except ValueError as e:            # forbidden symbol in this rank's slice
        err, counts = e, np.zeros(4, dtype=np.uint64)
    if not all_ranks_ok(err is None, device, group):
        raise err if err is not None else ValueError("forbidden symbol in another rank's slice of the input BWT")
    mine = torch.from_numpy(counts.astype(np.int64)).to(device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    allc = torch.stack(allc).cpu().numpy().astype(np.uint64)
    before = allc[:rank].sum(axis=0) if rank else np.zeros(4, dtype=np.uint64)
    sup = torch.from_numpy(ix.slice_super(before).astype(np.int64)).to(device)
    dist.all_reduce(sup, op=dist.ReduceOp.SUM, group=group)
    ix.slice_pack(bwt_slice, before, sup.cpu().numpy().astype(np.uint64))
    ptr, nbytes = ix.device_blocks()
    full = wrap_device_words(ptr, nbytes // 4, device)
    per_words = per * TILE // 2 // 4
    assert per_words * world == full.numel()
    dist.all_gather_into_tensor(full, full[rank * per_words:(rank + 1) * per_words], group=group)
    ix.finish(allc.sum(axis=0))
    return ix


def position_cuts(n: int, world: int):
    """Contiguous suffix-array ranges for phase 4: clusters are assigned by their START position."""
    return [n * r // world for r in range(world + 1)]


def or_reduce_words(tensors, group=None, chunk_words: int = 1 << 28):
    """In-place OR-combine of single-writer bitvectors across ranks (SUM all-reduce of int32 words)."""
    for t in tensors:
        flat = t.view(-1)
        for off in range(0, flat.numel(), chunk_words):
            dist.all_reduce(flat[off:off + chunk_words], op=dist.ReduceOp.SUM, group=group)


def format_sharded(api, recs, left, right, params, two_samples: bool, rank: int, world: int, device, group=None):
    """Every rank formats the records of its own suffix-array slice.  Cluster numbers are sequential
    over the whole run (ebwt2InDel.cpp:1250/1328): each rank first counts the numbers its slice
    consumes (e2i_snp_count, no text), the counts are exchanged, and every rank formats once from
    its true first number.  The text pieces are put together on rank 0 in rank (= SA) order.
    Returns (text on rank 0 as a uint8 numpy array else None, events, clusters_out) over all ranks."""
    mine = torch.tensor([api.snp_count(recs, left, right, params, two_samples)], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    counts = [int(t[0]) for t in allc]
    txt, fst = api.snp_format(recs, left, right, params, two_samples, first_cluster_nr=1 + sum(counts[:rank]), copy=False)
    assert int(fst.clusters_out) == counts[rank]
    ev = torch.tensor([int(fst.events)], dtype=torch.int64, device=device)
    dist.all_reduce(ev, op=dist.ReduceOp.SUM, group=group)
    return gather_bytes(txt.view(), rank, world, device, group), int(ev[0]), sum(counts)


class _DeviceBytes:
    """Zero-copy view of `n` bytes at a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "|u1", "data": (int(ptr), False), "version": 2, "strides": None}


def format_sharded_device(calls, rank: int, world: int, device, group=None):
    """format_sharded for records that stayed in HBM (Context.call_device): the cluster numbers a range consumes are
    counted on the device, exchanged, the text is written on the device from the rank's true first number and gathered
    over NVLink without touching the host; rank 0 copies the whole text out once."""
    mine = torch.tensor([calls.clusters()], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    counts = [int(t[0]) for t in allc]
    ev0 = int(calls.stats.events)
    ptr, ln = calls.snp_device(1 + sum(counts[:rank]))
    try:
        ev = torch.tensor([int(calls.stats.events) - ev0], dtype=torch.int64, device=device)
        dist.all_reduce(ev, op=dist.ReduceOp.SUM, group=group)
        piece = torch.as_tensor(_DeviceBytes(ptr.value, ln), device=device) if ln else torch.empty(0, dtype=torch.uint8, device=device)
        out = gather_bytes(piece, rank, world, device, group)
    finally:
        torch.cuda.synchronize(device)
        calls.free_device(ptr)
    return out, int(ev[0]), sum(counts)


def gather_bytes(view, rank: int, world: int, device, group=None):
    """Concatenation of every rank's bytes on rank 0 (rank order) as a uint8 numpy array.
    Sizes are exchanged first; then ONE gather collective of pieces padded to the largest size (NCCL over
    NVLink on GPUs, gloo on CPU) and one copy per piece into a page-locked host buffer kept across calls."""
    global _pinned
    on_device = isinstance(view, torch.Tensor)          # a piece that is already in device memory
    if not on_device:
        view = memoryview(view).cast("B") if len(view) else memoryview(b"")
    size = torch.tensor([len(view)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s) for s in sizes]
    cuda = torch.device(device).type == "cuda"
    pad = max(max(sizes), 1)
    mine = torch.empty(pad, dtype=torch.uint8, device=device)
    if len(view):
        mine[:len(view)].copy_(view if on_device else torch.from_numpy(np.frombuffer(view, dtype=np.uint8)), non_blocking=True)
    parts = [torch.empty(pad, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0, group=group)
    if rank != 0:
        return None
    total = sum(sizes)
    if _pinned is None or _pinned.numel() < total:
        _pinned = torch.empty(max(total, 1) * 5 // 4, dtype=torch.uint8, pin_memory=cuda)
    out = _pinned[:max(total, 1)]
    off = 0
    for r in range(world):
        if sizes[r]:
            out[off:off + sizes[r]].copy_(parts[r][:sizes[r]], non_blocking=True)
        off += sizes[r]
    if cuda:
        torch.cuda.synchronize(device)
    return out[:total].numpy()


def gather_calls(recs: np.ndarray, left: np.ndarray, right: np.ndarray, rank: int, world: int, group=None):
    """Gather the per-rank call records on rank 0 in rank (= suffix-array) order."""
    out = [None] * world if rank == 0 else None
    dist.gather_object((recs, left, right), out, dst=0, group=group)
    if rank != 0:
        return None
    return (np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out]),
            np.concatenate([o[2] for o in out]))


_SUM_FIELDS = ("leaves", "nodes", "lcp_values", "lcp_values_leaves", "n_min", "da_values", "n_clusters",
               "clust_size", "rank_leaves", "rank_nodes", "rank_call", "bit_updates", "candidates",
               "kernel_launches", "h2d_bytes", "d2h_bytes")
_MAX_FIELDS = ("ms_index", "ms_leaves", "ms_nodes", "ms_call", "ms_h2d", "max_frontier", "levels_leaves", "levels_nodes")


def reduce_stats(st: dict, device, group=None) -> dict:
    """Counters are summed over ranks (every unit of work is done by exactly one rank), phase times
    take the max over ranks."""
    out = dict(st)
    s = torch.tensor([int(st[k]) for k in _SUM_FIELDS] + list(st["clust_sizes"]), dtype=torch.int64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    m = torch.tensor([float(st[k]) for k in _MAX_FIELDS], dtype=torch.float64, device=device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    vals = s.tolist()
    for i, k in enumerate(_SUM_FIELDS):
        out[k] = vals[i]
    out["clust_sizes"] = vals[len(_SUM_FIELDS):]
    for i, k in enumerate(_MAX_FIELDS):
        out[k] = type(st[k])(m[i].item())
    for k in ("ms_format", "ms_wall"):
        out.setdefault(k, 0.0)
    return out


def run_sharded(ctx, api, bwt1, bwt2, da, params, rank: int, world: int, group=None, n1=None, n2=None):
    """Whole path on `world` GPUs.  bwt1 / bwt2 are CUDA uint8 tensors: either the whole eBWT
    (replicated on every rank) or, when n1 / n2 give the total lengths, only this rank's slice
    index_slices(n, world)[rank] -- all a rank needs, since the index is built slice-wise.  The
    document array (mode -d) is passed whole.
    Returns (.snp text on rank 0 else None, reduced stats dict, seconds spent in the exchange)."""
    import time
    device = bwt1.device
    term = params.term
    tm = {}
    tp = time.perf_counter()

    def lap(name):
        nonlocal tp
        torch.cuda.synchronize(device)
        t = time.perf_counter()
        tm[name] = tm.get(name, 0.0) + (t - tp) * 1e3
        tp = t
    def make_index(t, n_total):
        if n_total is not None and n_total != t.numel():
            return build_index_sharded(ctx, t, n_total, term, rank, world, group)      # t is this rank's slice
        n_t = t.numel()
        if world == 1 or n_t < world * 4 * TILE:
            return ctx.index(t, term)
        lo, hi, _ = index_slices(n_t, world)[0][rank]
        return build_index_sharded(ctx, t[lo:hi], n_t, term, rank, world, group)

    b1 = make_index(bwt1, n1)
    b2 = make_index(bwt2, n2) if bwt2 is not None else None
    dabits = ctx.document_array(da) if da is not None else None
    st = api.Stats()
    lap("index")
    import os
    if world > 1 and os.environ.get("E2I_SHARDING", "ranged") != "subtree" and torch.device(device).type == "cuda":
        # position-range sharding: every rank keeps a dense frontier and pulls its records from the peers' frames
        # over NVLink (e2i_navigate_ranged); E2I_SHARDING=subtree selects the independent subtree shards
        lcp, da_nav, st = ctx.navigate_ranged(shm_comm(api, rank, world, group), b1, b2, params, stats=st)
    else:
        lcp, da_nav, st = ctx.navigate(b1, b2, params, shard=rank, n_shards=world, stats=st)
    lap("navigate")
    (pt, wt), (pm, wm) = lcp.device_words()
    words = [wrap_device_words(pt, wt, device), wrap_device_words(pm, wm, device)]
    if da_nav is not None:
        pd, wd = da_nav.device_words()
        words.append(wrap_device_words(pd, wd, device))
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    t0.record()
    or_reduce_words(words, group)
    t1.record()
    torch.cuda.synchronize(device)
    lap("or_reduce")
    n = b1.n + (b2.n if b2 is not None else 0)
    cuts = position_cuts(n, world)
    on_gpu = torch.device(device).type == "cuda"
    if on_gpu:      # the records stay in HBM and are printed there
        calls = ctx.call_device(b1, b2, da_nav if b2 is not None else dabits, lcp, params, cuts[rank], cuts[rank + 1], stats=st)
    else:
        recs, left, right, st = ctx.call(b1, b2, da_nav if b2 is not None else dabits, lcp, params,
                                         cuts[rank], cuts[rank + 1], stats=st, copy=False)
    lap("call")
    stats = reduce_stats(st.as_dict(), device, group)
    lap("reduce_stats")
    # Every LCP position has exactly one writer over all shards ("Computed n/n LCP values", ebwt2InDel.cpp:670);
    # the SUM all-reduce above is an OR only under that condition, so a violated deal fails loudly here.
    if stats["lcp_values"] != n or (b2 is not None and stats["da_values"] != n):
        raise RuntimeError(f"sharded traversal wrote {stats['lcp_values']} LCP values for {n} positions: the shards disagree on the deal")
    if on_gpu:
        snp, events, clusters = format_sharded_device(calls, rank, world, device, group)
        calls.close()
    else:
        snp, events, clusters = format_sharded(api, recs, left, right, params, (b2 is not None or da is not None),
                                               rank, world, device, group)
    lap("format+gather")
    stats["events"], stats["clusters_out"] = events, clusters
    stats["host_ms"] = tm
    return snp, stats, t0.elapsed_time(t1) / 1e3
