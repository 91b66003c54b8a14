"""CPU suite: host-side logic of the multi-GPU driver (ebwt2indel_b200/distributed.py) under the
gloo backend with world_size 2: the OR-combine of single-writer bitvectors as an integer SUM
all-reduce, the phase-4 position cuts, the rank-ordered gather of call records and the counter
reduction.  The CUDA entry points are not called here (no GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ebwt2indel_b200 import api, distributed as dd
        from oracle import binding as ob

        # 1. OR-reduce: split the oracle's bitvectors into two disjoint single-writer halves
        g = load_golden("m1_default")
        o = ob.Bwt(g["bwt1"])
        thr, mn, _ = ob.navigate_one(o, ob.default_params())
        full = [torch.from_numpy(thr.view(np.int32).copy()), torch.from_numpy(mn.view(np.int32).copy())]
        rng = np.random.default_rng(42)          # same mask on both ranks
        parts = []
        for t in full:
            mask = torch.from_numpy(rng.integers(0, 2 ** 31, size=t.numel(), dtype=np.int64).astype(np.int32))
            parts.append((t & mask) if rank == 0 else (t & ~mask))
        dd.or_reduce_words(parts, chunk_words=4096)
        ok_or = all(torch.equal(a, b) for a, b in zip(parts, full))

        # 2. cuts cover [0, n) without overlap
        cuts = dd.position_cuts(1000003, world)
        ok_cuts = cuts[0] == 0 and cuts[-1] == 1000003 and all(a < b for a, b in zip(cuts, cuts[1:]))

        # 3. gather in rank order + format on rank 0 == formatting the concatenation
        p = api.default_params(k_left=8, k_right=6, max_gap=2, complexity=4)

        def fake(r):
            recs = np.zeros(2, dtype=api.CALL_REC_DTYPE)
            left = np.full(2 * 8 * 8, ord("A"), dtype=np.uint8)
            right = np.tile(np.frombuffer(b"GATTAC", dtype=np.uint8), 2).copy()
            for i in range(2):
                a, b = "ACGTACG" + "AC"[r], "ACGTACG" + "GT"[i]
                left[(i * 8 + 0) * 8:(i * 8 + 0) * 8 + 8] = np.frombuffer(a.encode(), dtype=np.uint8)
                left[(i * 8 + 1) * 8:(i * 8 + 1) * 8 + 8] = np.frombuffer(b.encode(), dtype=np.uint8)
                recs[i]["begin"], recs[i]["end"] = 100 * r + 10 * i, 100 * r + 10 * i + 7
                recs[i]["n0"], recs[i]["right_len"], recs[i]["has_right"] = 2, 6, 1
                recs[i]["support"][:2] = (4 + r, 3 + i)
            return recs, left, right

        mine = fake(rank)
        got = dd.gather_calls(*mine, rank, world)
        ok_gather = True
        if rank == 0:
            both = [fake(0), fake(1)]
            want = api.snp_format(np.concatenate([b[0] for b in both]), np.concatenate([b[1] for b in both]),
                                  np.concatenate([b[2] for b in both]), p, two_samples=False)[0]
            have = api.snp_format(got[0], got[1], got[2], p, two_samples=False)[0]
            ok_gather = have == want and have.count(b">cluster:") == 8 and b">cluster:4_id:2" in have
        else:
            ok_gather = got is None

        # 3b. every rank formats its own slice; numbering continues across ranks
        txt, ev, cl = dd.format_sharded(api, *mine, p, False, rank, world, torch.device("cpu"))
        if rank == 0:
            ok_gather = ok_gather and txt.tobytes() == want and cl == 4 and ev == 8
        else:
            ok_gather = ok_gather and txt is None and cl == 4

        # 4. counters: sums and maxima
        st = api.Stats().as_dict()
        st["nodes"], st["ms_nodes"], st["clust_sizes"][3] = 10 + rank, 5.0 + rank, 7
        red = dd.reduce_stats(st, torch.device("cpu"))
        ok_stats = red["nodes"] == 21 and red["ms_nodes"] == 6.0 and red["clust_sizes"][3] == 14
        ret[rank] = (ok_or, ok_cuts, ok_gather, ok_stats)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == (True, True, True, True), (r, ret[r])


def test_slice_and_cut_helpers():
    """index_slices: tile-aligned, every tile (the symbol-free tail tile included) has exactly one owner,
    symbols are covered in order, empty ranks are (0, 0, 0); position_cuts: monotone cover."""
    from ebwt2indel_b200 import distributed as dd
    for n in (1, 127, 128, 16384, 16385, 2 * 16384, 40 * 16384 + 5, 1 << 20, 1000003, 1 << 32, 15_100_000_000):
        for world in (1, 2, 3, 4, 8):
            slices, per = dd.index_slices(n, world)
            tiles = (n // 64 + 1 + 255) // 256
            assert len(slices) == world and per * world >= tiles
            assert sum(s[2] for s in slices) == tiles                      # one owner per tile
            pos, tile = 0, 0
            for lo, hi, nt in slices:
                if nt == 0:
                    assert (lo, hi) == (0, 0)
                    continue
                assert lo % dd.TILE == 0 and lo == min(n, tile * dd.TILE) and lo == pos and lo <= hi <= n
                assert hi - lo <= nt * dd.TILE and nt <= per
                pos, tile = hi, tile + nt
            assert pos == n and tile == tiles
            cuts = dd.position_cuts(n, world)
            assert cuts[0] == 0 and cuts[-1] == n and all(a <= b for a, b in zip(cuts, cuts[1:]))
