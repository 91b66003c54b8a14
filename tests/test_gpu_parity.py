"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the
golden vectors produced by the compiled reference.  Integer / byte / index work: bit-exact."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, resolved_fields

pytestmark = pytest.mark.gpu

COUNTERS = ("leaves", "nodes", "lcp_values", "lcp_values_leaves", "n_min", "da_values", "n_clusters",
            "clust_size", "clusters_out")


def mixed_bwt(n, seed, p_term=0.01):
    rng = np.random.default_rng(seed)
    p = [(1 - p_term) / 4] * 4 + [p_term]
    return np.frombuffer(b"ACGT#", dtype=np.uint8)[rng.choice(5, size=n, p=p)]


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 16383, 16384, 16385, 100000, 1 << 20])
def test_index_rank_access_fl(gpu_ctx, oracle, n):
    bwt = mixed_bwt(n, seed=n)
    ix = gpu_ctx.index(bwt)
    ob = oracle.Bwt(bwt)
    assert ix.n == n
    assert np.array_equal(ix.F(), ob.F())
    rng = np.random.default_rng(7)
    pos = np.unique(np.concatenate([np.arange(0, min(n, 300) + 1), rng.integers(0, n + 1, 4000), [n]])).astype(np.uint64)
    assert np.array_equal(ix.rank4(pos), ob.rank4(pos))
    ipos = pos[pos < n]
    assert np.array_equal(ix.access(ipos), bwt[ipos.astype(np.int64)])
    # FL (dna_bwt.hpp:115-133) on positions whose F symbol is not the terminator
    F = ob.F()
    fpos = ipos[ipos >= F[0]][:1500]
    if len(fpos):
        lib = oracle.lib()
        want = np.array([lib.orc_FL(ob.h, int(i)) for i in fpos], dtype=np.uint64)
        assert np.array_equal(ix.FL(fpos), want)


def test_forbidden_symbol(gpu_ctx):
    bwt = mixed_bwt(50000, 1).copy()
    bwt[31234] = ord("N")
    bwt[40000] = ord("x")
    with pytest.raises(ValueError, match="31234"):
        gpu_ctx.index(bwt)


def test_custom_terminator(gpu_ctx, oracle):
    bwt = mixed_bwt(5000, 2).copy()
    bwt[bwt == ord("#")] = ord("$")
    ix = gpu_ctx.index(bwt, term=ord("$"))
    assert np.array_equal(ix.F(), oracle.Bwt(bwt, term=ord("$")).F())
    with pytest.raises(ValueError):
        gpu_ctx.index(bwt)   # '$' is forbidden when the terminator is '#'


def test_document_array_pack(gpu_ctx):
    rng = np.random.default_rng(5)
    for n in (1, 31, 32, 33, 1000, 100003):
        raw = np.frombuffer(b"01x", dtype=np.uint8)[rng.choice(3, size=n, p=[.5, .45, .05])]
        bits = gpu_ctx.document_array(raw).fetch()
        want = np.packbits((raw == ord("1")).astype(np.uint8), bitorder="little")
        want = np.concatenate([want, np.zeros(-len(want) % 8, dtype=np.uint8)]).view(np.uint64)
        assert np.array_equal(bits, want)


def _case_params(mod, g):
    return mod.default_params(**resolved_fields(g["flags"]))


@pytest.mark.parametrize("name", golden_names())
def test_navigate_bitvectors_match_oracle(gpu_ctx, e2i, oracle, name):
    g = load_golden(name)
    term = resolved_fields(g["flags"]).get("term", ord("#"))
    p, po = _case_params(e2i, g), _case_params(oracle, g)
    b1 = gpu_ctx.index(g["bwt1"], term)
    b2 = gpu_ctx.index(g["bwt2"], term) if g["bwt2"] is not None else None
    lcp, da, st = gpu_ctx.navigate(b1, b2, p)
    n = len(g["bwt1"]) + (len(g["bwt2"]) if b2 else 0)
    thr, mn = lcp.fetch(n)
    o1 = oracle.Bwt(g["bwt1"], term)
    if b2:
        o2 = oracle.Bwt(g["bwt2"], term)
        othr, omn, oda, ost = oracle.navigate_two(o1, o2, po)
        assert np.array_equal(da.fetch(), oda[:(n + 63) // 64])
    else:
        othr, omn, ost = oracle.navigate_one(o1, po)
    assert np.array_equal(thr, othr[:len(thr)])
    assert np.array_equal(mn, omn[:len(mn)])
    for k in ("leaves", "nodes", "lcp_values", "lcp_values_leaves", "n_min", "da_values", "rank_leaves", "rank_nodes"):
        assert getattr(st, k) == getattr(ost, k), k
    for k, v in g["counters"].items():
        if k != "n_clusters":
            assert getattr(st, k) == v, f"{k} differs from the reference's printed counter"


@pytest.mark.parametrize("name", golden_names())
def test_whole_path_matches_reference_golden(gpu_ctx, e2i, oracle, name):
    g = load_golden(name)
    p = _case_params(e2i, g)
    snp, st = gpu_ctx.run(g["bwt1"], g["bwt2"], g["da"], p)
    assert snp == g["snp"], f"{name}: .snp differs from the compiled reference's"
    for k, v in g["counters"].items():
        assert getattr(st, k) == v, k
    osnp, ost = oracle.run(g["bwt1"], g["bwt2"], g["da"], _case_params(oracle, g))
    assert snp == osnp
    for k in COUNTERS:
        assert getattr(st, k) == getattr(ost, k), k
    assert list(st.clust_sizes) == list(ost.clust_sizes)
    if g["bwt2"] is None and g["da"] is None:
        assert st.events == ost.events


def test_whole_path_device_inputs(gpu_ctx, e2i):
    import torch
    g = load_golden("m3_default")
    d1 = torch.from_numpy(g["bwt1"].copy()).cuda()
    dd = torch.from_numpy(g["da"].copy()).cuda()
    snp, _ = gpu_ctx.run(d1, None, dd, e2i.default_params())
    assert snp == g["snp"]


@pytest.mark.parametrize("seed,mode", [(101, 1), (102, 1), (103, 3), (104, 2)])
def test_seeded_mid_size_vs_oracle(gpu_ctx, e2i, oracle, seed, mode):
    """Fresh seeded inputs at a size the oracle finishes in seconds (n ~ 2-4 M)."""
    from ebwt2indel_b200 import synth
    if mode == 1:
        reads = synth.diploid_reads(60000, 120, 30, 20, 100, seed=seed)
        bwt, _ = synth.ebwt_bcr_numpy(reads)
        args = (bwt, None, None)
    else:
        r0, r1 = synth.two_individuals_reads(20000, 40, 10, 20, 100, seed=seed)
        if mode == 3:
            args = synth.merged_ebwt_da(r0, r1, builder=synth.ebwt_bcr_numpy)
            args = (args[0], None, args[1])
        else:
            args = (synth.ebwt_bcr_numpy(r0)[0], synth.ebwt_bcr_numpy(r1)[0], None)
    snp, st = gpu_ctx.run(*args, e2i.default_params())
    osnp, ost = oracle.run(*args, oracle.default_params())
    assert snp == osnp
    for k in COUNTERS:
        assert getattr(st, k) == getattr(ost, k), k
    assert len(snp) > 0


def test_repetitive_and_degenerate_inputs(gpu_ctx, e2i, oracle):
    """Edge cases: a single read, identical reads (deep unary paths, large leaves), homopolymers."""
    from ebwt2indel_b200 import synth
    cases = []
    one = np.frombuffer(b"ACGTACGTTGCA", dtype=np.uint8)[None, :]
    cases.append(one)
    cases.append(np.repeat(one, 40, axis=0))
    cases.append(np.full((25, 30), ord("A"), dtype=np.uint8))
    rng = np.random.default_rng(9)
    rep = synth.BASES[rng.integers(0, 4, 50)]
    tandem = np.tile(rep, 20)
    idx = np.arange(60)
    cases.append(tandem[rng.integers(0, len(tandem) - 60, 300)[:, None] + idx[None, :]])
    for reads in cases:
        bwt, _ = synth.ebwt_naive(reads)
        for kw in ({}, {"K": 4, "k_left": 8, "k_right": 6, "mcov_out": 2, "complexity": 5, "max_gap": 3}):
            snp, st = gpu_ctx.run(bwt, None, None, e2i.default_params(**kw))
            osnp, ost = oracle.run(bwt, None, None, oracle.default_params(**kw))
            assert snp == osnp
            for k in COUNTERS:
                assert getattr(st, k) == getattr(ost, k), (k, reads.shape, kw)


def test_sharded_navigation_ors_to_full(gpu_ctx, e2i):
    """the bitvectors of the traversal shards OR-combine to the single-shard result,
    every bit has one writer (so the integer sum of the words equals their OR)."""
    g = load_golden("m1_default")
    p = e2i.default_params()
    b1 = gpu_ctx.index(g["bwt1"])
    n = len(g["bwt1"])
    full, _, fst = gpu_ctx.navigate(b1, None, p)
    fthr, fmn = full.fetch(n)
    for n_shards in (2, 3, 8):
        sthr = np.zeros_like(fthr)
        smn = np.zeros_like(fmn)
        tot = {k: 0 for k in ("leaves", "nodes", "lcp_values", "n_min")}
        for s in range(n_shards):
            l, _, st = gpu_ctx.navigate(b1, None, p, shard=s, n_shards=n_shards)
            t, m = l.fetch(n)
            assert not np.any(sthr & t) and not np.any(smn & m)
            sthr |= t
            smn |= m
            for k in tot:
                tot[k] += getattr(st, k)
        assert np.array_equal(sthr, fthr) and np.array_equal(smn, fmn)
        for k in tot:
            assert tot[k] == getattr(fst, k), (k, n_shards)


def test_sharded_calls_concatenate(gpu_ctx, e2i):
    g = load_golden("m3_default")
    p = e2i.default_params()
    b1 = gpu_ctx.index(g["bwt1"])
    da = gpu_ctx.document_array(g["da"])
    lcp, _, _ = gpu_ctx.navigate(b1, None, p)
    n = len(g["bwt1"])
    cuts = [0, n // 3 + 17, 2 * n // 3 + 5, n]
    parts = [gpu_ctx.call(b1, None, da, lcp, p, cuts[i], cuts[i + 1]) for i in range(3)]
    recs = np.concatenate([x[0] for x in parts])
    left = np.concatenate([x[1] for x in parts])
    right = np.concatenate([x[2] for x in parts])
    snp, _ = e2i.snp_format(recs, left, right, p, two_samples=True)
    assert snp == g["snp"]
    assert sum(x[3].n_clusters for x in parts) == g["counters"]["n_clusters"]


def test_small_frontier_budget_gives_same_result(e2i):
    """Bounded-memory traversal: a tiny frontier budget forces chunked depth-first sweeps."""
    g = load_golden("m1_default")
    ctx = e2i.Context(0, frontier_bytes=24 << 20)
    try:
        snp, st = ctx.run(g["bwt1"], None, None, e2i.default_params())
    finally:
        ctx.close()
    assert snp == g["snp"]
    assert st.nodes == g["counters"]["nodes"]


def _run_cli(tmp_path, g, env=None):
    exe = os.path.join(ROOT, "bin", "ebwt2InDel")
    f1 = tmp_path / "a.ebwt"
    g["bwt1"].tofile(f1)
    cmd, names = [exe, "-1", str(f1)], {str(f1): "<bwt1>"}
    if g["bwt2"] is not None:
        f2 = tmp_path / "b.ebwt"
        g["bwt2"].tofile(f2)
        cmd += ["-2", str(f2)]
        names[str(f2)] = "<bwt2>"
    if g["da"] is not None:
        f3 = tmp_path / "da.txt"
        g["da"].tofile(f3)
        cmd += ["-d", str(f3)]
        names[str(f3)] = "<da>"
    out = tmp_path / "out.snp"
    names[str(out)] = "<out>"
    inv = {v: k for k, v in {"-L": "k_left", "-R": "k_right", "-k": "K", "-g": "max_gap", "-v": "max_snvs",
                             "-m": "mcov_out", "-c": "complexity", "-q": "max_variants_per_position",
                             "-t": "term"}.items()}
    for k, v in g["flags"].items():
        cmd += [inv[k], str(v)]
    r = subprocess.run(cmd + ["-o", str(out)], capture_output=True, text=True, env=env)
    text = r.stdout
    for path, ph in names.items():
        text = text.replace(path, ph)
    return r, text, out


def test_cli_drop_in(tmp_path):
    """bin/ebwt2InDel with the reference's argv: same .snp bytes, and the same stdout as the compiled reference
    line for line (tests/golden/<case>.stdout.txt) -- banner, counters, averages, the cluster-length histogram --
    except the reference's progress percentages and its 'Max stack depth' lines (a property of its DFS order)."""
    for name in golden_names():
        g = load_golden(name)
        r, text, out = _run_cli(tmp_path, g)
        assert r.returncode == 0, r.stdout + r.stderr
        assert out.read_bytes() == g["snp"], name
        want = [ln for ln in open(os.path.join(ROOT, "tests", "golden", name + ".stdout.txt")).read().split("\n")
                if not ln.startswith("Max stack depth")]
        assert text.split("\n") == want, name
    bad = tmp_path / "bad.ebwt"
    bad.write_bytes(b"ACGTNACGT#")
    exe = os.path.join(ROOT, "bin", "ebwt2InDel")
    r = subprocess.run([exe, "-1", str(bad), "-o", str(tmp_path / "x.snp")], capture_output=True, text=True)
    assert r.returncode == 1 and "read forbidden character 'N' (ASCII code 78)" in r.stdout


def test_streaming_ingest_and_index_sidecar(gpu_ctx, e2i, oracle, tmp_path):
    """f1: e2i_index_build_file (reader thread -> page-locked ring -> copy stream -> counting pass) gives the
    index of the in-memory build, for sizes around the 64 MB chunk and the tile size; the packed sidecar
    (save / load) answers the same queries; e2i_run_files equals e2i_run; a short DA file repeats its last
    byte (ebwt2InDel.cpp:1503-1508); the CLI reuses the sidecar when E2I_INDEX_CACHE=1."""
    rng = np.random.default_rng(3)
    for n in (1, 16384, (64 << 20) - 1, (64 << 20) + 16385, 3 * (64 << 20) + 7):
        bwt = mixed_bwt(n, seed=n % 1000)
        f = tmp_path / "x.ebwt"
        bwt.tofile(f)
        a, b = gpu_ctx.index_file(str(f)), gpu_ctx.index(bwt)
        pos = np.unique(np.concatenate([rng.integers(0, n + 1, 3000), [0, n]])).astype(np.uint64)
        assert np.array_equal(a.F(), b.F()) and np.array_equal(a.rank4(pos), b.rank4(pos))
        a.save(str(tmp_path / "x.e2ix"))
        c = gpu_ctx.index_load(str(tmp_path / "x.e2ix"))
        assert c.n == n and np.array_equal(c.F(), b.F()) and np.array_equal(c.rank4(pos), b.rank4(pos))
        del a, b, c
    bad = mixed_bwt(100000, 1).copy()
    bad[70001] = ord("N")
    bad.tofile(tmp_path / "bad.ebwt")
    with pytest.raises(ValueError, match="70001"):
        gpu_ctx.index_file(str(tmp_path / "bad.ebwt"))
    for name in ("m1_default", "m2_flags", "m3_default"):
        g = load_golden(name)
        paths = {}
        for key in ("bwt1", "bwt2", "da"):
            if g[key] is not None:
                paths[key] = str(tmp_path / (key + ".bin"))
                g[key].tofile(paths[key])
        snp, st = gpu_ctx.run_files(paths["bwt1"], paths.get("bwt2"), paths.get("da"), _case_params(e2i, g))
        assert snp == g["snp"], name
    # short document array: the missing positions take the value of the last byte
    g = load_golden("m3_default")
    n = len(g["bwt1"])
    cut = n - 1000
    da_short = g["da"][:cut]
    da_short.tofile(tmp_path / "short.txt")
    want = np.concatenate([da_short, np.full(n - cut, da_short[-1], dtype=np.uint8)])
    bits = gpu_ctx.document_array_file(str(tmp_path / "short.txt"), n).fetch()
    assert np.array_equal(bits, gpu_ctx.document_array(want).fetch())
    # CLI with the sidecar cache: first run writes it, second run loads it; same output
    g = load_golden("m1_default")
    env = dict(os.environ, E2I_INDEX_CACHE="1")
    for _ in range(2):
        r, _, out = _run_cli(tmp_path, g, env)
        assert r.returncode == 0 and out.read_bytes() == g["snp"]
        assert os.path.exists(str(tmp_path / "a.ebwt") + ".e2ix")


def test_gap_longer_than_left_context_is_accepted(gpu_ctx, e2i, oracle):
    """-g > -L: accepted like the reference (its substr wraps, ebwt2InDel.cpp:208-219); same text as the oracle."""
    g = load_golden("m3_default")
    kw = {"k_left": 12, "max_gap": 20, "k_right": 10, "K": 8, "complexity": 6}
    snp, _ = gpu_ctx.run(g["bwt1"], None, g["da"], e2i.default_params(**kw))
    osnp, _ = oracle.run(g["bwt1"], None, g["da"], oracle.default_params(**kw))
    assert snp == osnp and len(snp) > 0


def test_gpu_ebwt_builder_matches_reference_builders(gpu_ctx):
    """Tooling check: the GPU BCR builder (product index + rank kernels for the LF step, csrc/tools.cu
    for the merge) gives the eBWT / DA of the naive suffix sort."""
    import torch
    from ebwt2indel_b200 import synth
    plan = synth.diploid_plan(3000, 8, 3, 20, 50, seed=31)
    want, _ = synth.ebwt_naive(plan.materialize())
    got = synth.ebwt_bcr_gpu(gpu_ctx, [plan], "cuda:0")
    assert np.array_equal(got.cpu().numpy(), want)
    p0, p1 = synth.two_individuals_plans(2000, 6, 2, 16, 40, seed=32)
    m, da = synth.merged_ebwt_da(p0.materialize(), p1.materialize())
    bwt, owner = synth.ebwt_bcr_gpu(gpu_ctx, [p0, p1], "cuda:0", want_owner=True)
    assert np.array_equal(bwt.cpu().numpy(), m)
    assert np.array_equal((owner + 48).cpu().numpy(), da)
    torch.cuda.synchronize()


@pytest.mark.parametrize("n,world", [(200000, 2), (1 << 20, 3), (1000003, 8), (40 * 16384 + 5, 8), (4 * 65536, 2)])
def test_slicewise_index_build_equals_whole_build(gpu_ctx, oracle, n, world):
    """Multi-GPU index construction, emulated on one GPU: every 'rank' counts and packs its own
    tile-aligned slice, the block ranges are put together, and the result answers rank queries
    exactly like the index built from the whole string (and like the oracle)."""
    import torch
    from ebwt2indel_b200 import distributed as dd
    bwt = mixed_bwt(n, seed=n + world)
    dev = torch.from_numpy(bwt.copy()).cuda()
    slices, per = dd.index_slices(n, world)
    parts = [gpu_ctx.index_alloc(n, ord("#"), tile_multiple=world) for _ in range(world)]
    counts = np.stack([parts[r].slice_count(dev[slices[r][0]:slices[r][1]], slices[r][0], slices[r][2]) for r in range(world)])
    sup = np.zeros(parts[0].n_super * 4, dtype=np.uint64)
    for r in range(world):
        sup += parts[r].slice_super(counts[:r].sum(axis=0))
    views = []
    for r in range(world):
        parts[r].slice_pack(dev[slices[r][0]:slices[r][1]], counts[:r].sum(axis=0), sup)
        ptr, nbytes = parts[r].device_blocks()
        views.append(dd.wrap_device_words(ptr, nbytes // 4, dev.device))
    w = per * dd.TILE // 2 // 4
    for r in range(1, world):
        views[0][r * w:(r + 1) * w].copy_(views[r][r * w:(r + 1) * w])   # the all-gather, by hand
    torch.cuda.synchronize()
    parts[0].finish(counts.sum(axis=0))
    whole = gpu_ctx.index(bwt)
    assert np.array_equal(parts[0].F(), whole.F())
    rng = np.random.default_rng(1)
    cuts = [s[0] for s in slices] + [s[1] for s in slices]
    pos = np.unique(np.concatenate([rng.integers(0, n + 1, 5000), [0, n], np.clip(np.array(cuts), 0, n),
                                    np.clip(np.array(cuts) + 1, 0, n), np.clip(np.array(cuts) - 1, 0, n)])).astype(np.uint64)
    assert np.array_equal(parts[0].rank4(pos), whole.rank4(pos))
    assert np.array_equal(parts[0].rank4(pos), oracle.Bwt(bwt).rank4(pos))


def _bench_workload(gpu_ctx, name, scale=1.0):
    import torch
    import bench
    cfg = bench.CONFIGS[name]
    if scale != 1.0:
        cfg = bench.scaled(cfg, scale)
    wl = bench.make_workload(cfg, torch.device("cuda:0"), gpu_ctx)
    torch.cuda.synchronize()
    return wl


def test_full_size_c1_matches_oracle(gpu_ctx, e2i, oracle):
    """BASELINE.json configs[0] at full size (n = 40.4 M): .snp bytes and every counter against the oracle."""
    wl = _bench_workload(gpu_ctx, "C1")
    assert wl["n"] == 40_400_000
    snp, st = gpu_ctx.run(wl["bwt1"], None, None, e2i.default_params())
    osnp, ost = oracle.run(wl["bwt1"].cpu().numpy(), None, None, oracle.default_params())
    assert snp == osnp and len(snp) > 100_000
    for k in COUNTERS + ("events", "rank_leaves", "rank_nodes"):
        assert getattr(st, k) == getattr(ost, k), k
    assert st.lcp_values == wl["n"]


@pytest.mark.parametrize("name,scale", [("C4", 1 / 16), ("C2", 1 / 4), ("C3", 1 / 8)])
def test_large_inputs_size_independent_properties(gpu_ctx, e2i, name, scale):
    """Sizes the oracle cannot finish in seconds (0.25 - 0.95 G symbols, shapes of configs 2-4):
    properties that do not need it.  (1) every LCP position is computed exactly once -- the
    reference's own 'Computed n/n LCP values' invariant; (2) every DA value too in mode -2;
    (3) traversal shards write disjoint bits whose sum is the unsharded bitvector, word for word;
    (4) the text does not depend on the frontier budget (chunked depth-first traversal), on where
    the inputs live (host / device) or on the run (idempotence); (5) clusters are reported in
    increasing cluster number."""
    import re
    import torch
    from ebwt2indel_b200 import distributed as dd
    wl = _bench_workload(gpu_ctx, name, scale)
    p = e2i.default_params()
    n = wl["n"]
    snp, st = gpu_ctx.run(wl["bwt1"], wl["bwt2"], wl["da"], p)
    assert st.lcp_values == n                                              # (1)
    if wl["bwt2"] is not None:
        assert st.da_values == n                                           # (2)
    assert st.nodes > n // 2 and st.n_clusters > 0 and len(snp) > 0
    # (3)
    b1 = gpu_ctx.index(wl["bwt1"])
    b2 = gpu_ctx.index(wl["bwt2"]) if wl["bwt2"] is not None else None
    full, fda, fst = gpu_ctx.navigate(b1, b2, p)
    (pt, wt), (pm, wm) = full.device_words()
    fthr = dd.wrap_device_words(pt, wt, "cuda:0").clone()
    fmin = dd.wrap_device_words(pm, wm, "cuda:0").clone()
    torch.cuda.synchronize()
    del full, fda
    sthr, smin, nodes = torch.zeros_like(fthr), torch.zeros_like(fmin), 0
    for s in range(3):
        part, pda, pst = gpu_ctx.navigate(b1, b2, p, shard=s, n_shards=3)
        (pt, wt), (pm, wm) = part.device_words()
        t, m = dd.wrap_device_words(pt, wt, "cuda:0"), dd.wrap_device_words(pm, wm, "cuda:0")
        assert not bool((sthr & t).any()) and not bool((smin & m).any())
        sthr += t
        smin += m
        nodes += pst.nodes
        torch.cuda.synchronize()          # torch reads the library's buffers on its own stream: finish before they are freed
        del part, pda
    assert torch.equal(sthr, fthr) and torch.equal(smin, fmin) and nodes == fst.nodes
    del b1, b2, sthr, smin, fthr, fmin
    # (4)
    host = [None if t is None else t.cpu().numpy() for t in (wl["bwt1"], wl["bwt2"], wl["da"])]
    assert gpu_ctx.run(*host, p)[0] == snp
    small = e2i.Context(0, frontier_bytes=max(32 << 20, n // 3))   # far below one level's frames: forces chunked sweeps
    try:
        s2, st2 = small.run(wl["bwt1"], wl["bwt2"], wl["da"], p)
    finally:
        small.close()
    assert s2 == snp and st2.levels_nodes > st.levels_nodes, (st2.levels_nodes, st.levels_nodes)   # really chunked
    # (5)
    nums = [int(x) for x in re.findall(rb">cluster:(\d+)_", snp)]
    assert nums == sorted(nums) and nums[0] == 1 and nums[-1] == st.clusters_out


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0]])
def test_multi_gpu_single_process_matches(e2i, oracle, devices):
    """e2i_run_multi (what bin/ebwt2InDel runs with E2I_GPUS / E2I_DEVICES): one process, one thread and one
    context per rank, slice-wise index + peer copies, sharded traversal, OR-combine kernel over peer pointers,
    phase 4 per suffix-array range.  Ranks on one device here (the logic does not care): the text and the
    counters equal the reference's in all three modes, and a mid-size seeded input equals the oracle."""
    from ebwt2indel_b200 import synth
    for name in ("m1_default", "m1_flags", "m2_default", "m3_default", "m2_flags", "m1_short_reads"):
        g = load_golden(name)
        snp, st = e2i.run_multi(devices, g["bwt1"], g["bwt2"], g["da"], _case_params(e2i, g))
        assert snp == g["snp"], name
        for k, v in g["counters"].items():
            assert getattr(st, k) == v, (name, k)
    reads = synth.diploid_reads(150000, 300, 60, 20, 100, seed=77)          # n = 6 M: the shards get real work
    bwt, _ = synth.ebwt_bcr_numpy(reads)
    snp, st = e2i.run_multi(devices, bwt, None, None, e2i.default_params())
    osnp, ost = oracle.run(bwt, None, None, oracle.default_params())
    assert snp == osnp and len(snp) > 0
    for k in COUNTERS + ("events",):
        assert getattr(st, k) == getattr(ost, k), k


def test_cli_multi_gpu_env(tmp_path):
    """bin/ebwt2InDel with E2I_DEVICES=0,0: same .snp bytes as the reference."""
    exe = os.path.join(ROOT, "bin", "ebwt2InDel")
    g = load_golden("m3_default")
    f1, f3, out = tmp_path / "a.ebwt", tmp_path / "da.txt", tmp_path / "o.snp"
    g["bwt1"].tofile(f1)
    g["da"].tofile(f3)
    r = subprocess.run([exe, "-1", str(f1), "-d", str(f3), "-o", str(out)], capture_output=True, text=True,
                       env=dict(os.environ, E2I_DEVICES="0,0"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert out.read_bytes() == g["snp"]


def test_product_ebwt_builder(gpu_ctx, e2i, tmp_path):
    """f2: e2i_ebwt_build / bin/ebwt_build (GPU BCR on the product's own index + rank kernels) give the eBWT and
    the document array of the naive suffix sort, and ebwt2InDel on their output equals the golden .snp."""
    from ebwt2indel_b200 import synth
    reads = synth.diploid_reads(3000, 8, 3, 20, 50, seed=31)
    want, _ = synth.ebwt_naive(reads)
    assert np.array_equal(gpu_ctx.ebwt_build(reads), want)
    r0, r1 = synth.two_individuals_reads(2000, 6, 2, 16, 40, seed=32)
    m, da = synth.merged_ebwt_da(r0, r1)
    bwt, got_da = gpu_ctx.ebwt_build(np.concatenate([r0, r1]), second_from=len(r0))
    assert np.array_equal(bwt, m) and np.array_equal(got_da, da)
    with pytest.raises(ValueError, match="read 3 at offset 7"):
        bad = reads.copy()
        bad[3, 7] = ord("N")
        gpu_ctx.ebwt_build(bad)
    # command line: FASTA in, -r adds the reverse complements; then the whole chain reads -> eBWT -> .snp
    fwd = synth.read_plan([synth.random_genome(4000, np.random.default_rng(5))], 1200, 60, np.random.default_rng(6), revcomp=False).materialize()
    fa = tmp_path / "reads.fa"
    with open(fa, "w") as f:
        for i, r in enumerate(fwd):
            f.write(f">r{i}\n{r.tobytes().decode()}\n")
    exe = os.path.join(ROOT, "bin", "ebwt_build")
    out = tmp_path / "reads.ebwt"
    r = subprocess.run([exe, "-i", str(fa), "-r", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    both = np.concatenate([fwd, synth._COMP[fwd[:, ::-1]]])
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), synth.ebwt_naive(both)[0])
    g = load_golden("m1_default")          # its reads are synth.diploid_reads(6000, 14, 5, 24, 100, seed=11)
    gr = synth.diploid_reads(6000, 14, 5, 24, 100, seed=11)
    assert np.array_equal(gpu_ctx.ebwt_build(gr), g["bwt1"])


def test_multi_gpu_ranged_node_pass(e2i, oracle, monkeypatch):
    """E2I_RANGED_NODES=1: the internal-node pass position-range sharded as well (every rank pulls its records out
    of the peers' compacted frames); same text and counters as the reference, ranks emulated on one device."""
    monkeypatch.setenv("E2I_RANGED_NODES", "1")
    for name in ("m1_default", "m2_default", "m3_flags"):
        g = load_golden(name)
        snp, st = e2i.run_multi([0, 0, 0], g["bwt1"], g["bwt2"], g["da"], _case_params(e2i, g))
        assert snp == g["snp"], name
        for k, v in g["counters"].items():
            assert getattr(st, k) == v, (name, k)


# ---- device formatter (e2i_snp_format_gpu / e2i_call_snp) against the host formatter ---------------
def _random_call_records(rng, n, kl, kr, alphabet, dtype):
    """Arbitrary records in the layout of e2i_call_rec: everything the formatter branches on varies."""
    recs = np.zeros(n, dtype=dtype)
    recs["begin"] = np.arange(n) * 7
    recs["end"] = recs["begin"] + 3
    recs["n0"] = rng.integers(0, 5, n)
    recs["n1"] = rng.integers(0, 5, n)
    recs["right_len"] = rng.integers(0, kr + 1, n)
    recs["has_right"] = rng.random(n) < 0.9
    recs["support"] = rng.integers(0, 12, (n, 8))
    letters = np.frombuffer(alphabet, dtype=np.uint8)
    # contexts that differ in few places, so that SNPs, insertions and deletions all occur
    base = letters[rng.integers(0, len(letters), (n, 1, kl + 12))]
    left = np.repeat(base, 8, axis=1)
    shift = rng.integers(0, 4, (n, 8))
    out = np.zeros((n, 8, kl), dtype=np.uint8)
    for s in range(4):
        m = shift == s
        out[m] = left[:, :, s:s + kl][m]
    flip = rng.random((n, 8, kl)) < 0.04
    out[flip] = letters[rng.integers(0, len(letters), int(flip.sum()))]
    out[:, :, kl - 1] = letters[rng.integers(0, min(4, len(letters)), (n, 8))]      # the variant character
    right = letters[rng.integers(0, len(letters), (n, kr))]
    runs = rng.random(n) < 0.1
    right[runs] = right[runs][:, :1]                                                    # low-complexity right contexts
    return recs, np.ascontiguousarray(out).reshape(-1), np.ascontiguousarray(right).reshape(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("kl,kr,gap,alphabet", [(31, 30, 10, b"ACGT"), (8, 5, 3, b"ACGT"), (40, 30, 10, b"ACGT"), (31, 30, 0, b"ACGT"),
                                                (5, 4, 9, b"ACGT"), (31, 30, 10, b"ACGTN#"), (32, 255, 33, b"ACGT"), (1, 1, 1, b"ACGT")])
@pytest.mark.parametrize("two", [False, True])
def test_device_formatter_equals_host_formatter(gpu_ctx, e2i, kl, kr, gap, alphabet, two):
    rng = np.random.default_rng(kl * 1000 + kr + gap + (7 if two else 0))
    p = e2i.default_params()
    p.k_left, p.k_right, p.max_gap, p.complexity = kl, kr, gap, min(20, max(1, kr // 2))
    p.max_snvs = 3
    recs, left, right = _random_call_records(rng, 20000, kl, kr, alphabet, e2i.CALL_REC_DTYPE)
    for first in (1, 95, 99990):                        # cluster numbers that cross a power of ten change the layout
        host, sh = e2i.snp_format(recs, left, right, p, two_samples=two, first_cluster_nr=first)
        dev, sd = gpu_ctx.snp_format(recs, left, right, p, two_samples=two, first_cluster_nr=first)
        assert len(host) > 1000
        assert dev == host
        assert (sd.events, sd.clusters_out) == (sh.events, sh.clusters_out)
        assert e2i.snp_count(recs, left, right, p, two) == sd.clusters_out
    # nothing to print
    dev, sd = gpu_ctx.snp_format(recs[:0], left[:0], right[:0], p, two_samples=two)
    assert dev == b"" and sd.clusters_out == 0
    none = recs.copy()
    none["has_right"] = 0
    dev, sd = gpu_ctx.snp_format(none, left, right, p, two_samples=two)
    assert dev == b"" and sd.clusters_out == 0 and sd.events == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["m1_default", "m2_default", "m3_default"])
def test_call_snp_ranges_concatenate_to_golden(gpu_ctx, e2i, name):
    """Phase 4 + text on the device, range by range with running cluster numbers (what a rank of a sharded run does)."""
    g = load_golden(name)
    p = e2i.default_params()
    b1 = gpu_ctx.index(g["bwt1"])
    b2 = gpu_ctx.index(g["bwt2"]) if g.get("bwt2") is not None else None
    da = gpu_ctx.document_array(g["da"]) if g.get("da") is not None else None
    lcp, da_nav, _ = gpu_ctx.navigate(b1, b2, p)
    n = len(g["bwt1"]) + (len(g["bwt2"]) if b2 else 0)
    whole, st = gpu_ctx.call_snp(b1, b2, da_nav if b2 else da, lcp, p)
    assert whole == g["snp"]
    assert st.n_clusters == g["counters"]["n_clusters"]
    cuts = [0, n // 3 + 17, 2 * n // 3 + 5, n]
    text, first = b"", 1
    for i in range(3):
        part, ps = gpu_ctx.call_snp(b1, b2, da_nav if b2 else da, lcp, p, cuts[i], cuts[i + 1], first_cluster_nr=first)
        text += part
        first += ps.clusters_out
    assert text == g["snp"]


@pytest.mark.parametrize("name", ["m1_default", "m2_default", "m3_default"])
def test_device_resident_calls_number_ranges_consecutively(gpu_ctx, e2i, name):
    """e2i_call_device: the records stay in HBM; counted there, then printed from each range's true first cluster
    number (the protocol of the multi-GPU drivers).  The host copy of the text and the device copy agree."""
    import torch
    g = load_golden(name)
    p = e2i.default_params()
    b1 = gpu_ctx.index(g["bwt1"])
    b2 = gpu_ctx.index(g["bwt2"]) if g.get("bwt2") is not None else None
    da = gpu_ctx.document_array(g["da"]) if g.get("da") is not None else None
    lcp, da_nav, _ = gpu_ctx.navigate(b1, b2, p)
    n = len(g["bwt1"]) + (len(g["bwt2"]) if b2 else 0)
    cuts = [0, n // 4 + 3, n // 4 + 3, 2 * n // 3 + 5, n]                 # one empty range
    parts = [gpu_ctx.call_device(b1, b2, da_nav if b2 else da, lcp, p, cuts[i], cuts[i + 1]) for i in range(4)]
    counts = [c.clusters() for c in parts]
    host_view = gpu_ctx.call(b1, b2, da_nav if b2 else da, lcp, p)
    assert sum(len(c) for c in parts) == len(host_view[0])
    assert sum(counts) == e2i.snp_count(host_view[0], host_view[1], host_view[2], p, b2 is not None or da is not None)
    text = b""
    for i, c in enumerate(parts):
        first = 1 + sum(counts[:i])
        piece = c.snp(first)
        ptr, ln = c.snp_device(first)
        try:
            assert ln == len(piece)
            if ln:
                from ebwt2indel_b200.distributed import _DeviceBytes
                dev = torch.as_tensor(_DeviceBytes(ptr.value, ln), device="cuda:0")
                assert bytes(dev.cpu().numpy()) == piece
        finally:
            c.free_device(ptr)
        text += piece
    assert text == g["snp"]
    with pytest.raises(e2i.E2iError):
        lib_view = e2i.lib().e2i_calls_view          # the records of a device handle cannot be viewed on the host
        import ctypes as C
        pr, pl, pt, nn = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        e2i._check(lib_view(parts[0].h, C.byref(pr), C.byref(pl), C.byref(pt), C.byref(nn))) if len(parts[0]) else e2i._check(3)
