"""GPU suite, large cases: the full-size BASELINE.json configurations and inputs with n > 2^32 against the
COMPILED, UNMODIFIED reference.  tests/golden/big/<case>.json holds what the reference produced (sha-256 of
its .snp, every printed counter) when tests/golden/make_big_golden.py ran it on the same seeded input in the
build container (0.5 - 3 h of one core per case); see tests/bigcase.py.  Here the input is rebuilt on the GPU,
proven to be the same string (checksums), and the product's output is compared bit for bit."""
import hashlib

import pytest

import bigcase

pytestmark = pytest.mark.gpu

COUNTERS = ("leaves", "nodes", "lcp_values", "lcp_values_leaves", "n_min", "da_values", "n_clusters")


@pytest.mark.parametrize("name", sorted(bigcase.CASES))
def test_big_case_matches_compiled_reference(gpu_ctx, e2i, name):
    import torch
    from ebwt2indel_b200 import workloads
    g = bigcase.load_golden(name)
    if g is None:
        pytest.skip(f"tests/golden/big/{name}.json not generated (tests/golden/make_big_golden.py)")
    free_b, _ = torch.cuda.mem_get_info()
    n = g["n1"] + g.get("n2", 0)
    if free_b < 3.2 * n + (4 << 30):
        pytest.skip("not enough free device memory for this case")
    dev = torch.device("cuda:0")
    wl = workloads.make_workload_gpu(bigcase.case_config(name), dev, gpu_ctx)
    torch.cuda.synchronize()
    gpu_ctx.trim()
    # the GPU-built input is the string the reference was run on
    assert wl["bwt1"].numel() == g["n1"]
    assert workloads.checksum(wl["bwt1"]) == g["bwt1_checksum"]
    if "n2" in g:
        assert wl["bwt2"].numel() == g["n2"] and workloads.checksum(wl["bwt2"]) == g["bwt2_checksum"]
    if "da_checksum" in g:
        assert workloads.checksum(wl["da"]) == g["da_checksum"]
    snp, st = gpu_ctx.run(wl["bwt1"], wl["bwt2"], wl["da"], e2i.default_params(), copy=False)
    del wl
    torch.cuda.empty_cache()
    for k in COUNTERS:
        if k in g["counters"]:
            assert getattr(st, k) == g["counters"][k], f"{name}: {k} differs from the reference's printed counter"
    assert st.lcp_values == n                       # "Computed n/n LCP values" (ebwt2InDel.cpp:670)
    assert len(snp) == g["snp_bytes"]
    assert hashlib.sha256(snp.view()).hexdigest() == g["snp_sha256"], f"{name}: .snp differs from the compiled reference's"
    # the cluster-length histogram the reference prints (ebwt2InDel.cpp:1454-1462) from its stdout
    hist = {}
    for ln in g.get("stdout_lines", []):
        f = ln.split()
        if len(f) in (2, 3) and f[0].isdigit() and f[-1].isdigit() and (len(f) == 2 or set(f[1]) == {"-"}):
            hist[int(f[0])] = int(f[-1])
    if hist:
        for i, v in hist.items():
            assert st.clust_sizes[i] == v, (name, i)
    del snp
    gpu_ctx.trim()


def test_rank_access_beyond_2_32_positions(gpu_ctx):
    """rank / access / F on a random 5.2 G-symbol string (n > 2^32: 64-bit positions, 79 k superblocks of 2^16)
    against prefix counts computed independently with torch on the GPU."""
    import numpy as np
    import torch
    n = 5_200_000_123
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 3 * n:
        pytest.skip("not enough free device memory")
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    lut = torch.tensor(list(b"ACGT") * 63 + list(b"####"), dtype=torch.uint8, device=dev)       # ~1.6 % terminators
    bwt = torch.empty(n, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for off in range(0, n, step):
        k = min(step, n - off)
        bwt[off:off + k] = lut[torch.randint(0, 256, (k,), device=dev, generator=g)]
    torch.cuda.synchronize()              # the library works on its own stream: torch's fill must have finished
    ix = gpu_ctx.index(bwt)
    rng = np.random.default_rng(11)
    pos = np.unique(np.concatenate([rng.integers(0, n + 1, 200_000), [0, 1, n - 1, n, 1 << 32, (1 << 32) - 1, (1 << 32) + 1,
                                                                        1 << 16, (1 << 16) - 1, 65 * (1 << 16) + 31, 65 * (1 << 16) + 32]])).astype(np.uint64)
    got = ix.rank4(pos)
    # independent prefix counts: per-chunk totals, then an exact count inside the chunk of every query
    tot = torch.zeros((n + step - 1) // step + 1, 4, dtype=torch.int64, device=dev)
    for c, off in enumerate(range(0, n, step)):
        chunk = bwt[off:off + step]
        for s, ch in enumerate(b"ACGT"):
            tot[c + 1, s] = (chunk == ch).sum()
    pre = torch.cumsum(tot, 0).cpu().numpy()
    want = np.zeros_like(got)
    order = np.argsort(pos)
    pp = pos.astype(np.int64)
    for c, off in enumerate(range(0, n, step)):
        sel = order[(pp[order] >= off) & (pp[order] < off + step)] if off + step < n else order[pp[order] >= off]
        if not len(sel):
            continue
        chunk = bwt[off:off + step]
        q = torch.from_numpy(pp[sel] - off).to(dev)
        for s, ch in enumerate(b"ACGT"):
            cs = torch.cumsum((chunk == ch).to(torch.int32), 0, dtype=torch.int64)
            inside = torch.where(q > 0, cs[torch.clamp(q - 1, min=0)], torch.zeros_like(q))
            want[sel, s] = pre[c, s] + inside.cpu().numpy()
            del cs
    assert np.array_equal(got, want)
    F = ix.F()
    n_term = n - int(pre[-1].sum())
    assert list(F) == [n_term, n_term + pre[-1, 0], n_term + pre[-1, 0] + pre[-1, 1], n_term + pre[-1, :3].sum()]
    ipos = pos[pos < n][::40]
    assert np.array_equal(ix.access(ipos), bwt[torch.from_numpy(ipos.astype(np.int64)).to(dev)].cpu().numpy())
    del ix, bwt
    gpu_ctx.trim()
    torch.cuda.empty_cache()
