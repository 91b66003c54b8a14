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
