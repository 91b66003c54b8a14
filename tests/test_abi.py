"""CPU suite: the C-ABI library loads and exports every symbol include/e2i.h declares; host-only
entry points (parameter resolution, distance, .snp formatting) behave like the reference; compute
entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "e2i.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(e2i_[a-z0-9_A-Z]+)\s*\(", src)))


def test_exports_every_declared_symbol(e2i):
    L = e2i.lib()
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"libe2i.so does not export {n}"
    assert set(names) == set(e2i.SYMBOLS)


def test_struct_layouts(e2i):
    assert C.sizeof(e2i.Params) == 36
    assert C.sizeof(e2i.CallRec) == 56
    assert C.sizeof(e2i.Stats) == 8 * (10 + 201 + 8) + 8 * 6 + 8 * 3 + 8 * 2 + 8


def test_params_default_and_resolve(e2i):
    p = e2i.default_params()
    assert (p.k_left, p.k_right, p.K, p.max_gap, p.max_snvs, p.mcov_out, p.complexity,
            p.max_variants_per_position, p.term) == (31, 30, 16, 10, 2, 3, 20, 0, ord("#"))
    # 0 means default, also for -g (ebwt2InDel.cpp:1740-1746); -c default does not follow -R (:64)
    q = e2i.Params(k_left=0, k_right=50, K=0, max_gap=0, max_snvs=0, mcov_out=0, complexity=0,
                   max_variants_per_position=0, term=36)
    e2i.resolve_params(q)
    assert (q.k_left, q.k_right, q.K, q.max_gap, q.max_snvs, q.mcov_out, q.complexity, q.term) == \
        (31, 50, 16, 10, 2, 3, 20, 36)


def test_distance_matches_reference_example_and_oracle(e2i, oracle):
    assert e2i.distance("ACCTACTG", "TTACTTAC", 8) == (1, 2)      # ebwt2InDel.cpp:186-189
    assert e2i.distance("TTACTTAC", "ACCTACTG", 8) == (1, -2)
    rng = np.random.default_rng(1)
    for _ in range(500):
        n = int(rng.integers(5, 40))
        a = "".join(rng.choice(list("ACGT"), n))
        b = list(a)
        for _ in range(int(rng.integers(0, 4))):
            b[int(rng.integers(0, n))] = str(rng.choice(list("ACGT")))
        sh = int(rng.integers(0, 6))
        b = "".join(b)
        b = ("".join(rng.choice(list("ACGT"), sh)) + b)[:n] if rng.random() < .5 else b
        g = int(rng.integers(0, 12))
        assert e2i.distance(a, b, g) == oracle.distance(a, b, g), (a, b, g)


def test_snp_format_host_only(e2i):
    """to_file(vector<variant_single_t>) quirks (ebwt2InDel.cpp:1254-1330) on hand-made records."""
    p = e2i.default_params(k_left=8, k_right=6, max_gap=2, complexity=4)
    recs = np.zeros(3, dtype=e2i.CALL_REC_DTYPE)
    left = np.full(3 * 8 * 8, ord("A"), dtype=np.uint8)
    right = np.zeros(3 * 6, dtype=np.uint8)

    def put(r, slot, s):
        left[(r * 8 + slot) * 8:(r * 8 + slot) * 8 + 8] = np.frombuffer(s.encode(), dtype=np.uint8)

    # cluster 1: two alleles, SNP A/C
    put(0, 0, "ACGTACGA"); put(0, 1, "ACGTACGC")
    recs["has_right"] = 1
    recs[0]["n0"] = 2; recs[0]["right_len"] = 6; recs[0]["support"][:2] = (5, 4)
    right[0:6] = np.frombuffer(b"GATTAC", dtype=np.uint8)
    # cluster 2: two alleles but one below coverage -> nothing printed, cluster number still advances (:1328)
    put(1, 0, "ACGTACGA"); put(1, 1, "ACGTACGC")
    recs[1]["n0"] = 2; recs[1]["right_len"] = 6; recs[1]["support"][:2] = (5, 1)
    right[6:12] = np.frombuffer(b"GATTAC", dtype=np.uint8)
    # cluster 3: low-complexity right context (run of 4) -> filtered by has_run, number advances
    put(2, 0, "ACGTACGA"); put(2, 1, "ACGTACGT")
    recs[2]["n0"] = 2; recs[2]["right_len"] = 6; recs[2]["support"][:2] = (3, 3)
    right[12:18] = np.frombuffer(b"AAAATC", dtype=np.uint8)
    snp, st = e2i.snp_format(recs, left, right, p, two_samples=False)
    assert snp == (b">cluster:1_id:1_right:6_cov:5_type:_SNP_event:A/C\nACGTACGAGATTAC\n"
                   b">cluster:1_id:2_right:6_cov:4_type:_SNP_event:A/C\nACGTACGCGATTAC\n")
    assert st.events == 2 and st.clusters_out == 3


def test_no_cuda_device_fails_loudly(e2i):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(e2i.E2iError) as ei:
        e2i.Context(0)
    assert ei.value.code == e2i.E2I_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_cli_help_and_missing_file():
    exe = os.path.join(ROOT, "bin", "ebwt2InDel")
    assert os.access(exe, os.X_OK), "bin/ebwt2InDel is missing: run `make`"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ebwt2InDel [options]")          # argc < 3 -> help, exit 0
    r = subprocess.run([exe, "-1", "/nonexistent.ebwt", "-o", "/tmp/x.snp"], capture_output=True, text=True)
    assert r.returncode == 0 and "Error: could not find file /nonexistent.ebwt" in r.stdout


def test_parallel_formatter_equals_sequential_chaining(e2i):
    """e2i_snp_format formats record ranges on several host threads and fills in the sequential
    cluster numbers (ebwt2InDel.cpp:1250/1328) afterwards: the text must equal small single-thread
    slices chained through first_cluster_nr."""
    rng = np.random.default_rng(0)
    p = e2i.default_params(k_left=12, k_right=8, max_gap=3, complexity=5)
    N = 30000
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = np.zeros(N, dtype=e2i.CALL_REC_DTYPE)
    left = acgt[rng.integers(0, 4, size=(N, 8, 12))].copy()
    left[:, 1:, :] = left[:, :1, :]
    left[:, :, -1] = acgt[rng.integers(0, 4, size=(N, 8))]
    left = left.reshape(-1)
    right = acgt[rng.integers(0, 4, size=N * 8)].copy()
    recs["n0"] = rng.integers(0, 5, N)
    recs["n1"] = rng.integers(0, 5, N)
    recs["right_len"] = rng.integers(5, 9, N)
    recs["has_right"] = rng.integers(0, 8, N) > 0
    recs["support"] = rng.integers(1, 8, (N, 8))
    for two in (False, True):
        whole, st = e2i.snp_format(recs, left, right, p, two)
        out, nr, ev = b"", 1, 0
        for i in range(0, N, 1000):
            s, st2 = e2i.snp_format(recs[i:i + 1000], left[i * 96:(i + 1000) * 96], right[i * 8:(i + 1000) * 8], p, two,
                                    first_cluster_nr=nr)
            out += s
            nr += st2.clusters_out
            ev += st2.events
        assert whole == out
        assert st.clusters_out == nr - 1 and st.events == ev and len(whole) > 100000
        assert e2i.snp_count(recs, left, right, p, two) == st.clusters_out


def test_filter_snp_matches_reference_tool(e2i, oracle):
    """e2i_filter_snp / bin/filter_snp against the reference's filter_snp (filter_snp.cpp:17-81):
    committed expectations for the golden .snp files, and the compiled reference where it exists."""
    from conftest import load_golden
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "filter_snp")
    my_tool = os.path.join(ROOT, "bin", "filter_snp")
    for name in ("m1_default", "m1_flags", "m3_default", "m2_flags", "m1_metagenome"):
        snp = load_golden(name)["snp"]
        for m, M in ((0, 0), (3, 0), (5, 0), (10, 20), (8, 8), (1000, 0), (4, 2)):
            got = e2i.filter_snp(snp, m, M)
            # independent restatement in Python of the reference's loop
            lines = snp.split(b"\n")
            if lines and lines[-1] == b"":
                lines.pop()
            want = b""
            for i in range(0, len(lines) - 1, 2):
                tok = lines[i].split(b"_")
                cov = 0
                if len(tok) >= 4:
                    f = tok[3].split(b":")
                    if len(f) >= 2:
                        d = f[1].decode()
                        k = 0
                        while k < len(d) and (d[k].isdigit() or (k == 0 and d[k] in "+-")):
                            k += 1
                        cov = int(d[:k]) if d[:k] not in ("", "+", "-") else 0
                if cov >= m and (M == 0 or cov <= M):
                    want += lines[i] + b"\n" + lines[i + 1] + b"\n"
            assert got == want, (name, m, M)
            if os.access(ref_tool, os.X_OK):
                import tempfile
                with tempfile.NamedTemporaryFile(suffix=".snp") as f:
                    f.write(snp)
                    f.flush()
                    args = [str(m)] + ([str(M)] if M else [])
                    ref = subprocess.run([ref_tool, f.name] + args, capture_output=True).stdout
                    mine = subprocess.run([my_tool, f.name] + args, capture_output=True).stdout
                assert ref == got == mine, (name, m, M)
    assert len(e2i.filter_snp(load_golden("m1_default")["snp"], 5)) > 0


def test_cli_text_identical_to_reference_binary():
    """Help text and the missing-file message, byte for byte against the compiled reference (no GPU needed:
    both programs print and exit before touching the input, ebwt2InDel.cpp:76-103, 1748-1753)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ebwt2InDel")
    if not os.access(ref, os.X_OK):
        pytest.skip("compiled reference absent")
    mine = os.path.join(ROOT, "bin", "ebwt2InDel")
    for argv in ([], ["-h", "x"], ["-1", "/nonexistent.ebwt", "-o", "/tmp/e2i_x.snp"],
                 ["-1", os.path.join(ROOT, "README.md"), "-2", "/nonexistent2", "-o", "/tmp/e2i_x.snp"],
                 ["-1", os.path.join(ROOT, "README.md"), "-2", os.path.join(ROOT, "README.md"), "-d", "x", "-o", "/tmp/e2i_x.snp"]):
        a = subprocess.run([ref] + argv, capture_output=True)
        b = subprocess.run([mine] + argv, capture_output=True)
        assert a.stdout == b.stdout and a.returncode == b.returncode, argv
