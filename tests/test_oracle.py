"""CPU suite: the oracle (oracle/e2i_oracle.c) against the reference's golden vectors.

The reference's only known-answer test is the distance() example at ebwt2InDel.cpp:186-189; every
other vector under tests/golden/ was produced by the compiled, unmodified reference
(tests/golden/make_golden.py).  When the compiled reference is present (build container) the
oracle is also compared with it live on fresh seeded inputs.
"""
import os

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, resolved_fields


def test_distance_known_answer(oracle):
    # /root/reference/ebwt2InDel.cpp:186-189 (max_gap = default 10)
    assert oracle.distance("ACCTACTG", "TTACTTAC", 8) == (1, 2)
    assert oracle.distance("TTACTTAC", "ACCTACTG", 8) == (1, -2)


def test_distance_properties(oracle):
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = "".join(rng.choice(list("ACGT"), 31))
        assert oracle.distance(a, a, 10)[0] == 0
        b = a[:-1] + ("A" if a[-1] != "A" else "C")
        d = oracle.distance(a, b, 10)
        assert d[0] + abs(d[1]) <= 1
        # max_gap = 0 is plain right-aligned Hamming distance
        c = "".join(rng.choice(list("ACGT"), 31))
        assert oracle.distance(a, c, 0) == (sum(x != y for x, y in zip(a, c)), 0)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(oracle, name):
    g = load_golden(name)
    p = oracle.default_params(**resolved_fields(g["flags"]))
    snp, st = oracle.run(g["bwt1"], g["bwt2"], g["da"], p)
    d = st.as_dict()
    for k, v in g["counters"].items():
        assert d[k] == v, f"{name}: counter {k}: oracle {d[k]} != reference {v}"
    assert snp == g["snp"], f"{name}: .snp differs from the reference's"


def test_oracle_rank_access_select(oracle):
    rng = np.random.default_rng(3)
    n = 5000
    bwt = np.frombuffer(b"ACGT#", dtype=np.uint8)[rng.choice(5, size=n, p=[.24, .25, .25, .25, .01])]
    b = oracle.Bwt(bwt)
    pos = np.concatenate([np.arange(0, 300), rng.integers(0, n + 1, 500), [n]]).astype(np.uint64)
    got = b.rank4(pos)
    pref = np.zeros((n + 1, 4), dtype=np.uint64)
    for c, ch in enumerate(b"ACGT"):
        pref[1:, c] = np.cumsum(bwt == ch)
    assert np.array_equal(got, pref[pos.astype(np.int64)])
    F = b.F()
    cnt = [(bwt == ch).sum() for ch in b"#ACG"]
    assert list(F) == list(np.cumsum(cnt))
    lib = oracle.lib()
    for i in rng.integers(0, n, 200):
        assert lib.orc_access(b.h, int(i)) == bwt[i]
    for c, ch in enumerate(b"ACGT"):
        occ = np.flatnonzero(bwt == ch)
        for r in rng.integers(0, len(occ), 50):
            assert lib.orc_select(b.h, int(r), ch) == occ[r]


def test_forbidden_symbol(oracle):
    with pytest.raises(ValueError, match="position 3"):
        oracle.Bwt(np.frombuffer(b"ACGNACGT#", dtype=np.uint8))


def test_oracle_vs_compiled_reference_live(oracle):
    if not oracle.ref_available():
        pytest.skip("compiled reference absent (only built where /root/reference exists)")
    from ebwt2indel_b200 import synth
    reads = synth.diploid_reads(3000, 8, 3, 20, 60, seed=21)
    bwt, _ = synth.ebwt_naive(reads)
    snp_o, st = oracle.run(bwt)
    snp_r, cnt = oracle.run_ref(bwt)
    assert snp_o == snp_r
    d = st.as_dict()
    assert all(d[k] == v for k, v in cnt.items())
    r0, r1 = synth.two_individuals_reads(2000, 6, 2, 20, 60, seed=22)
    m, da = synth.merged_ebwt_da(r0, r1)
    assert oracle.run(m, None, da)[0] == oracle.run_ref(m, None, da)[0]
    b0, _ = synth.ebwt_naive(r0)
    b1, _ = synth.ebwt_naive(r1)
    assert oracle.run(b0, b1)[0] == oracle.run_ref(b0, b1)[0]


def test_bcr_builder_matches_naive():
    from ebwt2indel_b200 import synth
    reads = synth.diploid_reads(800, 4, 1, 12, 30, seed=5)
    a, oa = synth.ebwt_naive(reads)
    b, ob_ = synth.ebwt_bcr_numpy(reads)
    assert np.array_equal(a, b) and np.array_equal(oa, ob_)


@pytest.mark.skipif(not os.access(os.path.join(ROOT, "oracle", "_ref", "ebwt2InDel"), os.X_OK), reason="compiled reference absent")
def test_gap_longer_than_left_context_like_reference(oracle):
    """-g > -L: the reference accepts it (its substr(0, len - g) wraps to the whole string, ebwt2InDel.cpp:208-219);
    the oracle gives the same bytes, in modes -1 and -d."""
    for name in ("m1_default", "m3_default"):
        g = load_golden(name)
        flags = ("-L", 12, "-g", 20, "-R", 10, "-k", 8, "-c", 6)
        kw = {"k_left": 12, "max_gap": 20, "k_right": 10, "K": 8, "complexity": 6}
        want, _ = oracle.run_ref(g["bwt1"], g["bwt2"], g["da"], flags)
        snp, _ = oracle.run(g["bwt1"], g["bwt2"], g["da"], oracle.default_params(**kw))
        assert snp == want and len(snp) > 0
