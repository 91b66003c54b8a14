"""Seeded synthetic inputs for the ebwt2InDel hot path (test + bench tooling, not product code).

The reference ships no data: every parity corpus is generated here.
Shapes follow SURVEY.md §8(d): an i.i.d. uniform genome, a second haplotype / individual
derived from it by SNPs and short indels, error-free fixed-length reads with uniform starts,
optionally with reverse complements, and the eBWT of the read collection under BCR's
convention (``#_i < #_j`` for i < j, ``# < A < C < G < T``; raw ASCII, one byte per symbol,
no header).  The input format is what ``dna_bwt(path, TERM)`` reads
(/root/reference/internal/dna_bwt.hpp:36-62, dna_string.hpp:55-110) and the document array is
the ASCII '0'/'1' file read at /root/reference/ebwt2InDel.cpp:1495-1508.
"""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def random_genome(length: int, rng: np.random.Generator) -> np.ndarray:
    """i.i.d. uniform ACGT genome as ASCII bytes."""
    return BASES[rng.integers(0, 4, size=length, dtype=np.uint8)]


def mutate(genome: np.ndarray, n_snps: int, n_indels: int, rng: np.random.Generator,
           margin: int = 200, max_indel: int = 10) -> np.ndarray:
    """Second haplotype: SNPs (uniform alt base) and indels (length U[1,max_indel], 50/50 ins/del)."""
    g = genome.copy()
    n = len(g)
    margin = min(margin, max(0, n // 4))
    lo, hi = margin, max(margin + 1, n - margin)
    if n_snps:
        pos = rng.choice(np.arange(lo, hi), size=min(n_snps, hi - lo), replace=False)
        shift = rng.integers(1, 4, size=len(pos))
        code = np.searchsorted(BASES, g[pos])
        g[pos] = BASES[(code + shift) % 4]
    if not n_indels:
        return g
    ipos = np.sort(rng.choice(np.arange(lo, hi), size=min(n_indels, hi - lo), replace=False))
    pieces = []
    prev = 0
    for p in ipos:
        ln = int(rng.integers(1, max_indel + 1))
        if p < prev:
            continue
        pieces.append(g[prev:p])
        if rng.random() < 0.5:
            pieces.append(random_genome(ln, rng))  # insertion
            prev = p
        else:
            prev = min(n, p + ln)  # deletion
    pieces.append(g[prev:])
    return np.concatenate(pieces)


class ReadPlan:
    """Reads described by start offsets into the concatenated haplotypes (nothing materialised):
    forward read i = hap[starts[i] : starts[i] + read_len]; with `revcomp`, read n_fwd + i is the
    reverse complement of forward read i.  Lets the GPU builder make 10^8-read sets in place."""

    def __init__(self, hap: np.ndarray, starts: np.ndarray, read_len: int, revcomp: bool):
        self.hap, self.starts, self.read_len, self.revcomp = hap, starts, read_len, revcomp

    @property
    def n_fwd(self) -> int:
        return len(self.starts)

    @property
    def n_reads(self) -> int:
        return len(self.starts) * (2 if self.revcomp else 1)

    def materialize(self) -> np.ndarray:
        idx = np.arange(self.read_len)
        fwd = self.hap[self.starts[:, None] + idx[None, :]]
        if not self.revcomp:
            return fwd
        return np.concatenate([fwd, _COMP[fwd[:, ::-1]]], axis=0)


def read_plan(haplotypes, n_reads: int, read_len: int, rng: np.random.Generator, revcomp: bool = True) -> ReadPlan:
    """Error-free reads, uniform starts, split evenly over `haplotypes` (same draws as sample_reads)."""
    per = n_reads // len(haplotypes)
    starts, off = [], 0
    for h in haplotypes:
        starts.append(off + rng.integers(0, len(h) - read_len + 1, size=per))
        off += len(h)
    return ReadPlan(np.concatenate(haplotypes), np.concatenate(starts).astype(np.int64), read_len, revcomp)


def sample_reads(haplotypes, n_reads: int, read_len: int, rng: np.random.Generator,
                 revcomp: bool = True) -> np.ndarray:
    """Error-free reads, uniform starts, split evenly over `haplotypes`; (m, read_len) ASCII matrix.

    With `revcomp`, the reverse complement of every read is appended after all forward reads.
    """
    return read_plan(haplotypes, n_reads, read_len, rng, revcomp).materialize()


def ebwt_naive(reads: np.ndarray, term: int = ord("#")):
    """eBWT of a read collection by sorting every read suffix (ground truth for small inputs).

    Suffix (r, k) = reads[r, k:] followed by #_r, k = 0..L.  Order: # < A < C < G < T, ties
    (identical suffixes of different reads) by read index.  BWT symbol = reads[r, k-1], or #
    for k = 0.  Returns (bwt ASCII bytes, read index of every suffix).
    """
    m, L = reads.shape
    code = np.zeros(256, dtype=np.uint8)
    for i, b in enumerate(b"ACGT"):
        code[b] = i + 1
    enc = code[reads]
    keys = np.zeros((m, L + 1, L + 1), dtype=np.uint8)
    for k in range(L + 1):
        keys[:, k, : L - k] = enc[:, k:]
    flat = np.ascontiguousarray(keys.reshape(m * (L + 1), L + 1)).view(f"S{L + 1}").ravel()
    order = np.argsort(flat, kind="stable")  # initial order is read-major -> ties by read index
    r = order // (L + 1)
    k = order % (L + 1)
    bwt = np.where(k == 0, np.uint8(term), reads[r, np.maximum(k, 1) - 1]).astype(np.uint8)
    return bwt, r


def ebwt_bcr_numpy(reads: np.ndarray, term: int = ord("#")):
    """eBWT by BCR-style column insertion (numpy restatement of the GPU builder's algorithm).

    Iteration k holds BWT_k, the symbols preceding all suffixes of length <= k, and P[r], the
    position of read r's length-k suffix.  The length-(k+1) suffix sits at LF_k(P[r]) in
    BWT_{k+1}; the old symbols keep their relative order.  Returns (bwt, read index per position).
    """
    m, L = reads.shape
    bwt = reads[:, L - 1].copy()
    owner = np.arange(m, dtype=np.int64)
    P = np.arange(m, dtype=np.int64)
    for k in range(L):
        c = bwt[P]
        newP = np.empty(m, dtype=np.int64)
        base = m
        for sym in b"ACGT":
            is_c = bwt == sym
            rank = np.cumsum(is_c) - is_c  # exclusive rank
            sel = c == sym
            newP[sel] = base + rank[P[sel]]
            base += int(is_c.sum())
        new_sym = reads[:, L - k - 2] if k + 1 < L else np.full(m, term, dtype=np.uint8)
        size = len(bwt) + m
        mark = np.zeros(size, dtype=bool)
        mark[newP] = True
        nb = np.empty(size, dtype=np.uint8)
        no = np.empty(size, dtype=np.int64)
        nb[~mark] = bwt
        no[~mark] = owner
        nb[newP] = new_sym
        no[newP] = np.arange(m)
        bwt, owner, P = nb, no, newP
    return bwt, owner


def diploid_plan(genome_len: int, n_snps: int, n_indels: int, coverage: float, read_len: int,
                 seed: int, revcomp: bool = True) -> ReadPlan:
    """Mode -1 shape: one diploid individual (two haplotypes), total `coverage` split over both."""
    rng = np.random.default_rng(seed)
    h1 = random_genome(genome_len, rng)
    h2 = mutate(h1, n_snps, n_indels, rng)
    n_reads = int(round(coverage * genome_len / read_len))
    return read_plan([h1, h2], n_reads, read_len, rng, revcomp)


def diploid_reads(genome_len: int, n_snps: int, n_indels: int, coverage: float, read_len: int,
                  seed: int, revcomp: bool = True) -> np.ndarray:
    return diploid_plan(genome_len, n_snps, n_indels, coverage, read_len, seed, revcomp).materialize()


def two_individuals_plans(genome_len: int, n_snps: int, n_indels: int, coverage: float,
                          read_len: int, seed: int, revcomp: bool = True):
    """Modes -2/-d shape: two haploid individuals, `coverage` each; returns (plan0, plan1)."""
    rng = np.random.default_rng(seed)
    g1 = random_genome(genome_len, rng)
    g2 = mutate(g1, n_snps, n_indels, rng)
    n_reads = int(round(coverage * genome_len / read_len))
    p0 = read_plan([g1], n_reads, read_len, rng, revcomp)
    p1 = read_plan([g2], n_reads, read_len, rng, revcomp)
    return p0, p1


def two_individuals_reads(genome_len: int, n_snps: int, n_indels: int, coverage: float,
                          read_len: int, seed: int, revcomp: bool = True):
    p0, p1 = two_individuals_plans(genome_len, n_snps, n_indels, coverage, read_len, seed, revcomp)
    return p0.materialize(), p1.materialize()


def metagenome_plan(n_species: int, strains_per_species: int, genome_len: int, snp_rate: float, coverage: float,
                    read_len: int, seed: int, revcomp: bool = True, sigma: float = 1.0) -> ReadPlan:
    """Mode -1 metagenome shape (BASELINE.json configs[4]): `n_species` random base genomes, each with
    `strains_per_species` strains that differ from it by `snp_rate` SNPs; strain abundances are
    log-normal(0, sigma), so most variants are low-frequency; `coverage` is the mean over the total length."""
    rng = np.random.default_rng(seed)
    strains = []
    for _ in range(n_species):
        base = random_genome(genome_len, rng)
        for _ in range(strains_per_species):
            strains.append(mutate(base, int(round(snp_rate * genome_len)), 0, rng))
    w = rng.lognormal(0.0, sigma, size=len(strains))
    w /= w.sum()
    n_reads = int(round(coverage * genome_len * len(strains) / read_len))
    per = np.maximum(1, np.round(w * n_reads).astype(np.int64))
    starts, off = [], 0
    for h, k in zip(strains, per):
        starts.append(off + rng.integers(0, len(h) - read_len + 1, size=int(k)))
        off += len(h)
    return ReadPlan(np.concatenate(strains), np.concatenate(starts).astype(np.int64), read_len, revcomp)


def merged_ebwt_da(reads0: np.ndarray, reads1: np.ndarray, builder=ebwt_naive):
    """Merged eBWT of (reads0 then reads1) plus the ASCII '0'/'1' document array (mode -d input)."""
    bwt, owner = builder(np.concatenate([reads0, reads1], axis=0))
    da = np.where(owner >= len(reads0), np.uint8(ord("1")), np.uint8(ord("0"))).astype(np.uint8)
    return bwt, da


def ebwt_bcr_torch(reads, device=None, term: int = ord("#"), want_owner: bool = False):
    """eBWT by BCR-style column insertion with torch ops (runs on the GPU for bench-sized inputs).

    Same algorithm and result as `ebwt_bcr_numpy`; `reads` is an (m, L) uint8 array/tensor.  Returns
    a uint8 tensor on `device` (and, with `want_owner`, a bool tensor: suffix belongs to the second
    half of the reads -- used to derive the document array of a merged collection).
    Tooling for synthetic inputs only: not part of the product path.
    """
    import torch

    if not torch.is_tensor(reads):
        reads = torch.from_numpy(np.ascontiguousarray(reads))
    dev = torch.device(device) if device is not None else reads.device
    reads = reads.to(dev)
    m, L = reads.shape
    idt = torch.int32 if m * (L + 1) < 2 ** 31 - 1 else torch.int64
    bwt = reads[:, L - 1].clone()
    P = torch.arange(m, device=dev, dtype=idt)
    owner = (torch.arange(m, device=dev) >= m // 2) if want_owner else None
    second = owner.clone() if want_owner else None
    syms = [ord(c) for c in "ACGT"]
    for k in range(L):
        c = reads[:, L - 1 - k]                     # == bwt[P]
        newP = torch.empty(m, device=dev, dtype=idt)
        base = m
        for sym in syms:
            is_c = bwt == sym
            incl = torch.cumsum(is_c, 0, dtype=idt)
            sel = c == sym
            Ps = P[sel].long()
            newP[sel] = (base + incl[Ps] - is_c[Ps].to(idt)).to(idt)
            base += int(incl[-1])
            del incl, is_c
        new_sym = reads[:, L - k - 2] if k + 1 < L else torch.full((m,), term, dtype=torch.uint8, device=dev)
        size = bwt.numel() + m
        mark = torch.zeros(size, dtype=torch.bool, device=dev)
        idx = newP.long()
        mark[idx] = True
        nb = torch.empty(size, dtype=torch.uint8, device=dev)
        keep = ~mark
        nb[keep] = bwt
        nb[idx] = new_sym
        if want_owner:
            no = torch.empty(size, dtype=torch.bool, device=dev)
            no[keep] = owner
            no[idx] = second
            owner = no
        bwt, P = nb, newP
        del mark, keep, idx
    return (bwt, owner) if want_owner else bwt


# ---- GPU builder (bench-sized inputs) ------------------------------------------------------------
_tools = None


def tools_lib():
    """libe2i_tools.so: merge kernels of the GPU eBWT builder (csrc/tools.cu); tooling, not product."""
    global _tools
    if _tools is None:
        import ctypes as C
        import os
        L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libe2i_tools.so"))
        vp, u64 = C.c_void_p, C.c_uint64
        L.e2i_tool_bcr_mark.argtypes = [vp, u64, vp, u64, vp]
        L.e2i_tool_bcr_mark.restype = C.c_int
        L.e2i_tool_bcr_merge.argtypes = [vp, vp, vp, vp, u64, u64, vp, vp, vp, vp]
        L.e2i_tool_bcr_merge.restype = C.c_int
        L.e2i_tool_gather_bench.argtypes = [vp, u64, C.c_int, u64, vp, C.POINTER(C.c_float)]
        L.e2i_tool_gather_bench.restype = C.c_int
        _tools = L
    return _tools


def random_sector_bandwidth(device, buf_bytes: int = 8 << 30, n_access: int = 1 << 28):
    """Measured GB/s of independent random 32 / 64 / 128-byte gathers over `buf_bytes` of HBM: the
    random-sector roofline of an unsorted rank-query stream (BASELINE.json north_star)."""
    import ctypes as C
    import torch
    T = tools_lib()
    buf = torch.empty(buf_bytes, dtype=torch.uint8, device=device)
    buf.random_(0, 255)
    sink = torch.zeros(4, dtype=torch.int32, device=device)
    out = {}
    torch.cuda.synchronize(device)
    for sector in (32, 64, 128):
        ms = C.c_float(0)
        rc = T.e2i_tool_gather_bench(buf.data_ptr(), buf_bytes, sector, n_access, sink.data_ptr(), C.byref(ms))
        assert rc == 0
        out[str(sector)] = n_access * sector / (ms.value / 1e3) / 1e9
    del buf
    return out


class DevicePlans:
    """One or more ReadPlans (read sets, concatenated in order) resident on the GPU; column(j) is
    reads[:, j] over all reads, computed by gathers from the haplotypes."""

    def __init__(self, plans, device):
        import torch
        self.dev = torch.device(device)
        self.L = plans[0].read_len
        self.parts = []
        comp = torch.zeros(256, dtype=torch.uint8)
        for a, b in zip(b"ACGT", b"TGCA"):
            comp[a] = b
        self.comp = comp.to(self.dev)
        for p in plans:
            assert p.read_len == self.L
            self.parts.append((torch.from_numpy(p.hap).to(self.dev), torch.from_numpy(p.starts).to(self.dev), p.revcomp))
        self.m = sum(p.n_reads for p in plans)
        self.first_of_second_set = plans[0].n_reads if len(plans) > 1 else None

    def column(self, j: int):
        import torch
        out = []
        for hap, starts, rc in self.parts:
            out.append(hap[starts + j])
            if rc:
                out.append(self.comp[hap[starts + (self.L - 1 - j)].long()])
        return torch.cat(out)


def ebwt_bcr_gpu(ctx, plans, device, term: int = ord("#"), want_owner: bool = False):
    """eBWT of the reads of `plans` by BCR-style column insertion on the GPU.

    Same result as `ebwt_bcr_torch` on the materialised reads.  The LF step of every iteration runs
    on the product's own kernels (index build + batched rank through `ctx`, an api.Context); the
    merge runs on csrc/tools.cu.  The reads stay sorted by the position of their newest suffix
    (a stable 4-way partition by symbol keeps that order), so targets come out sorted and both the
    rank queries and the merge are monotone sweeps.  Returns a uint8 tensor (and, with
    `want_owner`, a uint8 0/1 tensor: suffix belongs to the second read set).
    """
    import torch
    T = tools_lib()
    dp = DevicePlans(plans, device)
    dev, m, L = dp.dev, dp.m, dp.L
    lut = torch.full((256,), 0, dtype=torch.int64, device=dev)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    bwt = dp.column(L - 1).clone()
    P = torch.arange(m, dtype=torch.int64, device=dev)
    order = torch.arange(m, dtype=torch.int64, device=dev)
    read_owner = owner = None
    if want_owner:
        read_owner = (order >= dp.first_of_second_set).to(torch.uint8)
        owner = read_owner.clone()
    ranks = torch.empty((m, 4), dtype=torch.int64, device=dev)
    for k in range(L):
        S = bwt.numel()
        ix = ctx.index(bwt, term)
        torch.cuda.synchronize(dev)
        ix.rank4_device(P, ranks)
        F = torch.from_numpy(ix.F().astype(np.int64)).to(dev)
        ix.close()
        code = lut[dp.column(L - 1 - k)[order].long()]
        newP = m + F[code] + ranks.gather(1, code[:, None]).squeeze(1)
        perm = torch.cat([torch.nonzero(code == c).squeeze(1) for c in range(4)])   # stable 4-way partition
        order = order[perm]
        newP = newP[perm].contiguous()
        del perm, code
        new_sym = dp.column(L - 2 - k)[order] if k + 1 < L else torch.full((m,), term, dtype=torch.uint8, device=dev)
        n_out = S + m
        n_words = ((n_out + 31) // 32 + 127) // 128 * 128
        bits = torch.zeros(n_words, dtype=torch.int32, device=dev)
        groups = torch.empty(n_words // 128, dtype=torch.int64, device=dev)
        rc = T.e2i_tool_bcr_mark(newP.data_ptr(), m, bits.data_ptr(), n_words, groups.data_ptr())
        assert rc == 0
        gex = torch.cumsum(groups, 0) - groups
        out = torch.empty(n_out, dtype=torch.uint8, device=dev)
        new_own = out_own = None
        if want_owner:
            new_own = read_owner[order]
            out_own = torch.empty(n_out, dtype=torch.uint8, device=dev)
        rc = T.e2i_tool_bcr_merge(bwt.data_ptr(), bits.data_ptr(), gex.data_ptr(), new_sym.data_ptr(), n_out, n_words,
                                  out.data_ptr(), owner.data_ptr() if want_owner else None,
                                  new_own.data_ptr() if want_owner else None, out_own.data_ptr() if want_owner else None)
        assert rc == 0
        torch.cuda.synchronize(dev)
        bwt, P, owner = out, newP, out_own
        del bits, groups, gex, new_sym
    return (bwt, owner) if want_owner else bwt
