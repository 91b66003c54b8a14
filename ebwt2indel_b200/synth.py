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


def sample_reads(haplotypes, n_reads: int, read_len: int, rng: np.random.Generator,
                 revcomp: bool = True) -> np.ndarray:
    """Error-free reads, uniform starts, split evenly over `haplotypes`; (m, read_len) ASCII matrix.

    With `revcomp`, the reverse complement of every read is appended after all forward reads.
    """
    per = n_reads // len(haplotypes)
    out = []
    idx = np.arange(read_len)
    for h in haplotypes:
        starts = rng.integers(0, len(h) - read_len + 1, size=per)
        out.append(h[starts[:, None] + idx[None, :]])
    fwd = np.concatenate(out, axis=0)
    if not revcomp:
        return fwd
    rc = _COMP[fwd[:, ::-1]]
    return np.concatenate([fwd, rc], axis=0)


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


def diploid_reads(genome_len: int, n_snps: int, n_indels: int, coverage: float, read_len: int,
                  seed: int, revcomp: bool = True) -> np.ndarray:
    """Mode -1 shape: one diploid individual (two haplotypes), total `coverage` split over both."""
    rng = np.random.default_rng(seed)
    h1 = random_genome(genome_len, rng)
    h2 = mutate(h1, n_snps, n_indels, rng)
    n_reads = int(round(coverage * genome_len / read_len))
    return sample_reads([h1, h2], n_reads, read_len, rng, revcomp)


def two_individuals_reads(genome_len: int, n_snps: int, n_indels: int, coverage: float,
                          read_len: int, seed: int, revcomp: bool = True):
    """Modes -2/-d shape: two haploid individuals, `coverage` each; returns (reads0, reads1)."""
    rng = np.random.default_rng(seed)
    g1 = random_genome(genome_len, rng)
    g2 = mutate(g1, n_snps, n_indels, rng)
    n_reads = int(round(coverage * genome_len / read_len))
    r0 = sample_reads([g1], n_reads, read_len, rng, revcomp)
    r1 = sample_reads([g2], n_reads, read_len, rng, revcomp)
    return r0, r1


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
