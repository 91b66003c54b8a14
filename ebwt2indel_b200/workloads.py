"""Named synthetic workloads (SURVEY.md §8(d) / BASELINE.json configs): test + bench tooling, not product.

One table so that bench.py, the GPU tests and tests/golden/make_big_golden.py (which runs the compiled
reference on the same inputs) agree on what "C4 at scale 0.3" means.  `scale` < 1 shrinks genome
length and variant counts together (same coverage, read length and variant density).
"""
from __future__ import annotations

from . import synth

CONFIGS = {
    "C1": dict(mode=1, genome=1_000_000, snps=1000, indels=200, cov=20, read_len=100, revcomp=True, seed=1,
               desc="mode -1: 1 Mbp random diploid (1k SNPs, 200 indels), 20x 100bp + revcomp, n=40.4M"),
    "C2": dict(mode=3, genome=5_000_000, snps=5000, indels=1000, cov=50, read_len=100, revcomp=True, seed=2,
               desc="mode -d: two 5 Mbp individuals, 50x 100bp + revcomp each, merged eBWT + DA, n=1.01G"),
    "C3": dict(mode=2, genome=50_000_000, snps=50_000, indels=10_000, cov=30, read_len=150, revcomp=False, seed=3,
               desc="mode -2: two eBWTs of 50 Mbp genomes, 30x 150bp, n=1.51G each"),
    "C4": dict(mode=1, genome=250_000_000, snps=250_000, indels=50_000, cov=30, read_len=150, revcomp=True, seed=4,
               desc="mode -1: 250 Mbp diploid, 30x 150bp + revcomp, n=15.1G"),
    "C5": dict(mode=1, meta=True, species=10, strains=10, genome=10_000_000, snp_rate=0.001, cov=10, read_len=100, revcomp=True,
               seed=5, snps=0, indels=0,
               desc="mode -1: metagenome, 10 species x 10 strains x 10 Mbp (0.1% SNPs between strains, log-normal abundances), "
                    "10x mean 100bp + revcomp, n=20.2G"),
    # intermediate mode -1 sizes (same shape as C4)
    "C4s16": dict(mode=1, genome=15_625_000, snps=15_625, indels=3_125, cov=30, read_len=150, revcomp=True, seed=4,
                  desc="mode -1: 1/16 of C4 (15.6 Mbp diploid, 30x 150bp + revcomp), n=0.94G"),
    "C4s4": dict(mode=1, genome=62_500_000, snps=62_500, indels=12_500, cov=30, read_len=150, revcomp=True, seed=4,
                 desc="mode -1: 1/4 of C4 (62.5 Mbp diploid, 30x 150bp + revcomp), n=3.8G"),
}


def scaled(cfg: dict, scale: float) -> dict:
    c = dict(cfg)
    c["genome"] = max(2000, int(cfg["genome"] * scale))
    c["snps"] = max(1, int(cfg["snps"] * scale))
    c["indels"] = max(1, int(cfg["indels"] * scale))
    return c


def plans_for(cfg: dict):
    """Read plans of a workload: (mode, plans of eBWT 1, plans of eBWT 2 or None).
    mode 1: one read set; mode 3: two read sets merged into ONE eBWT (+ document array);
    mode 2: one read set per eBWT."""
    if cfg.get("meta"):
        return 1, [synth.metagenome_plan(cfg["species"], cfg["strains"], cfg["genome"], cfg["snp_rate"], cfg["cov"],
                                         cfg["read_len"], cfg["seed"], cfg["revcomp"])], None
    if cfg["mode"] == 1:
        return 1, [synth.diploid_plan(cfg["genome"], cfg["snps"], cfg["indels"], cfg["cov"], cfg["read_len"], cfg["seed"],
                                      cfg["revcomp"])], None
    p0, p1 = synth.two_individuals_plans(cfg["genome"], cfg["snps"], cfg["indels"], cfg["cov"], cfg["read_len"], cfg["seed"],
                                         cfg["revcomp"])
    if cfg["mode"] == 3:
        return 3, [p0, p1], None
    return 2, [p0], [p1]


def make_workload_gpu(cfg: dict, device, ctx):
    """Inputs of the named shape built on the GPU (synth.ebwt_bcr_gpu; nothing is materialised on the
    host): dict(mode, bwt1, bwt2, da, n, reads) of uint8 tensors on `device`."""
    mode, pl1, pl2 = plans_for(cfg)
    reads = sum(p.n_reads for p in pl1) + (sum(p.n_reads for p in pl2) if pl2 else 0)
    if mode == 3:
        bwt, owner = synth.ebwt_bcr_gpu(ctx, pl1, device, want_owner=True)
        return dict(mode=3, bwt1=bwt, bwt2=None, da=owner + 48, n=bwt.numel(), reads=reads)      # ASCII '0' / '1'
    b1 = synth.ebwt_bcr_gpu(ctx, pl1, device)
    b2 = synth.ebwt_bcr_gpu(ctx, pl2, device) if pl2 else None
    return dict(mode=mode, bwt1=b1, bwt2=b2, da=None, n=b1.numel() + (b2.numel() if b2 is not None else 0), reads=reads)


def checksum(t) -> int:
    """Position-weighted checksum of a uint8 torch tensor (any device), equal to oracle/bcr_build.c's
    orc_checksum: sum_i b[i] * ((i mod 2^20) + 1) + (sum_i b[i]) * 2^40 (mod 2^64)."""
    import torch
    n = t.numel()
    w = torch.arange(1, (1 << 20) + 1, dtype=torch.int64, device=t.device)
    a = 0
    s = 0
    step = 1 << 27                                   # multiple of 2^20: every chunk starts at weight 1
    for off in range(0, n, step):
        c = t[off:off + step].to(torch.int64)
        full = c.numel() >> 20 << 20
        if full:
            a += int((c[:full].view(-1, 1 << 20) * w).sum().item())
        if c.numel() > full:
            a += int((c[full:] * w[:c.numel() - full]).sum().item())
        s += int(c.sum().item())
    return (a + (s << 40)) & 0xFFFFFFFFFFFFFFFF
