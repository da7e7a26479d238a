"""Seeded synthetic inputs for the seed-and-extend hot path.

Reference genomes: several contigs, planted repeat families (diverged copies), tandem repeats,
low-complexity runs and a few N runs (so .amb is non-trivial).  Reads: 2x151 bp pairs, insert
N(400,40), substitutions / insertions / deletions / N, plus chimeric and garbage reads.
Pure numpy; deterministic for a given seed.  Not part of the product path.
"""
from __future__ import annotations
import numpy as np

_ALPHA = np.frombuffer(b"ACGTN", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def make_reference(total_bp: int, seed: int = 1, n_contigs: int = 4, repeat_frac: float = 0.15,
                   n_runs: int = 3):
    """Returns list of (name, codes uint8 array in 0..4)."""
    rng = np.random.default_rng(seed)
    # contig lengths: geometric-ish split
    w = np.array([0.5 ** i for i in range(n_contigs)], dtype=np.float64)
    lens = np.maximum((w / w.sum() * total_bp).astype(np.int64), 1000)
    contigs = [rng.integers(0, 4, size=int(l), dtype=np.uint8) for l in lens]
    # repeat families
    budget = int(total_bp * repeat_frac)
    while budget > 0:
        L = int(rng.integers(200, 3000))
        copies = int(rng.integers(3, 60))
        div = float(rng.choice([0.0, 0.005, 0.02, 0.05, 0.10]))
        unit = rng.integers(0, 4, size=L, dtype=np.uint8)
        for _ in range(copies):
            c = int(rng.integers(0, n_contigs))
            if len(contigs[c]) <= L + 10:
                continue
            p = int(rng.integers(0, len(contigs[c]) - L))
            u = unit.copy()
            if div > 0:
                m = rng.random(L) < div
                u[m] = (u[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            if rng.random() < 0.5:
                u = (3 - u)[::-1]
            contigs[c][p:p + L] = u
            budget -= L
    # tandem repeats and low-complexity
    for _ in range(max(4, total_bp // 500_000)):
        c = int(rng.integers(0, n_contigs))
        ul = int(rng.integers(1, 40))
        tot = int(rng.integers(60, 600))
        if len(contigs[c]) <= tot + 10:
            continue
        p = int(rng.integers(0, len(contigs[c]) - tot))
        unit = rng.integers(0, 4, size=ul, dtype=np.uint8)
        contigs[c][p:p + tot] = np.tile(unit, tot // ul + 1)[:tot]
    # N runs
    for _ in range(n_runs):
        c = int(rng.integers(0, n_contigs))
        tot = int(rng.integers(50, 800))
        if len(contigs[c]) <= tot + 10:
            continue
        p = int(rng.integers(0, len(contigs[c]) - tot))
        contigs[c][p:p + tot] = 4
    return [(f"chr{i + 1}", contigs[i]) for i in range(n_contigs)]


def write_fasta(path: str, contigs, width: int = 80):
    with open(path, "wb") as f:
        for name, codes in contigs:
            f.write(b">" + name.encode() + b"\n")
            s = _ALPHA[codes]
            n = len(s)
            full = (n // width) * width
            if full:
                body = np.empty((n // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def _mutate(rng, codes, sub, ins, dele, nrate, out_len):
    """Apply errors to a template (codes longer than out_len), return exactly out_len codes."""
    out = []
    i = 0
    n = len(codes)
    r = rng.random(size=3 * out_len + 16)
    k = 0
    while len(out) < out_len:
        if i >= n:
            out.append(int(rng.integers(0, 4)))
            continue
        x = r[k % len(r)]; k += 1
        if x < dele:
            i += 1
            continue
        if x < dele + ins:
            out.append(int(rng.integers(0, 4)))
            continue
        b = int(codes[i]); i += 1
        if b > 3:
            b = 4
        elif x < dele + ins + sub:
            b = (b + int(rng.integers(1, 4))) & 3
        elif x < dele + ins + sub + nrate:
            b = 4
        out.append(b)
    return np.array(out, dtype=np.uint8)


def make_pairs(contigs, n_pairs: int, read_len: int = 151, seed: int = 2, ins_mean: float = 400.0,
               ins_sd: float = 40.0, sub: float = 0.01, ins: float = 0.0015, dele: float = 0.0005,
               nrate: float = 0.001, garbage: float = 0.02):
    """Returns (r1, r2): arrays [n_pairs, read_len] of codes 0..4."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for _, c in contigs], dtype=np.float64)
    prob = lens / lens.sum()
    r1 = np.empty((n_pairs, read_len), dtype=np.uint8)
    r2 = np.empty((n_pairs, read_len), dtype=np.uint8)
    pad = read_len // 8 + 8
    for p in range(n_pairs):
        g = rng.random()
        if g < garbage / 2:
            r1[p] = rng.integers(0, 4, size=read_len, dtype=np.uint8)
            r2[p] = rng.integers(0, 4, size=read_len, dtype=np.uint8)
            continue
        while True:
            c = int(rng.choice(len(contigs), p=prob))
            ref = contigs[c][1]
            isz = max(int(rng.normal(ins_mean, ins_sd)), read_len + 5)
            if len(ref) > isz + 2 * pad:
                break
        st = int(rng.integers(0, len(ref) - isz - pad))
        frag = ref[st:st + isz + pad]
        a = frag[:read_len + pad]
        b = _COMP[frag[max(0, isz - read_len - pad):isz]][::-1]
        if rng.random() < 0.5:  # fragment from the reverse strand
            a, b = b, a
        m1 = _mutate(rng, a, sub, ins, dele, nrate, read_len)
        m2 = _mutate(rng, b, sub, ins, dele, nrate, read_len)
        if g < garbage:  # chimeric: second half of read 1 from elsewhere
            c2 = int(rng.choice(len(contigs), p=prob))
            ref2 = contigs[c2][1]
            s2 = int(rng.integers(0, max(1, len(ref2) - read_len)))
            h = int(rng.integers(40, read_len - 40))
            alt = ref2[s2:s2 + read_len - h]
            if len(alt) == read_len - h:
                m1[h:] = alt if rng.random() < 0.5 else _COMP[alt][::-1]
        r1[p] = m1
        r2[p] = m2
    return r1, r2


def make_long_reads(contigs, n_reads: int, read_len: int = 10000, seed: int = 3, sub: float = 0.04,
                    ins: float = 0.03, dele: float = 0.03):
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for _, c in contigs], dtype=np.float64)
    prob = lens / lens.sum()
    out = []
    for _ in range(n_reads):
        while True:
            c = int(rng.choice(len(contigs), p=prob))
            ref = contigs[c][1]
            if len(ref) > read_len * 1.2 + 100:
                break
        st = int(rng.integers(0, len(ref) - int(read_len * 1.2)))
        t = ref[st:st + int(read_len * 1.2)]
        if rng.random() < 0.5:
            t = _COMP[t][::-1]
        out.append(_mutate(rng, t, sub, ins, dele, 0.0, read_len))
    return out


def write_fastq(path: str, reads, prefix: str = "r", suffix: str = ""):
    with open(path, "wb") as f:
        for i, codes in enumerate(reads):
            s = _ALPHA[codes].tobytes()
            f.write(b"@" + f"{prefix}{i}{suffix}".encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")


if __name__ == "__main__":
    import argparse, os
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--ref-bp", type=int, default=10_000_000)
    ap.add_argument("--pairs", type=int, default=10_000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--contigs", type=int, default=4)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    ctg = make_reference(a.ref_bp, seed=a.seed, n_contigs=a.contigs)
    write_fasta(os.path.join(a.out, "ref.fa"), ctg)
    r1, r2 = make_pairs(ctg, a.pairs, seed=a.seed + 1)
    write_fastq(os.path.join(a.out, "r1.fq"), r1, "p")
    write_fastq(os.path.join(a.out, "r2.fq"), r2, "p")


def make_pairs_fast(contigs, n_pairs: int, read_len: int = 151, seed: int = 2, ins_mean: float = 400.0, ins_sd: float = 40.0,
                    sub: float = 0.01, indel_frac: float = 0.25, nrate: float = 0.001, garbage: float = 0.01):
    """Vectorised generator for large read sets (1M+ pairs): same error model as make_pairs except that
    each read carries at most one indel (1..4 bp) with probability `indel_frac`.  Returns (r1, r2)."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for _, c in contigs], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    genome = np.concatenate([c for _, c in contigs])
    cid = rng.choice(len(contigs), size=n_pairs, p=lens / lens.sum())
    isz = np.clip(rng.normal(ins_mean, ins_sd, n_pairs).astype(np.int64), read_len + 10, None)
    isz = np.minimum(isz, lens[cid] - 20)
    pos = (rng.random(n_pairs) * (lens[cid] - isz - 8)).astype(np.int64) + starts[cid]
    ext = read_len + 8
    ar = np.arange(ext, dtype=np.int64)
    fwd = genome[pos[:, None] + ar[None, :]]                                   # [n, ext] from the fragment start
    rev_idx = (pos + isz - 1)[:, None] - ar[None, :]
    rev = _COMP[genome[np.maximum(rev_idx, 0)]]                                # reverse complement from the fragment end
    flip = rng.random(n_pairs) < 0.5
    a = np.where(flip[:, None], rev, fwd); b = np.where(flip[:, None], fwd, rev)

    def mutate(t):
        n = t.shape[0]
        j = np.arange(read_len, dtype=np.int64)[None, :]
        has = rng.random(n) < indel_frac
        is_del = rng.random(n) < 0.5
        p = rng.integers(10, read_len - 10, n)[:, None]
        d = rng.integers(1, 5, n)[:, None]
        # deletion: skip d template bases at p; insertion: d random bases at p
        src_del = j + (j >= p) * d
        src_ins = np.where(j < p, j, np.maximum(j - d, p))
        src = np.where((has & is_del)[:, None], src_del, np.where((has & ~is_del)[:, None], src_ins, j))
        out = np.take_along_axis(t, src, axis=1)
        insmask = (has & ~is_del)[:, None] & (j >= p) & (j < p + d)
        out = np.where(insmask, rng.integers(0, 4, out.shape, dtype=np.uint8), out)
        m = (rng.random(out.shape) < sub) & (out < 4)
        out = np.where(m, (out + rng.integers(1, 4, out.shape, dtype=np.uint8)) & 3, out)
        out = np.where(rng.random(out.shape) < nrate, 4, out)
        return out.astype(np.uint8)

    r1 = mutate(a); r2 = mutate(b)
    g = rng.random(n_pairs) < garbage
    ng = int(g.sum())
    if ng:
        r1[g] = rng.integers(0, 4, (ng, read_len), dtype=np.uint8)
        r2[g] = rng.integers(0, 4, (ng, read_len), dtype=np.uint8)
    return r1, r2


def write_fastq_fast(path: str, reads: np.ndarray, prefix: bytes = b"p"):
    """Vectorised FASTQ writer: fixed-width names <prefix>%09d, constant quality 'I'."""
    n, L = reads.shape
    name_w = 1 + len(prefix) + 9 + 1
    rec = np.empty((n, name_w + L + 1 + 2 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord("@")
    rec[:, 1:1 + len(prefix)] = np.frombuffer(prefix, np.uint8)
    ids = np.arange(n, dtype=np.int64)
    for k in range(9):
        rec[:, 1 + len(prefix) + k] = ord("0") + (ids // 10 ** (8 - k)) % 10
    rec[:, name_w - 1] = 10
    rec[:, name_w:name_w + L] = _ALPHA[reads]
    rec[:, name_w + L] = 10
    rec[:, name_w + L + 1] = ord("+"); rec[:, name_w + L + 2] = 10
    rec[:, name_w + L + 3:name_w + 2 * L + 3] = ord("I")
    rec[:, -1] = 10
    rec.tofile(path)


def make_pairs_torch(genome, contig_lens, n_pairs: int, read_len: int = 151, seed: int = 2, ins_mean: float = 400.0, ins_sd: float = 40.0,
                     sub: float = 0.01, indel_frac: float = 0.25, nrate: float = 0.001, garbage: float = 0.01):
    """make_pairs_fast on a torch device (multi-Gbp genomes): `genome` uint8 tensor of all contigs concatenated."""
    import torch
    dev = genome.device
    g = torch.Generator(device=dev); g.manual_seed(seed)
    lens = torch.as_tensor(np.asarray(contig_lens, dtype=np.int64), device=dev)
    starts = torch.cumsum(lens, 0) - lens
    cid = torch.multinomial(lens.to(torch.float64) / lens.sum(), n_pairs, replacement=True, generator=g)
    isz = torch.clamp((torch.randn(n_pairs, device=dev, generator=g) * ins_sd + ins_mean).to(torch.int64), min=read_len + 10)
    isz = torch.minimum(isz, lens[cid] - 20)
    pos = (torch.rand(n_pairs, device=dev, generator=g, dtype=torch.float64) * (lens[cid] - isz - 8).to(torch.float64)).to(torch.int64) + starts[cid]
    ext = read_len + 8
    ar = torch.arange(ext, device=dev)
    fwd = genome[pos[:, None] + ar[None, :]]
    rev = 3 - genome[torch.clamp((pos + isz - 1)[:, None] - ar[None, :], min=0)]
    flip = torch.rand(n_pairs, device=dev, generator=g) < 0.5
    a = torch.where(flip[:, None], rev, fwd); b = torch.where(flip[:, None], fwd, rev)

    def mutate(t):
        n = t.shape[0]
        j = torch.arange(read_len, device=dev)[None, :]
        has = torch.rand(n, device=dev, generator=g) < indel_frac
        is_del = torch.rand(n, device=dev, generator=g) < 0.5
        p = torch.randint(10, read_len - 10, (n, 1), device=dev, generator=g)
        d = torch.randint(1, 5, (n, 1), device=dev, generator=g)
        src_del = j + (j >= p) * d
        src_ins = torch.where(j < p, j.expand(n, -1), torch.maximum(j - d, p))
        src = torch.where((has & is_del)[:, None], src_del, torch.where((has & ~is_del)[:, None], src_ins, j.expand(n, -1)))
        out = torch.gather(t, 1, src)
        insmask = (has & ~is_del)[:, None] & (j >= p) & (j < p + d)
        out = torch.where(insmask, torch.randint(0, 4, out.shape, dtype=torch.uint8, device=dev, generator=g), out)
        m = torch.rand(out.shape, device=dev, generator=g) < sub
        out = torch.where(m, (out + torch.randint(1, 4, out.shape, dtype=torch.uint8, device=dev, generator=g)) & 3, out)
        out = torch.where(torch.rand(out.shape, device=dev, generator=g) < nrate, torch.full_like(out, 4), out)
        return out

    r1 = mutate(a); r2 = mutate(b)
    gb = torch.rand(n_pairs, device=dev, generator=g) < garbage
    r1 = torch.where(gb[:, None], torch.randint(0, 4, r1.shape, dtype=torch.uint8, device=dev, generator=g), r1)
    r2 = torch.where(gb[:, None], torch.randint(0, 4, r2.shape, dtype=torch.uint8, device=dev, generator=g), r2)
    return r1.cpu().numpy(), r2.cpu().numpy()


def make_long_reads_torch(genome, contig_lens, n_reads: int, read_len: int = 10000, seed: int = 3, sub: float = 0.04, ins: float = 0.03,
                          dele: float = 0.03, block: int = 1024):
    """make_long_reads on a torch device (multi-Gbp genomes; config 5: 10 kbp reads, 4 % substitutions, 3 % insertions, 3 % deletions).
    Per template base: deleted with probability `dele`, else copied (substituted with probability `sub`); after it a random base is
    inserted with probability `ins`.  The read is the first read_len emitted bases; half of the reads come from the reverse strand.
    -> uint8 array (n_reads, read_len)."""
    import torch
    dev = genome.device
    g = torch.Generator(device=dev); g.manual_seed(seed)
    lens = torch.as_tensor(np.asarray(contig_lens, dtype=np.int64), device=dev)
    starts = torch.cumsum(lens, 0) - lens
    lt = int(read_len * 1.25) + 64
    ok = lens > lt + 16
    pr = torch.where(ok, lens.to(torch.float64), torch.zeros_like(lens, dtype=torch.float64))
    out = np.empty((n_reads, read_len), np.uint8)
    for b0 in range(0, n_reads, block):
        n = min(block, n_reads - b0)
        cid = torch.multinomial(pr / pr.sum(), n, replacement=True, generator=g)
        pos = (torch.rand(n, device=dev, generator=g, dtype=torch.float64) * (lens[cid] - lt - 8).to(torch.float64)).to(torch.int64) + starts[cid]
        ar = torch.arange(lt, device=dev)
        fwd = genome[pos[:, None] + ar[None, :]]
        rev = 3 - genome[(pos + lt - 1)[:, None] - ar[None, :]]
        flip = torch.rand(n, device=dev, generator=g) < 0.5
        t = torch.where(flip[:, None], rev, fwd)
        u = torch.rand((n, lt), device=dev, generator=g)
        is_del = u < dele
        is_sub = (u >= dele) & (u < dele + sub)
        has_ins = torch.rand((n, lt), device=dev, generator=g) < ins
        base = torch.where(is_sub, (t + torch.randint(1, 4, t.shape, dtype=torch.uint8, device=dev, generator=g)) & 3, t)
        rnd = torch.randint(0, 4, t.shape, dtype=torch.uint8, device=dev, generator=g)
        emit = (~is_del).to(torch.int32) + has_ins.to(torch.int32)
        cum = torch.cumsum(emit, 1)
        j = torch.arange(read_len, device=dev, dtype=torch.int32)[None, :].expand(n, -1).contiguous()
        src = torch.searchsorted(cum, j, right=True).clamp(max=lt - 1)                 # template base that emits output position j
        k = j - (torch.gather(cum, 1, src) - torch.gather(emit, 1, src))                 # 0: first emitted base of that template position
        first_is_copy = ~torch.gather(is_del, 1, src)
        r = torch.where((k == 0) & first_is_copy, torch.gather(base, 1, src), torch.gather(rnd, 1, src))
        out[b0:b0 + n] = r.cpu().numpy()
    return out
