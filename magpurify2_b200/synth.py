"""Deterministic synthetic MAGs and alignments.

Used by tests/ (small, NumPy on the host) and bench.py (large, generated on the device with torch
used purely as RNG + buffer plumbing).  Nothing here is on the product path.

MAG model: total length ~N(3.0 Mbp, 0.3 Mbp); ~300 contigs with log-normal lengths (sigma 1.0)
rescaled to the total, floor 1500 bp, plus 1 % deliberately tiny contigs (0..200 bp, which covers
the < 4 bp and <= 150 bp edge rules); per-MAG GC ~U(0.30, 0.70); 5 % "contaminant" contigs at
GC +- 0.12; N runs (Poisson, 1 per 100 kbp, length U(1,100)); 2 % lower-case stretches; one
IUPAC byte per Mbp.  `long_contigs=True` gives the C5 stress shape (5-10 Mbp, 3..20 contigs, the
largest >= 50 % of the MAG).
"""
import numpy as np

SEED_SEQ = 0x4D414732
SEED_COV = 0x434F5631
READ_LEN = 150


def mag_layout(n_mags, seed=0, long_contigs=False, mean_bp=3.0e6, n_contigs=300):
    """Contig lengths for n_mags MAGs.

    Returns (contig_len int64[n], mag_off int64[n_mags+1], gc float64[n] per-contig GC target)."""
    lens, gcs, mag_off = [], [], [0]
    for m in range(n_mags):
        rng = np.random.Generator(np.random.Philox(key=(SEED_SEQ ^ m) + (seed << 32)))
        if long_contigs:
            total = rng.uniform(5e6, 10e6)
            k = int(rng.integers(3, 21))
            w = rng.lognormal(0.0, 1.0, k)
            w[0] = w[1:].sum() * rng.uniform(1.0, 3.0)  # largest contig >= 50 %
            ln = np.maximum(1500, (w / w.sum() * total)).astype(np.int64)
        else:
            total = max(2e5, rng.normal(mean_bp, mean_bp * 0.1))
            k = max(3, int(rng.normal(n_contigs, n_contigs * 0.1)))
            w = rng.lognormal(0.0, 1.0, k)
            ln = np.maximum(1500, (w / w.sum() * total)).astype(np.int64)
            tiny = rng.random(k) < 0.01
            ln[tiny] = rng.integers(1, 201, int(tiny.sum()))
        gc0 = rng.uniform(0.30, 0.70)
        g = np.full(len(ln), gc0)
        cont = rng.random(len(ln)) < 0.05
        g[cont] = np.clip(gc0 + rng.choice([-0.12, 0.12], int(cont.sum())), 0.05, 0.95)
        lens.append(ln)
        gcs.append(g)
        mag_off.append(mag_off[-1] + len(ln))
    return (np.concatenate(lens) if lens else np.zeros(0, np.int64),
            np.asarray(mag_off, dtype=np.int64),
            np.concatenate(gcs) if gcs else np.zeros(0))


def _decorate(bases, offsets, rng):
    """N runs, lower-case stretches and IUPAC bytes, in place (host arrays)."""
    L = len(bases)
    if L == 0:
        return
    for _ in range(int(rng.poisson(L / 1e5))):
        p = int(rng.integers(0, L))
        bases[p:p + int(rng.integers(1, 101))] = ord("N")
    n_low = int(rng.poisson(L * 0.02 / 500)) + 1
    for _ in range(n_low):
        p = int(rng.integers(0, L))
        bases[p:p + 500] |= 0x20
    for _ in range(int(rng.poisson(L / 1e6)) + 1):
        bases[int(rng.integers(0, L))] = ord("RYKMSW"[int(rng.integers(0, 6))])


def make_mags(n_mags, seed=0, long_contigs=False, mean_bp=3.0e6, n_contigs=300, decorate=True):
    """Host (NumPy) generator.  Returns dict(bases u8[L], offsets i64[n+1], mag_off, lengths)."""
    clen, mag_off, gc = mag_layout(n_mags, seed, long_contigs, mean_bp, n_contigs)
    offsets = np.zeros(len(clen) + 1, dtype=np.int64)
    np.cumsum(clen, out=offsets[1:])
    L = int(offsets[-1])
    rng = np.random.Generator(np.random.Philox(key=SEED_SEQ + 7919 * (seed + 1)))
    u = rng.random(L, dtype=np.float32)
    v = rng.random(L, dtype=np.float32) < 0.5
    gc_base = np.repeat(gc.astype(np.float32), clen)
    is_gc = u < gc_base
    bases = np.where(is_gc, np.where(v, ord("G"), ord("C")), np.where(v, ord("A"), ord("T"))).astype(np.uint8)
    if decorate:
        _decorate(bases, offsets, rng)
    return dict(bases=bases, offsets=offsets, mag_off=mag_off, lengths=clen)


def subset_layout(clen, mag_off, gc, mags):
    """The layout (mag_layout's triple) of the MAGs `mags` (indices) of a larger layout, in that order."""
    rows = np.concatenate([np.arange(mag_off[m], mag_off[m + 1]) for m in mags]) if len(mags) else np.zeros(0, np.int64)
    sizes = np.array([mag_off[m + 1] - mag_off[m] for m in mags], dtype=np.int64)
    return clen[rows], np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64), gc[rows]


def make_mags_device(n_mags, device, seed=0, long_contigs=False, mean_bp=3.0e6, n_contigs=300,
                     mags_per_block=64, layout=None):
    """Same model, bases generated on `device` with torch (bench sizes: 3 Gbp would take minutes
    on the host).  Returns dict(bases torch.uint8[L] on device, offsets/mag_off/lengths NumPy).
    layout: a (contig_len, mag_off, gc) triple to fill instead of mag_layout(n_mags, ...) (a rank's share of a
    larger MAG set)."""
    import torch

    clen, mag_off, gc = layout if layout is not None else mag_layout(n_mags, seed, long_contigs, mean_bp, n_contigs)
    n_mags = len(mag_off) - 1
    offsets = np.zeros(len(clen) + 1, dtype=np.int64)
    np.cumsum(clen, out=offsets[1:])
    L = int(offsets[-1])
    bases = torch.empty(L + 64, dtype=torch.uint8, device=device)[:L]
    gen = torch.Generator(device=device)
    gen.manual_seed(SEED_SEQ + 7919 * (seed + 1))
    lut = torch.tensor([ord("A"), ord("T"), ord("C"), ord("G")], dtype=torch.uint8, device=device)
    for m0 in range(0, n_mags, mags_per_block):
        m1 = min(n_mags, m0 + mags_per_block)
        c0, c1 = int(mag_off[m0]), int(mag_off[m1])
        b0, b1 = int(offsets[c0]), int(offsets[c1])
        if b1 == b0:
            continue
        gcb = torch.repeat_interleave(
            torch.from_numpy(gc[c0:c1].astype(np.float32)).to(device),
            torch.from_numpy(clen[c0:c1]).to(device), output_size=b1 - b0)
        u = torch.rand(b1 - b0, device=device, generator=gen)
        v = torch.rand(b1 - b0, device=device, generator=gen) < 0.5
        idx = (u < gcb).to(torch.int64) * 2 + v.to(torch.int64)
        bases[b0:b1] = lut[idx]
        del gcb, u, v, idx
    # decorations: N runs / lower case / IUPAC at host-chosen positions (few thousand tiny writes)
    rng = np.random.Generator(np.random.Philox(key=SEED_SEQ + 104729 * (seed + 1)))
    n_runs = int(rng.poisson(L / 1e5))
    if n_runs:
        p = torch.from_numpy(rng.integers(0, L, n_runs)).to(device)
        ln = torch.from_numpy(rng.integers(1, 101, n_runs)).to(device)
        for k in range(100):
            q = p + k
            ok = (ln > k) & (q < L)
            bases[q[ok]] = ord("N")
    n_low = int(rng.poisson(L * 0.02 / 500)) + 1
    p = torch.from_numpy(rng.integers(0, max(1, L - 512), n_low)).to(device)
    idx = (p[:, None] + torch.arange(500, device=device)[None, :]).reshape(-1)
    bases[idx] = bases[idx] | 0x20
    n_iu = int(rng.poisson(L / 1e6)) + 1
    p = torch.from_numpy(rng.integers(0, L, n_iu)).to(device)
    bases[p] = torch.from_numpy(np.frombuffer(b"RYKMSW", dtype=np.uint8)[rng.integers(0, 6, n_iu)].copy()).to(device)
    return dict(bases=bases, offsets=offsets, mag_off=mag_off, lengths=clen)


# ---- alignments --------------------------------------------------------------------------

def abundances(n_mags, n_samples, seed=0):
    """a[m, s] ~ LogNormal(ln 20, 1.0); 10 % of (MAG, sample) pairs at 0.2x."""
    rng = np.random.Generator(np.random.Philox(key=SEED_COV + 31 * (seed + 1)))
    a = rng.lognormal(np.log(20.0), 1.0, (n_mags, n_samples))
    a[rng.random((n_mags, n_samples)) < 0.10] *= 0.2 / 20.0
    return a


def make_events(lengths, mag_off, sample, seed=0, read_len=READ_LEN, variable=False, depth_scale=1.0, indel_frac=0.0):
    """Host generator of one sample's alignment blocks, grouped by contig, in BAM order (reads sorted by
    position inside a contig, the blocks of a read one after the other).  Returns (starts u32, ends u32,
    ev_off i64[n+1]).  variable: read length N(read_len, 15) clipped to [50, 250]; indel_frac: that share
    of the reads carries one insertion or deletion of 1-3 bp (SURVEY 8(d) variant), i.e. two M blocks --
    the second one starts AFTER the first blocks of the following reads, so the starts are not sorted."""
    n_mags = len(mag_off) - 1
    ab = abundances(n_mags, sample + 1, seed)[:, sample] * depth_scale
    rng = np.random.Generator(np.random.Philox(key=(SEED_COV ^ (sample << 20)) + (seed << 40)))
    per_contig = np.repeat(ab, np.diff(mag_off))
    cont = rng.random(len(lengths)) < 0.05
    per_contig[cont] = rng.lognormal(np.log(20.0), 1.0, int(cont.sum())) * depth_scale
    lam = per_contig * np.maximum(lengths - read_len + 1, 0) / read_len
    n_reads = rng.poisson(lam).astype(np.int64)
    n_reads[lengths < read_len] = 0
    ev_off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(n_reads, out=ev_off[1:])
    R = int(ev_off[-1])
    cid = np.repeat(np.arange(len(lengths), dtype=np.int64), n_reads)
    span = (lengths - read_len + 1)[cid]
    st = np.floor(rng.random(R) * span).astype(np.int64)
    if variable:
        rl = np.clip(np.rint(rng.normal(read_len, 15, R)), 50, 250).astype(np.int64)
    else:
        rl = np.full(R, read_len, dtype=np.int64)
    order = np.lexsort((st, cid))
    st, rl = st[order], rl[order]
    clen = lengths[cid]
    if indel_frac <= 0.0:
        en = np.minimum(st + rl, clen)  # a BAM never aligns past the contig end
        return st.astype(np.uint32), en.astype(np.uint32), ev_off
    two = rng.random(R) < indel_frac
    a = np.floor(rng.random(R) * (rl - 20)).astype(np.int64) + 10       # bases before the indel
    k = rng.integers(1, 4, R)                                            # its length
    dele = rng.random(R) < 0.5                                           # deletion: gap on the reference
    s2 = st + a + np.where(dele, k, 0)
    e2 = s2 + (rl - a - np.where(dele, 0, k))                            # insertion: k read bases are not aligned
    two &= (e2 <= clen) & (e2 > s2)
    nb = 1 + two.astype(np.int64)
    at = np.zeros(R + 1, dtype=np.int64)
    np.cumsum(nb, out=at[1:])
    S = np.empty(int(at[-1]), dtype=np.int64)
    E = np.empty(int(at[-1]), dtype=np.int64)
    S[at[:-1]] = st
    E[at[:-1]] = np.where(two, st + a, np.minimum(st + rl, clen))
    S[at[:-1][two] + 1] = s2[two]
    E[at[:-1][two] + 1] = e2[two]
    ev_off2 = at[ev_off]
    return S.astype(np.uint32), E.astype(np.uint32), ev_off2


def make_events_device(lengths, mag_off, sample, device, seed=0, read_len=READ_LEN, depth_scale=1.0, variant=False):
    """Device generator (bench sizes).  Same model, torch RNG; returns torch tensors
    (starts u32-as-int32 storage, ends, ev_off int64) on `device`.  variant: read length N(150, 15)
    clipped to [50, 250] and 2 % of the reads with one 1-3 bp insertion or deletion (two blocks)."""
    import torch

    n_mags = len(mag_off) - 1
    ab = abundances(n_mags, sample + 1, seed)[:, sample] * depth_scale
    rng = np.random.Generator(np.random.Philox(key=(SEED_COV ^ (sample << 20)) + (seed << 40)))
    per_contig = np.repeat(ab, np.diff(mag_off))
    cont = rng.random(len(lengths)) < 0.05
    per_contig[cont] = rng.lognormal(np.log(20.0), 1.0, int(cont.sum())) * depth_scale
    lam = per_contig * np.maximum(lengths - read_len + 1, 0) / read_len
    n_reads = rng.poisson(lam).astype(np.int64)
    n_reads[lengths < read_len] = 0
    ev_off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(n_reads, out=ev_off[1:])
    R = int(ev_off[-1])
    gen = torch.Generator(device=device)
    gen.manual_seed((SEED_COV ^ (sample << 20)) + seed)
    # contigs are handled in blocks of ~64 M reads so that the temporaries (64-bit sort keys) stay small next to
    # the result: a C4 shard has billions of reads per sample
    n_c = len(lengths)
    cuts = [0]
    while cuts[-1] < n_c:
        k = int(np.searchsorted(ev_off, ev_off[cuts[-1]] + (64 << 20), side="left"))
        cuts.append(min(n_c, max(k, cuts[-1] + 1)))
    n_out = None
    if variant:
        out_s, out_e, per_contig_blocks = [], [], []
    else:
        starts = torch.empty(R, dtype=torch.int32, device=device)
        ends = torch.empty(R, dtype=torch.int32, device=device)
    for c0, c1 in zip(cuts[:-1], cuts[1:]):
        r0, r1 = int(ev_off[c0]), int(ev_off[c1])
        Rb = r1 - r0
        if Rb == 0:
            if variant:
                per_contig_blocks.append(np.zeros(c1 - c0, dtype=np.int64))
            continue
        d_n = torch.from_numpy(n_reads[c0:c1]).to(device)
        cid = torch.repeat_interleave(torch.arange(c1 - c0, device=device, dtype=torch.int64), d_n, output_size=Rb)
        span = torch.from_numpy((lengths[c0:c1] - read_len + 1)).to(device)[cid]
        st = torch.floor(torch.rand(Rb, device=device, generator=gen, dtype=torch.float64) * span).to(torch.int64)
        key, _ = torch.sort((cid << 32) | st)
        del cid, span
        st = key & 0xFFFFFFFF
        if not variant:
            starts[r0:r1] = st.to(torch.int32)  # bit pattern == uint32 (contigs are < 2^31)
            ends[r0:r1] = (st + read_len).to(torch.int32)
            del key, st
            continue
        cid = key >> 32
        del key
        clen = torch.from_numpy(lengths[c0:c1].astype(np.int64)).to(device)[cid]
        rl = torch.clamp(torch.round(torch.randn(Rb, device=device, generator=gen) * 15.0 + read_len), 50, 250).to(torch.int64)
        two = torch.rand(Rb, device=device, generator=gen) < 0.02
        a = torch.floor(torch.rand(Rb, device=device, generator=gen) * (rl - 20)).to(torch.int64) + 10
        k = torch.randint(1, 4, (Rb,), device=device, generator=gen)
        dele = torch.rand(Rb, device=device, generator=gen) < 0.5
        s2 = st + a + torch.where(dele, k, torch.zeros_like(k))
        e2 = s2 + (rl - a - torch.where(dele, torch.zeros_like(k), k))
        two &= (e2 <= clen) & (e2 > s2)
        nb = 1 + two.to(torch.int64)
        at = torch.zeros(Rb + 1, dtype=torch.int64, device=device)
        torch.cumsum(nb, 0, out=at[1:])
        nblk = int(at[-1].item())
        S = torch.empty(nblk, dtype=torch.int32, device=device)
        E = torch.empty(nblk, dtype=torch.int32, device=device)
        first = at[:-1]
        S[first] = st.to(torch.int32)
        E[first] = torch.where(two, st + a, torch.minimum(st + rl, clen)).to(torch.int32)
        S[first[two] + 1] = s2[two].to(torch.int32)
        E[first[two] + 1] = e2[two].to(torch.int32)
        out_s.append(S)
        out_e.append(E)
        per_contig_blocks.append(torch.bincount(cid, weights=nb.to(torch.float64), minlength=c1 - c0).to(torch.int64).cpu().numpy())
        del cid, clen, rl, two, a, k, dele, s2, e2, nb, at, first, st
    if not variant:
        return starts, ends, torch.from_numpy(ev_off).to(device), R
    S = torch.cat(out_s) if out_s else torch.empty(0, dtype=torch.int32, device=device)
    E = torch.cat(out_e) if out_e else torch.empty(0, dtype=torch.int32, device=device)
    blocks = np.concatenate(per_contig_blocks) if per_contig_blocks else np.zeros(0, np.int64)
    ev_off2 = np.concatenate([[0], np.cumsum(blocks)]).astype(np.int64)
    return S, E, torch.from_numpy(ev_off2).to(device), int(S.numel())

