"""Synthetic WGS-shaped read batches (input "S1").

Used by ``bench.py`` (on the GPU, 250 Mb x 30x) and by the parity tests (small, on the CPU
through numpy).  Produces the structure-of-arrays batch of ``include/pss_b200.h`` directly
-- the BAM container is not part of the timed path.  torch is used as the array library
so the same code fills device memory for the bench and host memory for tests.

Recipe per pair: leftmost start uniform, insert size ~ N(400, 60) clipped to [L, 700], FR
orientation (flags 99/147 or 163/83), 1 % mate-unmapped (the mate is kept as an
unmapped-placed record), 1 % mate on another contig, 1 % same-strand, 1 % face-away, 2 %
duplicates, mapq 85 % 60 / 5 % 0 / rest uniform 1-59, CIGAR 92 % ``LM``, 3 % one 1-10 bp
insertion, 3 % one 1-10 bp deletion, 2 % soft-clipped end, bases = reference with 0.5 %
substitutions and 0.1 % N, qualities ~ N(35, 5) clipped to [2, 41].
"""
import numpy as np
import torch

_CODE = torch.tensor([1, 2, 4, 8], dtype=torch.uint8)     # A C G T in BAM 4-bit
_ASCII = torch.tensor([65, 67, 71, 84], dtype=torch.uint8)


def make_reference(length, seed=1, device='cpu'):
    """Random ACGT reference with ~0.1 % of N runs; returns (ascii uint8 tensor, 2-bit codes)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    base = torch.randint(0, 4, (length,), generator=g, device=device, dtype=torch.uint8)
    ascii_ = _ASCII.to(device)[base.long()]
    n_runs = max(1, length // 50000)
    starts = torch.randint(0, max(1, length - 50), (n_runs,), generator=g, device=device)
    idx = (starts[:, None] + torch.arange(50, device=device)[None, :]).reshape(-1)
    idx = idx[idx < length]
    is_n = torch.zeros(length, dtype=torch.bool, device=device)
    is_n[idx] = True
    ascii_[is_n] = 78
    return ascii_, base, is_n


def make_batch(contig_len, n_pairs, read_len=150, seed=20261019, device='cpu', ref=None, chunk_pairs=1 << 19,
               with_qual=True, lo=0, hi=None):
    """Generate a coordinate-sorted batch of 2 * n_pairs reads starting in [lo, hi).
    Returns a dict of torch tensors (the pss_reads arrays + sizes) and the reference."""
    dev = torch.device(device)
    if ref is None:
        ref = make_reference(contig_len, seed=1, device=device)
    ref_ascii, ref_base, ref_is_n = ref
    hi = contig_len if hi is None else hi
    L = read_len
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    parts = []
    done = 0
    while done < n_pairs:
        m = min(chunk_pairs, n_pairs - done)
        parts.append(_make_chunk(m, done, L, lo, max(lo + 1, hi - 700 - 20), ref_base, 0, ref_is_n, g, dev, with_qual))
        done += m
    cat = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
    del parts
    return _finish(cat, L, dev, with_qual, None, None), ref


def _finish(cat, L, dev, with_qual, lo, hi):
    """Coordinate sort (stable: equal starts keep generation order, like a sorted BAM), keep
    the records overlapping [lo, hi) (htslib's region rule: pos < hi and endpos > lo) and
    lay the batch out as the pss_reads arrays."""
    order = torch.argsort(cat['pos'], stable=True)
    n_all = order.numel()
    pos_s = cat['pos'][order]
    if lo is None:
        keep = torch.ones(n_all, dtype=torch.bool, device=dev)
    else:
        end_s = pos_s.long() + cat['reflen'][order].long()
        keep = (pos_s.long() < hi) & (end_s > lo)
    newidx = torch.cumsum(keep.long(), 0) - 1
    # name_prev: the earlier record of the same pair, if it is in the batch
    inv = torch.empty(n_all, dtype=torch.int64, device=dev)
    inv[order] = torch.arange(n_all, device=dev)
    mate_sorted = inv[order ^ 1]                                # sorted index of each read's mate
    me = torch.arange(n_all, device=dev)
    has_prev = (mate_sorted < me) & keep[mate_sorted]
    prev = torch.where(has_prev, newidx[mate_sorted], torch.full_like(me, -1))
    order = order[keep]
    n = order.numel()
    out = {}
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq'):
        out[k] = cat[k][order].contiguous()
    ncig = cat['ncig'][order]
    out['cigar_off'] = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    out['cigar_off'][1:] = torch.cumsum(ncig.long(), 0)
    cig = cat['cig3'][order]                                     # [n, 3]
    kmask = torch.arange(3, device=dev)[None, :] < ncig[:, None]
    out['cigar'] = cig[kmask].contiguous()
    nb = (L + 1) // 2
    out['seq_off'] = torch.arange(n + 1, dtype=torch.int64, device=dev) * nb
    out['seq4'] = cat['seq4'][order].reshape(-1).contiguous()
    if with_qual:
        out['qual_off'] = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
        out['qual'] = cat['qual'][order].reshape(-1).contiguous()
    else:
        out['qual_off'] = None
        out['qual'] = None
    out['name_prev'] = prev[keep].to(torch.int32)
    out['endpos'] = (out['pos'].long() + cat['reflen'][order].long())
    out['n_reads'] = n
    out['n_cigar_words'] = int(out['cigar'].numel())
    out['n_seq_bytes'] = int(out['seq4'].numel())
    out['n_qual_bytes'] = int(out['qual'].numel()) if with_qual else 0
    out['tid'] = 0
    return out


def reference_span(lo, hi, seed=1, device='cpu', block=1000000):
    """Reference bytes of [lo, hi) of the synthetic genome.  The genome is defined block by
    block (``block`` positions, generator seeded by the block number), so any rank can
    materialise any interval and neighbouring shards agree on their shared positions.
    Returns (ascii uint8, 2-bit base, is_N), each of length hi - lo."""
    b0, b1 = lo // block, (hi + block - 1) // block
    parts = [make_reference(block, seed=seed * 1000003 + b, device=device) for b in range(b0, max(b1, b0 + 1))]
    cat = [torch.cat([p[i] for p in parts]) for i in range(3)]
    o = lo - b0 * block
    return tuple(c[o:o + (hi - lo)].contiguous() for c in cat)


def make_interval_batch(lo, hi, genome_len, depth=30, read_len=150, seed=20261019, device='cpu', with_qual=True,
                        block=1000000):
    """The records of a ``depth``-fold synthetic WGS run over a genome of ``genome_len``
    positions that overlap [lo, hi) (so including the halo of reads that start before lo),
    coordinate sorted -- exactly what a region query of the sorted BAM would return.

    Pairs are generated per block of leftmost starts with a generator seeded by the block
    number: every shard of a multi-GPU run sees the same genome-wide read set no matter how
    the genome is cut.  Returns (batch dict, (ref_ascii, ref_start)) with the reference
    covering [ref_start, ref_start + len(ref_ascii)) >= [lo, hi)."""
    dev = torch.device(device)
    L = read_len
    pairs_per_block = int(round(depth * block / (2.0 * L)))
    b0 = max(0, (lo - 720 - L) // block)         # a pair spans <= 700 + deletion <= 720
    b1 = min((hi + block - 1) // block, (genome_len + block - 1) // block)
    ref_lo = b0 * block
    ref_hi = min(genome_len, (b1 + 1) * block)    # reads of the last block reach < b1 * block + 720
    ref_ascii, ref_base, ref_is_n = reference_span(ref_lo, ref_hi, seed=1, device=device, block=block)
    parts = []
    for b in range(b0, b1):
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1000003 + b)
        s_lo = b * block
        s_hi = min((b + 1) * block, genome_len - 720)
        if s_hi <= s_lo:
            continue
        parts.append(_make_chunk(pairs_per_block, 0, L, s_lo, s_hi, ref_base, ref_lo, ref_is_n, g, dev,
                                 with_qual))
    cat = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
    del parts
    out = _finish(cat, L, dev, with_qual, lo, hi)
    return out, (ref_ascii, ref_lo)


def _make_chunk(m, pair0, L, lo, hi, ref_base, ref_lo, ref_is_n, g, dev, with_qual):
    def rnd(*shape):
        return torch.rand(*shape, generator=g, device=dev)

    def rint(a, b, *shape):
        return torch.randint(a, b, shape, generator=g, device=dev)

    start = rint(lo, hi, m)
    isize = torch.clamp(torch.round(400 + 60 * torch.randn(m, generator=g, device=dev)), L, 700).long()
    cls = rnd(m)                  # pair class
    left_is_r1 = rnd(m) < 0.5
    dup = rnd(m) < 0.02
    # records are generated as (left, right) interleaved: index 2i, 2i+1
    n = 2 * m
    pos = torch.empty(n, dtype=torch.int64, device=dev)
    pos[0::2] = start
    pos[1::2] = start + isize - L
    rev = torch.zeros(n, dtype=torch.bool, device=dev)
    rev[1::2] = True
    mrev = torch.zeros(n, dtype=torch.bool, device=dev)
    mrev[0::2] = True
    flag = torch.ones(n, dtype=torch.int64, device=dev)          # paired
    r1 = torch.empty(n, dtype=torch.bool, device=dev)
    r1[0::2] = left_is_r1
    r1[1::2] = ~left_is_r1
    flag |= torch.where(r1, 64, 128)
    tlen = torch.empty(n, dtype=torch.int64, device=dev)
    tlen[0::2] = isize
    tlen[1::2] = -isize
    mtid = torch.zeros(n, dtype=torch.int64, device=dev)
    pcls = cls.repeat_interleave(2)
    proper = pcls >= 0.04
    mate_unmapped = pcls < 0.01              # right record becomes an unmapped-placed read
    other_chr = (pcls >= 0.01) & (pcls < 0.02)
    same_strand = (pcls >= 0.02) & (pcls < 0.03)
    faceaway = (pcls >= 0.03) & (pcls < 0.04)
    is_right = torch.zeros(n, dtype=torch.bool, device=dev)
    is_right[1::2] = True
    # same strand: both forward
    rev = torch.where(same_strand, torch.zeros_like(rev), rev)
    mrev = torch.where(same_strand, torch.zeros_like(mrev), mrev)
    # face-away: left read reverse, right read forward (RF)
    rev = torch.where(faceaway, ~is_right, rev)
    mrev = torch.where(faceaway, is_right, mrev)
    mtid = torch.where(other_chr, torch.ones_like(mtid), mtid)
    tlen = torch.where(other_chr | mate_unmapped, torch.zeros_like(tlen), tlen)
    unmapped = mate_unmapped & is_right
    pos = torch.where(unmapped, pos[0::2].repeat_interleave(2), pos)   # placed at the mate's position
    flag |= torch.where(proper, 2, 0)
    flag |= torch.where(rev & ~unmapped, 16, 0)
    flag |= torch.where(mrev & ~mate_unmapped, 32, 0)
    flag |= torch.where(mate_unmapped & ~is_right, 8, 0)
    flag |= torch.where(unmapped, 4, 0)
    flag |= torch.where(dup.repeat_interleave(2), 1024, 0)
    mq = rnd(n)
    mapq = torch.where(mq < 0.85, torch.full((n,), 60, device=dev),
                       torch.where(mq < 0.90, torch.zeros(n, dtype=torch.int64, device=dev), rint(1, 60, n)))
    mapq = torch.where(unmapped, torch.zeros_like(mapq), mapq)
    # CIGAR classes
    cg = rnd(n)
    k = rint(1, 11, n)                                           # indel / clip length
    a = rint(10, L - 20, n)                                      # bases before the event
    typ = torch.zeros(n, dtype=torch.int64, device=dev)          # 0 M, 1 I, 2 D, 3 S-left, 4 S-right
    typ = torch.where((cg >= 0.92) & (cg < 0.95), 1, typ)
    typ = torch.where((cg >= 0.95) & (cg < 0.98), 2, typ)
    typ = torch.where((cg >= 0.98) & (cg < 0.99), 3, typ)
    typ = torch.where(cg >= 0.99, 4, typ)
    typ = torch.where(unmapped, torch.zeros_like(typ), typ)
    M, I, D, S = 0, 1, 2, 4
    cig3 = torch.zeros(n, 3, dtype=torch.int64, device=dev)
    ncig = torch.ones(n, dtype=torch.int64, device=dev)
    cig3[:, 0] = (L << 4) | M
    sel = typ == 1
    cig3[sel, 0] = (a[sel] << 4) | M
    cig3[sel, 1] = (k[sel] << 4) | I
    cig3[sel, 2] = ((L - a[sel] - k[sel]) << 4) | M
    ncig[sel] = 3
    sel = typ == 2
    cig3[sel, 0] = (a[sel] << 4) | M
    cig3[sel, 1] = (k[sel] << 4) | D
    cig3[sel, 2] = ((L - a[sel]) << 4) | M
    ncig[sel] = 3
    sel = typ == 3
    cig3[sel, 0] = (k[sel] << 4) | S
    cig3[sel, 1] = ((L - k[sel]) << 4) | M
    ncig[sel] = 2
    sel = typ == 4
    cig3[sel, 0] = ((L - k[sel]) << 4) | M
    cig3[sel, 1] = (k[sel] << 4) | S
    ncig[sel] = 2
    ncig = torch.where(unmapped, torch.zeros_like(ncig), ncig)   # unmapped: no CIGAR
    # query -> reference offset map, -1 = not templated by the reference
    q = torch.arange(L, device=dev)[None, :].expand(n, L)
    a_, k_ = a[:, None], k[:, None]
    t_ = typ[:, None]
    off = q.clone()
    off = torch.where((t_ == 1) & (q >= a_ + k_), q - k_, off)
    off = torch.where((t_ == 1) & (q >= a_) & (q < a_ + k_), torch.full_like(off, -1), off)
    off = torch.where((t_ == 2) & (q >= a_), q + k_, off)
    off = torch.where((t_ == 3), torch.where(q < k_, torch.full_like(off, -1), q - k_), off)
    off = torch.where((t_ == 4) & (q >= L - k_), torch.full_like(off, -1), off)
    rpos = pos[:, None] + off - ref_lo                           # index into the reference span held
    templ = (off >= 0) & (rpos >= 0) & (rpos < ref_base.numel())
    rb = ref_base[torch.clamp(rpos, 0, ref_base.numel() - 1)].long()
    rn = ref_is_n[torch.clamp(rpos, 0, ref_base.numel() - 1)]
    rand_base = rint(0, 4, n, L)
    base = torch.where(templ, rb, rand_base)
    sub = rnd(n, L) < 0.005
    base = torch.where(sub, (base + rint(1, 4, n, L)) % 4, base)
    code = _CODE.to(dev)[base]
    isn = (rnd(n, L) < 0.001) | (templ & rn)
    code = torch.where(isn, torch.full_like(code, 15), code)
    if L % 2:
        code = torch.cat([code, torch.zeros(n, 1, dtype=torch.uint8, device=dev)], 1)
    seq4 = (code[:, 0::2] << 4) | code[:, 1::2]
    # flag / cigar are kept as int16 / int32: bit-identical to the ABI's uint16 / uint32 for
    # the values generated here, and fully supported by torch indexing
    reflen = torch.full((n,), L, dtype=torch.int64, device=dev)
    reflen = torch.where(typ == 1, L - k, reflen)
    reflen = torch.where(typ == 2, L + k, reflen)
    reflen = torch.where((typ == 3) | (typ == 4), L - k, reflen)
    reflen = torch.where(unmapped, torch.ones_like(reflen), reflen)   # bam_endpos(): pos + 1
    out = {'pos': pos.to(torch.int32), 'flag': flag.to(torch.int16), 'mapq': mapq.to(torch.uint8),
           'reflen': reflen.to(torch.int32),
           'tlen': tlen.to(torch.int32), 'mtid': mtid.to(torch.int32),
           'l_seq': torch.full((n,), L, dtype=torch.int32, device=dev), 'ncig': ncig,
           'cig3': cig3.to(torch.int32), 'seq4': seq4}
    if with_qual:
        qual = torch.clamp(torch.round(35 + 5 * torch.randn(n, L, generator=g, device=dev)), 2, 41)
        out['qual'] = qual.to(torch.uint8)
    return out


def batch_to_numpy(batch):
    """Host numpy copy of a generated batch (for the oracle / the *_host entry points)."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            a = v.cpu().numpy()
            if k == 'flag':
                a = a.view(np.uint16)
            elif k == 'cigar':
                a = a.view(np.uint32)
            out[k] = a
        else:
            out[k] = v
    return out
