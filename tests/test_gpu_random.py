"""GPU: randomised differential test -- small batches with arbitrary CIGARs (M I D N S H P = X),
every flag combination the filters look at, reads without sequence, IUPAC / '=' codes in reads
and reference, overlapping mates (same QNAME) and deep same-start groups, pushed through
EVERY statistic on both kernels and compared bit-exactly with the CPU oracle.
"""
import numpy as np
import pytest

from pysamstats_b200 import _cabi, config
from oracle import oracle as orc
from _util import oracle_reads_from_arrays

pytestmark = pytest.mark.gpu

M, I, D, N, S, H, P, EQ, X = range(9)


def random_cigar(rng, maxlen):
    """A CIGAR htslib's pileup accepts: optional H/S clips at the ends, a first and last
    aligned op, anything in between (never two adjacent ops of the same kind)."""
    ops = []
    if rng.random() < 0.15:
        ops.append((H, int(rng.integers(1, 20))))
    if rng.random() < 0.25:
        ops.append((S, int(rng.integers(1, 15))))
    n_mid = int(rng.choice([0, 0, 0, 1, 1, 2, 3, 5]))
    body = [(int(rng.choice([M, M, M, EQ, X])), int(rng.integers(1, maxlen)))]
    for _ in range(n_mid):
        k = int(rng.choice([I, D, N, P, I, D]))
        ln = int(rng.integers(1, 12)) if k != N else int(rng.integers(1, 300))
        if k == P and rng.random() < 0.7:          # pad followed by an insertion (htslib sums those)
            body.append((P, ln))
            body.append((I, int(rng.integers(1, 6))))
        else:
            body.append((k, ln))
        body.append((int(rng.choice([M, M, EQ, X])), int(rng.integers(1, maxlen))))
    ops += body
    if rng.random() < 0.25:
        ops.append((S, int(rng.integers(1, 15))))
    if rng.random() < 0.1:
        ops.append((H, int(rng.integers(1, 20))))
    return ops


def random_batch(seed, n_reads=600, span=3000, maxlen=60, deep=False):
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.integers(50, span, n_reads)).astype('i4')
    if deep:                                       # many reads sharing a few starts
        pos = np.sort(rng.choice(rng.integers(50, span, 6), n_reads)).astype('i4')
    flag = np.zeros(n_reads, 'u2')
    cig, cig_off, seqs, seq_off, quals, qual_off, l_seq = [], [0], [], [0], [], [0], []
    names = rng.integers(0, n_reads // 2, n_reads)            # QNAMEs shared by ~2 reads
    tlen = np.zeros(n_reads, 'i4')
    mtid = np.zeros(n_reads, 'i4')
    mapq = rng.choice([0, 1, 5, 20, 30, 60, 255], n_reads).astype('u1')
    for i in range(n_reads):
        f = 0
        r = rng.random()
        if r < 0.8:
            f |= 1
            if rng.random() < 0.6:
                f |= 2
            if rng.random() < 0.08:
                f |= 8
            if rng.random() < 0.5:
                f |= 32
            f |= 64 if rng.random() < 0.5 else 128
            mtid[i] = 0 if rng.random() < 0.85 else int(rng.choice([-1, 1]))
            tlen[i] = int(rng.choice([0, 40, -40, 90, -90, 150, -150, 400, -400, 5000]))
        else:
            mtid[i] = -1
        if rng.random() < 0.5:
            f |= 16
        if rng.random() < 0.04:
            f |= 4
        if rng.random() < 0.04:
            f |= 256
        if rng.random() < 0.03:
            f |= 512
        if rng.random() < 0.06:
            f |= 1024
        flag[i] = f
        ops = random_cigar(rng, maxlen) if not (f & 4 and rng.random() < 0.5) else []
        cig += [(ln << 4) | op for op, ln in ops]
        cig_off.append(len(cig))
        qlen = sum(ln for op, ln in ops if op in (M, I, S, EQ, X))
        if rng.random() < 0.03 or not ops:
            qlen = 0 if rng.random() < 0.7 else qlen          # SEQ '*'
        l_seq.append(qlen)
        codes = rng.choice([1, 2, 4, 8, 1, 2, 4, 8, 15, 0, 3, 5, 10], qlen).astype('u1')
        if qlen % 2:
            codes = np.append(codes, 0).astype('u1')
        seqs.append(((codes[0::2] << 4) | codes[1::2]).astype('u1'))
        seq_off.append(seq_off[-1] + len(seqs[-1]))
        q = rng.integers(0, 42, qlen).astype('u1')
        if rng.random() < 0.03:
            q[:] = 255                                         # QUAL '*'
        quals.append(q)
        qual_off.append(qual_off[-1] + qlen)
    last = {}
    prev = np.full(n_reads, -1, 'i4')
    for i, nm in enumerate(names):
        prev[i] = last.get(nm, -1)
        last[nm] = i
    cat = lambda parts, dt: (np.concatenate(parts).astype(dt) if sum(len(p) for p in parts) else np.zeros(1, dt))
    nb = dict(n_reads=n_reads, tid=0, pos=pos, flag=flag, mapq=mapq, tlen=tlen, mtid=mtid,
              l_seq=np.array(l_seq, 'i4'), cigar=np.array(cig if cig else [0], 'u4'),
              cigar_off=np.array(cig_off, 'i8'), seq4=cat(seqs, 'u1'), seq_off=np.array(seq_off, 'i8'),
              qual=cat(quals, 'u1'), qual_off=np.array(qual_off, 'i8'), name_prev=prev)
    ref = rng.choice(list(b'ACGTacgtNnRYM=*'), span + 400, p=[.2, .2, .2, .2, .04, .04, .03, .03, .01, .01, .01, .01, .01,
                                                            .005, .005]).astype('u1')
    return nb, ref


def run(ctx, nb, ref, stat, lo, hi, flags=0, **kw):
    need_qual = bool(ctx.lib.pss_pileup_needs_qual(_cabi.PILEUP_STATS.index(stat), int(kw.get('min_baseq', 0))))
    d = dict(nb)
    if not need_qual:
        d['qual'] = d['qual_off'] = d['name_prev'] = None
    withref = stat in config.stats_types_withref
    p = ctx.pileup_params(stat, start=lo, end=hi, ref=ref if withref else None, ref_start=0,
                          ref_len=len(ref) if withref else 0, contig_len=len(ref), flags=flags, **kw)
    out = np.empty((len(config.value_fields(stat)), hi - lo), dtype=np.int32)
    mask = np.empty(hi - lo, dtype=np.uint8)
    ctx.pileup_host(_cabi.make_reads_struct(d), p, out, mask)
    return out, mask


def check(ctx, nb, ref, reads, stat, lo, hi, **kw):
    withref = stat in config.stats_types_withref
    pos, F, _ = orc.pileup_packed(stat, reads, lo, hi, ref=ref.tobytes() if withref else None, truncate=True, **kw)
    for flags in ((0, _cabi.PSS_FLAG_GENERIC_KERNEL, _cabi.PSS_FLAG_LONG_READS) if stat != 'coverage_gc' else (0,)):
        out, mask = run(ctx, nb, ref, stat, lo, hi, flags=flags, **kw)
        rows = np.flatnonzero(mask)
        label = '%s %r flags=%d' % (stat, kw, flags)
        assert np.array_equal(rows + lo, pos), label
        got = out[:, rows].T
        if stat == 'coverage_gc':
            got, F = got[:, 1:], F[:, 1:]
        if not np.array_equal(got, F):
            bad = np.argwhere(got != F)[0]
            raise AssertionError('%s: pos %d field %s: got %d expect %d' % (
                label, pos[bad[0]], config.value_fields(stat)[bad[1] + (stat == 'coverage_gc')], got[tuple(bad)],
                F[tuple(bad)]))


CASES = (dict(), dict(no_del=True), dict(min_mapq=5, no_dup=True), dict(stepper='nofilter'),
         dict(stepper='samtools', min_baseq=13), dict(min_baseq=25, no_del=True), dict(max_depth=7),
         dict(stepper='nofilter', max_depth=3, no_dup=True))


@pytest.mark.parametrize('seed', [1, 2, 3])
@pytest.mark.parametrize('stat', [s for s in config.pileup_types])
def test_random_batches(stat, seed):
    ctx = _cabi.get_context()
    nb, ref = random_batch(seed, deep=(seed == 3))
    reads = oracle_reads_from_arrays(nb)
    hi = int(nb['pos'].max()) + 900
    for kw in CASES:
        check(ctx, nb, ref, reads, stat, 0, hi, **kw)
    check(ctx, nb, ref, reads, stat, 777, 1999)          # odd range inside the data


def _read_batch(nb):
    from pysamstats_b200 import bamio
    b = bamio.ReadBatch()
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'cigar', 'seq_off', 'seq4', 'qual_off', 'qual'):
        setattr(b, k, nb[k])
    b.name_prev = nb['name_prev']
    b.endpos = np.array([orc_endpos(nb, i) for i in range(nb['n_reads'])], dtype='i4')
    b.tid, b.names = 0, None
    return b


@pytest.mark.parametrize('seed', [11, 12])
def test_quality_requests_in_position_pieces(seed, monkeypatch):
    # a request that reads base qualities is cut into position pieces: the pairs of the mate-overlap rewrite are
    # computed once for the whole batch from its metadata (pss_pairs_host) and every piece receives them as
    # two-read name chains in its own indices (opt._position_chunks) -- the pieces side by side must be the single
    # call, and the pairs must be the ones the whole-batch rewrite forms (pss_overlap_tweak).  QNAMEs shared by two,
    # three and more reads, arbitrary CIGARs / flags, dense enough that hundreds of mates overlap
    import torch
    from pysamstats_b200 import opt
    ctx = _cabi.get_context()
    nb, ref = random_batch(seed, n_reads=4000, span=6000, maxlen=70)
    # QNAMEs shared by two or three NEARBY reads (a stored read expires from htslib's table once the pileup has moved
    # past it), most of them qualifying for the rewrite (proper pair, mate mapped, |tlen| < 2 l_seq)
    rng = np.random.default_rng(seed)
    n = nb['n_reads']
    order = np.concatenate([k + rng.permutation(min(12, n - k)) for k in range(0, n, 12)])
    names = np.empty(n, 'i8')
    i = g = 0
    while i < n:
        k = int(rng.choice([2, 2, 3, 1]))
        names[order[i:i + k]] = g
        i, g = i + k, g + 1
    last, prev = {}, np.full(n, -1, 'i4')
    for i, nm in enumerate(names.tolist()):
        prev[i] = last.get(nm, -1)
        last[nm] = i
    ovl = rng.random(n) < 0.8
    nb['name_prev'] = prev
    nb['flag'] = np.where(ovl, (nb['flag'] | 3) & ~np.uint16(8), nb['flag']).astype('u2')
    nb['tlen'] = np.where(ovl, rng.choice([40, -40, 25, -90], n), nb['tlen']).astype('i4')
    b = _read_batch(nb)
    monkeypatch.setenv('PSS_TABLE_CHUNK', '1024')
    lo, hi = 0, int(nb['pos'].max()) + 900
    for stepper in ('all', 'nofilter'):
        links = ctx.pairs_host(_cabi.make_reads_struct(nb), stepper, 8000)
        dev = torch.device('cuda', ctx.device)
        dn = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v) for k, v in nb.items()}
        qout = torch.empty(len(nb['qual']), dtype=torch.uint8, device=dev)
        pf = torch.empty(nb['n_reads'], dtype=torch.int32, device=dev)
        ctx.overlap_tweak(_cabi.make_reads_struct(dn), _cabi.STEPPERS[stepper], 8000, qout, pf)
        ctx.synchronize()
        assert np.array_equal(links, pf.cpu().numpy()) and (links >= 0).sum() > 50, stepper
    for stat, kw in (('baseq_strand', dict()), ('baseq_ext', dict(stepper='nofilter')), ('mapq', dict(min_baseq=20)),
                     ('tlen_strand', dict(min_baseq=13, no_del=True)), ('variation', dict(min_baseq=30)),
                     ('baseq_ext_strand', dict(min_baseq=10, no_dup=True))):
        whole, wmask = run(ctx, nb, ref, stat, lo, hi, **kw)
        pieces = list(opt._position_chunks(b, lo, hi, True, 8000, kw.get('stepper', 'all')))
        assert len(pieces) >= 6
        withref = stat in config.stats_types_withref
        for c_lo, c_hi, rs in pieces:
            p = ctx.pileup_params(stat, start=c_lo, end=c_hi, ref=ref if withref else None, ref_start=0,
                                  ref_len=len(ref) if withref else 0, contig_len=len(ref), **kw)
            out = np.empty((whole.shape[0], c_hi - c_lo), dtype=np.int32)
            mask = np.empty(c_hi - c_lo, dtype=np.uint8)
            ctx.pileup_host(rs, p, out, mask)
            assert np.array_equal(mask, wmask[c_lo - lo:c_hi - lo]), (stat, kw, c_lo)
            rows = np.flatnonzero(mask)
            assert np.array_equal(out[:, rows], whole[:, c_lo - lo:c_hi - lo][:, rows]), (stat, kw, c_lo)


def test_random_binned():
    ctx = _cabi.get_context()
    for seed in (4, 5):
        nb, ref = random_batch(seed, n_reads=2000, span=20000)
        reads = oracle_reads_from_arrays(nb)
        d = dict(nb)
        d['qual'] = d['qual_off'] = d['name_prev'] = None
        for stat in config.binned_types:
            for (start, end, W, kw) in ((0, 21000, 300, {}), (1000, 15000, 77, dict(min_mapq=20, no_dup=True)),
                                        (5000, 5100, 500, {})):
                # the region iterator only yields overlapping reads
                endpos = np.array([orc_endpos(nb, i) for i in range(nb['n_reads'])])
                keep = np.flatnonzero((nb['pos'] < end) & (endpos > start))
                sub = subset(nb, keep)
                p = ctx.binned_params(stat, start, end, W, ref=ref, ref_start=0, ref_len=len(ref), **kw)
                out = np.empty((len(config.value_fields_binned(stat)), int(p.n_bins)), dtype=np.int32)
                s2 = dict(sub)
                s2['qual'] = s2['qual_off'] = s2['name_prev'] = None
                ctx.binned_host(_cabi.make_reads_struct(s2), p, out)
                pos, F = orc.binned_packed(stat, oracle_reads_from_arrays(sub), start, end, W, W // 2, ref=ref.tobytes(),
                                           **kw)
                assert F.shape == out.T.shape, (stat, start, end, W)
                assert np.array_equal(out.T, F), (stat, start, end, W, kw)


def orc_endpos(nb, i):
    if nb['flag'][i] & 4:
        return int(nb['pos'][i]) + 1
    c = nb['cigar'][nb['cigar_off'][i]:nb['cigar_off'][i + 1]]
    rl = int(sum(int(w >> 4) for w in c if (w & 15) in (M, D, N, EQ, X)))
    return int(nb['pos'][i]) + (rl if rl else 1)


def subset(nb, idx):
    out = dict(n_reads=len(idx), tid=0)
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq'):
        out[k] = np.ascontiguousarray(nb[k][idx])
    for data, off in (('cigar', 'cigar_off'), ('seq4', 'seq_off'), ('qual', 'qual_off')):
        parts = [nb[data][nb[off][i]:nb[off][i + 1]] for i in idx]
        o = np.zeros(len(idx) + 1, 'i8')
        if len(idx):
            o[1:] = np.cumsum([len(p) for p in parts])
        out[data] = np.concatenate(parts).astype(nb[data].dtype) if o[-1] else np.zeros(1, nb[data].dtype)
        out[off] = o
    old2new = {int(j): k for k, j in enumerate(idx)}
    out['name_prev'] = np.array([old2new.get(int(nb['name_prev'][j]), -1) for j in idx], 'i4')
    return out
