"""GPU: synthetic WGS-shaped batches straight through the C ABI (device and host entry points),
compared bit-exactly with the CPU oracle on the same arrays, plus size-independent
properties at bench scale (linearity of the counts, shard concatenation, kernel agreement).
"""
import numpy as np
import pytest

from pysamstats_b200 import _cabi, config, shard, synth
from oracle import oracle as orc
from _util import oracle_reads_from_arrays

pytestmark = pytest.mark.gpu

G = 300_000          # oracle-checkable in about a second per statistic
TILE_STATS = ('coverage', 'coverage_strand', 'coverage_ext', 'coverage_ext_strand', 'variation', 'variation_strand',
              'mapq', 'mapq_strand', 'baseq', 'baseq_strand', 'baseq_ext', 'baseq_ext_strand', 'tlen', 'tlen_strand')


@pytest.fixture(scope='module')
def ctx():
    return _cabi.get_context()


@pytest.fixture(scope='module')
def small():
    batch, (ref, ref_lo) = synth.make_interval_batch(0, G, G, depth=30, device='cpu', with_qual=True, block=100000)
    nb = synth.batch_to_numpy(batch)
    return nb, ref.numpy(), oracle_reads_from_arrays(nb)


def _run_host(ctx, nb, ref, stat, lo, hi, flags=0, **kw):
    need_qual = bool(ctx.lib.pss_pileup_needs_qual(_cabi.PILEUP_STATS.index(stat), int(kw.get('min_baseq', 0))))
    d = dict(nb)
    if not need_qual:
        d['qual'] = d['qual_off'] = d['name_prev'] = None
    rs = _cabi.make_reads_struct(d)
    withref = stat in config.stats_types_withref
    p = ctx.pileup_params(stat, start=lo, end=hi, ref=ref if withref else None, ref_start=0,
                          ref_len=len(ref) if withref else 0, contig_len=len(ref), flags=flags, **kw)
    nf = len(config.value_fields(stat))
    out = np.empty((nf, hi - lo), dtype=np.int32)
    mask = np.empty(hi - lo, dtype=np.uint8)
    ctx.pileup_host(rs, p, out, mask)
    return out, mask


def _oracle(stat, reads, ref, lo, hi, **kw):
    withref = stat in config.stats_types_withref
    pos, F, rc = orc.pileup_packed(stat, reads, lo, hi, ref=ref.tobytes() if withref else None, truncate=True, **kw)
    return pos, F


def _compare(out, mask, pos, F, lo, label):
    rows = np.flatnonzero(mask)
    assert np.array_equal(rows + lo, pos), label
    got = out[:, rows].T
    if not np.array_equal(got, F):
        bad = np.argwhere(got != F)[0]
        raise AssertionError('%s: row %d field %d: got %d expect %d' % (label, bad[0], bad[1], got[tuple(bad)], F[tuple(bad)]))


@pytest.mark.parametrize('stat', config.pileup_types)
def test_synthetic_vs_oracle(stat, ctx, small):
    nb, ref, reads = small
    kws = (dict(), dict(no_del=True), dict(min_mapq=20, no_dup=True), dict(stepper='nofilter'), dict(min_baseq=20))
    if stat == 'coverage_gc':
        # the synthetic reference has N runs and the windows are clipped at both contig ends
        kws += (dict(window_size=300, window_offset=150), dict(window_size=101, window_offset=0),
                dict(window_size=1000, window_offset=999), dict(window_size=64, window_offset=200))
    for kw in kws:
        out, mask = _run_host(ctx, nb, ref, stat, 0, G, **kw)
        pos, F = _oracle(stat, reads, ref, 0, G, **kw)
        _compare(out, mask, pos, F, 0, '%s %r' % (stat, kw))


@pytest.mark.parametrize('stat', TILE_STATS)
def test_tile_kernel_equals_generic_kernel(stat, ctx, small):
    nb, ref, _ = small
    for kw in (dict(), dict(no_del=True, min_mapq=30), dict(min_baseq=25), dict(min_baseq=36, no_dup=True),
               dict(min_baseq=30, no_del=True)):
        a, ma = _run_host(ctx, nb, ref, stat, 1000, G - 777, **kw)
        b, mb = _run_host(ctx, nb, ref, stat, 1000, G - 777, flags=_cabi.PSS_FLAG_GENERIC_KERNEL, **kw)
        assert np.array_equal(a, b) and np.array_equal(ma, mb), (stat, kw)


def test_reference_beyond_fasta_and_odd_ranges(ctx, small):
    # positions past the end of the reference behave like an empty fetch (every base a mismatch,
    # opt.pyx:1987-1994): the tile kernel's per-base slow path; odd start/length exercise the
    # unaligned (scalar-store) finalisation
    nb, ref, reads = small
    short = ref[:G // 2 + 13]
    for stat in ('variation', 'variation_strand'):
        out, mask = _run_host(ctx, nb, short, stat, 3, G - 5)
        pos, F = _oracle(stat, reads, short, 3, G - 5)
        _compare(out, mask, pos, F, 3, stat)


def test_sharded_equals_whole(ctx, small):
    nb, ref, _ = small
    endpos = nb['endpos']
    whole, wmask = _run_host(ctx, nb, ref, 'variation_strand', 0, G)
    parts = []
    for lo, hi in shard.plan_intervals(0, G, 3):
        i0, i1, _ = shard.halo_range(nb['pos'], endpos, lo, hi)
        view = {k: (nb[k][i0:] if k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off') else nb[k])
                for k in _cabi.READ_FIELDS}
        view['qual'] = view['qual_off'] = view['name_prev'] = None
        view['n_reads'] = i1 - i0
        view['tid'] = 0
        rs = _cabi.make_reads_struct(view, n_reads=i1 - i0)
        p = ctx.pileup_params('variation_strand', start=lo, end=hi, ref=ref[lo:], ref_start=lo, ref_len=hi - lo,
                              contig_len=len(ref))
        out = np.empty((60, hi - lo), dtype=np.int32)
        mask = np.empty(hi - lo, dtype=np.uint8)
        ctx.pileup_host(rs, p, out, mask)
        parts.append((mask, out))
    mask, out = shard.concat_shards(parts)
    assert np.array_equal(mask, wmask) and np.array_equal(out, whole)


def test_device_entry_points_and_prepared_batch(ctx):
    import torch
    dev = torch.device('cuda', ctx.device)
    L = 2_000_000
    batch, (ref, ref_lo) = synth.make_interval_batch(0, L, L, depth=30, device=dev, with_qual=False)
    rs = _cabi.make_reads_struct(batch)
    s = torch.cuda.current_stream().cuda_stream
    outs = {}
    ctx.prepare(rs, stepper='all', max_depth=8000, stream=s)
    for stat in ('variation_strand', 'coverage_ext', 'variation', 'coverage'):
        nf = len(config.value_fields(stat))
        o = torch.empty((nf, L), dtype=torch.int32, device=dev)
        m = torch.empty(L, dtype=torch.uint8, device=dev)
        withref = stat.startswith('variation')
        p = ctx.pileup_params(stat, start=0, end=L, ref=ref if withref else None, ref_start=ref_lo,
                              ref_len=int(ref.numel()) if withref else 0, contig_len=L)
        ctx.pileup_prepared(p, o, m, stream=s)
        outs[stat] = (o, m)
    ctx.synchronize(s)
    vs, cov = outs['variation_strand'][0].long(), outs['coverage'][0].long()
    # linear identities of the counters, at a size the oracle is not run on
    assert torch.equal(vs[0], cov[0]) and torch.equal(vs[3], cov[1])                 # reads_all, reads_pp
    assert torch.equal(vs[1] + vs[2], vs[0])                                          # fwd + rev = all
    assert torch.equal(outs['coverage_ext'][0][0].long(), cov[0])
    nondel = vs[6] + vs[12]                                                           # matches + mismatches
    assert torch.equal(nondel + vs[18], vs[0])                                        # + deletions (no N ops here)
    bases = vs[30] + vs[36] + vs[42] + vs[48] + vs[54]                                # A C T G N cover every base
    assert torch.equal(bases, nondel)
    v2 = outs['variation'][0].long()
    assert torch.equal(v2[0], vs[0]) and torch.equal(v2[2], vs[6]) and torch.equal(v2[10], vs[30])
    # one-shot entry point == prepared path
    o2 = torch.empty_like(outs['variation_strand'][0])
    m2 = torch.empty_like(outs['variation_strand'][1])
    p = ctx.pileup_params('variation_strand', start=0, end=L, ref=ref, ref_start=ref_lo, ref_len=int(ref.numel()),
                          contig_len=L)
    ctx.pileup(rs, p, o2, m2, stream=s)
    ctx.synchronize(s)
    assert torch.equal(o2, outs['variation_strand'][0]) and torch.equal(m2, outs['variation_strand'][1])
    # a prepared batch refuses mismatching read-level arguments
    p_bad = ctx.pileup_params('coverage', start=0, end=L, min_mapq=5)
    with pytest.raises(ValueError):
        ctx.pileup_prepared(p_bad, outs['coverage'][0], outs['coverage'][1], stream=s)


def test_very_deep_single_position(ctx):
    # 70 000 reads at one start: tiles with many chunks, counters far beyond 16 bits
    n = 70000
    L = 100
    nb = dict(n_reads=n, tid=0,
              pos=np.full(n, 1000, 'i4'), flag=np.where(np.arange(n) % 2, 16, 0).astype('u2'),
              mapq=np.full(n, 60, 'u1'), tlen=np.zeros(n, 'i4'), mtid=np.full(n, -1, 'i4'),
              l_seq=np.full(n, L, 'i4'), cigar_off=np.arange(n + 1, dtype='i8'),
              cigar=np.full(n, (L << 4), 'u4'), seq_off=np.arange(n + 1, dtype='i8') * (L // 2),
              seq4=np.tile(np.array([0x12, 0x48], 'u1'), n * L // 4), qual=None, qual_off=None, name_prev=None)
    ref = np.frombuffer(b'ACGT' * 600, dtype='u1')
    out, mask = _run_host(ctx, nb, ref, 'variation_strand', 900, 1300, max_depth=1000000)
    names = config.value_fields('variation_strand')
    col = {f: out[j] for j, f in enumerate(names)}
    cov = np.zeros(400, 'i8'); cov[100:200] = n
    assert np.array_equal(col['reads_all'], cov) and np.array_equal(mask, (cov > 0).astype('u1'))
    assert np.array_equal(col['reads_fwd'], cov // 2) and np.array_equal(col['reads_rev'], cov // 2)
    assert np.array_equal(col['matches'], cov) and col['mismatches'].sum() == 0       # read = ACGT... = reference phase
    for b, base in enumerate('ACGT'):
        e = np.zeros(400, 'i8'); e[100 + b:200:4] = n
        assert np.array_equal(col[base], e), base
    # variation without the strand split (one packed word per group: all | proper pair << 16): the same tile in two
    # passes, every third read a proper pair
    nbp = dict(nb, flag=(nb['flag'] | np.where(np.arange(n) % 3 == 0, 3, 0)).astype('u2'))
    out, mask = _run_host(ctx, nbp, ref, 'variation', 900, 1300, max_depth=1000000)
    col = {f: out[j] for j, f in enumerate(config.value_fields('variation'))}
    npp = (n + 2) // 3
    assert np.array_equal(col['reads_all'], cov) and np.array_equal(col['reads_pp'], cov // n * npp)
    assert np.array_equal(col['matches'], cov) and np.array_equal(col['matches_pp'], cov // n * npp) and col['mismatches'].sum() == 0
    for b, base in enumerate('ACGT'):
        e = np.zeros(400, 'i8'); e[100 + b:200:4] = n
        assert np.array_equal(col[base], e) and np.array_equal(col[base + '_pp'], e // n * npp), base
    # the depth cap at the default 8000: the first read is always admitted, then up to the cap
    out, mask = _run_host(ctx, nb, ref, 'coverage', 900, 1300)
    assert out[0][150] == 8000
    # the quality kinds keep 64-bit squared sums: a tile this deep is still one exact pass
    nq = dict(nb, qual=np.full(n * L, 30, 'u1'), qual_off=np.arange(n + 1, dtype='i8') * L,
              name_prev=np.full(n, -1, 'i4'))
    for stat, rms in (('baseq_strand', 'rms_baseq'), ('mapq_strand', 'rms_mapq'), ('baseq_ext', 'rms_baseq')):
        out, mask = _run_host(ctx, nq, ref, stat, 900, 1300, max_depth=1000000)
        col = {f: out[j] for j, f in enumerate(config.value_fields(stat))}
        assert np.array_equal(col['reads_all'], cov), stat
        assert np.array_equal(col[rms], np.where(cov > 0, 30 if 'baseq' in stat else 60, 0)), stat
        if stat == 'mapq_strand':
            assert np.array_equal(col['max_mapq_fwd'], np.where(cov > 0, 60, 0))


@pytest.mark.parametrize('stat', ('mapq', 'mapq_strand'))
def test_mapq_tile_deeper_than_16_bits(stat, ctx):
    # k_tilem packs n | nz << 16 per bank while a tile has at most 65535 candidate reads; a deeper tile is done in two
    # sweeps, one per half of its positions, with full-word counters.  72 000 reads in one 512-position tile (70 000 of
    # them over the same positions), starts in both halves, every mapping quality incl. 0, deletions and a ref-skip in the CIGARs, against the oracle --
    # plain, min_baseq (events of low bases dropped: the kept runs are the intervals) and no_del (aligned spans only)
    rng = np.random.default_rng(7)
    n, L = 72000, 60
    pos = np.sort(np.concatenate([rng.integers(1000, 1010, 35000), rng.integers(1020, 1030, 35000),
                                  rng.integers(1290, 1330, 2000)])).astype('i4')
    kind = rng.integers(0, 4, n)                      # 60M | 20M 5D 40M | 30M 40N 30M | 10S 50M
    cig, off = [], [0]
    for k in kind:
        ops = ([(0, 60)], [(0, 20), (2, 5), (0, 40)], [(0, 30), (3, 40), (0, 30)], [(4, 10), (0, 50)])[k]
        cig += [(l << 4) | op for op, l in ops]
        off.append(len(cig))
    flag = (np.where(rng.random(n) < 0.5, 16, 0) | np.where(rng.random(n) < 0.6, 3, 1)).astype('u2')
    nb = dict(n_reads=n, tid=0, pos=pos, flag=flag, mapq=rng.choice([0, 0, 1, 13, 29, 60, 254, 255], n).astype('u1'),
              tlen=np.zeros(n, 'i4'), mtid=np.zeros(n, 'i4'), l_seq=np.full(n, L, 'i4'),
              cigar_off=np.array(off, 'i8'), cigar=np.array(cig, 'u4'), seq_off=np.arange(n + 1, dtype='i8') * (L // 2),
              seq4=rng.choice(np.array([0x11, 0x12, 0x48, 0x84], 'u1'), n * L // 2),
              qual=rng.integers(0, 41, n * L).astype('u1'), qual_off=np.arange(n + 1, dtype='i8') * L,
              name_prev=np.full(n, -1, 'i4'))
    ref = np.frombuffer(b'ACGT' * 600, dtype='u1')
    reads = oracle_reads_from_arrays(nb)
    for kw in (dict(), dict(min_baseq=20), dict(no_del=True), dict(min_baseq=30, no_del=True, min_mapq=1)):
        out, mask = _run_host(ctx, nb, ref, stat, 900, 1500, max_depth=1000000, **kw)
        p, F = _oracle(stat, reads, ref, 900, 1500, max_depth=1000000, **kw)
        _compare(out, mask, p, F, 900, '%s %r' % (stat, kw))


@pytest.mark.parametrize('stat', config.binned_types)
def test_binned_synthetic_vs_oracle_and_sharded(stat, ctx, small):
    nb, ref, reads = small
    W = 300
    p = ctx.binned_params(stat, 0, G, W, ref=ref, ref_start=0, ref_len=len(ref))
    nbins = int(p.n_bins)
    nf = len(config.value_fields_binned(stat))
    d = dict(nb)
    d['qual'] = d['qual_off'] = d['name_prev'] = None
    whole = np.empty((nf, nbins), dtype=np.int32)
    ctx.binned_host(_cabi.make_reads_struct(d), p, whole)
    pos, F = orc.binned_packed(stat, reads, 0, G, W, W // 2, ref=ref.tobytes())
    assert np.array_equal(whole.T, F), stat
    # window-range shards: bin edges use the global start, every read belongs to exactly one shard
    parts = []
    for k_lo, k_hi in shard.plan_windows(0, G, W, 3):
        i0, i1 = shard.window_read_range(nb['pos'], 0, W, k_lo, k_hi, nbins)
        view = {k: (nb[k][i0:] if k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off') else nb[k])
                for k in _cabi.READ_FIELDS}
        view['qual'] = view['qual_off'] = view['name_prev'] = None
        view['n_reads'], view['tid'] = i1 - i0, 0
        s0, s1 = k_lo * W, min(G, k_hi * W)
        ps = ctx.binned_params(stat, s0, s1, W, ref=ref, ref_start=0, ref_len=len(ref))
        assert int(ps.n_bins) == k_hi - k_lo
        o = np.empty((nf, k_hi - k_lo), dtype=np.int32)
        ctx.binned_host(_cabi.make_reads_struct(view, n_reads=i1 - i0), ps, o)
        parts.append((o,))
    assert np.array_equal(shard.concat_shards(parts)[0], whole), stat


@pytest.mark.parametrize('stat', ('coverage_strand', 'coverage_ext', 'variation_strand'))
def test_narrow_transport_levels(stat, ctx, monkeypatch):
    # The host entry points bring the counts back as 8-bit columns, redo a position chunk with
    # 16-bit columns when a count does not fit, and then skip the 8-bit attempt for a few chunks:
    # depth 20 / 300 / 20 / 600 / 20 along the range, 1024 positions per chunk.
    monkeypatch.setenv('PSS_HOST_CHUNK', '1024')
    rng = np.random.default_rng(7)
    L, G = 100, 40000
    dens = np.full(G - L, 0.2)
    dens[9000:9600] = 3.0
    dens[30000:30300] = 6.0
    pos = np.repeat(np.arange(G - L), rng.poisson(dens)).astype('i4')
    n = len(pos)
    ref = rng.choice(list(b'ACGT'), G).astype('u1')
    code = np.zeros(256, 'u1')
    code[list(b'ACGT')] = (1, 2, 4, 8)
    bases = code[ref[pos[:, None] + np.arange(L)[None, :]]]
    sub = rng.random(bases.shape) < 0.01
    bases[sub] = rng.choice([1, 2, 4, 8, 15], int(sub.sum()))
    flag = (rng.choice([0, 16], n) | rng.choice([0, 3], n) | rng.choice([0, 1024], n, p=[.95, .05])).astype('u2')
    nb = dict(n_reads=n, tid=0, pos=pos, flag=flag, mapq=np.full(n, 60, 'u1'),
              tlen=rng.choice([-300, 300], n).astype('i4'), mtid=np.zeros(n, 'i4'), l_seq=np.full(n, L, 'i4'),
              cigar_off=np.arange(n + 1, dtype='i8'), cigar=np.full(n, L << 4, 'u4'),
              seq_off=np.arange(n + 1, dtype='i8') * (L // 2),
              seq4=((bases[:, 0::2] << 4) | bases[:, 1::2]).astype('u1').ravel(),
              qual=None, qual_off=None, name_prev=np.full(n, -1, 'i4'))
    reads = oracle_reads_from_arrays(nb)
    for kw in (dict(max_depth=100000), dict(max_depth=100000, stepper='nofilter', no_dup=True)):
        h0, d0 = ctx.transfer_bytes()
        out, mask = _run_host(ctx, nb, ref, stat, 0, G, **kw)
        h1, d1 = ctx.transfer_bytes()
        pos_o, F = _oracle(stat, reads, ref, 0, G, **kw)
        _compare(out, mask, pos_o, F, 0, '%s %r' % (stat, kw))
        assert out.max() > 255
        # most chunks travelled as bytes: well under the 16-bit volume
        assert d1 - d0 < 0.9 * (2 * out.size + mask.size), (d1 - d0, out.size)


def test_streamed_host_call_equals_single_call(ctx, small, monkeypatch):
    # pss_pileup_host_stream: the range is cut into chunks, chunk k + 1's reads are uploaded (a sub-range of
    # the batch found from max_span) while chunk k is computed; same columns as one pss_pileup_host call
    nb, ref, _ = small
    monkeypatch.setenv('PSS_HOST_CHUNK', '4096')
    d = dict(nb)
    d['qual'] = d['qual_off'] = d['name_prev'] = None
    rs = _cabi.make_reads_struct(d)
    lo, hi = 1000, G - 321
    stats = ('variation_strand', 'coverage_ext', 'coverage')
    max_span = int((nb['endpos'] - nb['pos']).max())
    # max_depth = 12 at 30x: the depth cap refuses reads everywhere, also across every chunk boundary -- the
    # library must not cut such a batch (the admission rule is order dependent), results = the single call's
    for kw in (dict(), dict(no_del=True, min_mapq=30), dict(max_depth=12)):
        ps, outs, masks = [], [], []
        for s in stats:
            withref = s in config.stats_types_withref
            ps.append(ctx.pileup_params(s, start=lo, end=hi, ref=ref if withref else None, ref_start=0,
                                        ref_len=len(ref) if withref else 0, contig_len=len(ref), **kw))
            outs.append(np.full((len(config.value_fields(s)), hi - lo), -7, dtype=np.int32))
            masks.append(np.full(hi - lo, 9, dtype=np.uint8))
        ctx.pileup_host_stream(rs, ps, outs, masks, max_span, 20480)
        for s, o, m in zip(stats, outs, masks):
            e, em = _run_host(ctx, nb, ref, s, lo, hi, **kw)
            assert np.array_equal(o, e) and np.array_equal(m, em), (s, kw)


@pytest.mark.parametrize('stat', ('tlen', 'tlen_strand'))
def test_tlen_rounding_boundaries_and_replay(stat, ctx, small, monkeypatch):
    # std_tlen is decided from exact integer sums unless it is too close to a rounding boundary; then the tile
    # is redone by the in-order Welford replay.  Nine reads with tlen 1,1,1,0,0,0,0,0,0 have std exactly 0.5
    # (a tie whose rounding depends on the order the reference saw the reads in); other multisets sit on
    # (k + 0.5) boundaries too.  Every permutation must agree with the oracle's in-order replay.
    rng = np.random.default_rng(21)
    L = 50
    groups = []
    for base in ([1, 1, 1, 0, 0, 0, 0, 0, 0], [3, 3, 3, 0, 0, 0, 0, 0, 0], [401, 401, 401, 400, 400, 400, 400, 400, 400],
                 [-7, -7, -7, -8, -8, -8, -8, -8, -8], [100000001, 100000001, 100000001] + [100000000] * 6):
        for _ in range(6):
            groups.append(rng.permutation(base))
    n = sum(len(g) for g in groups)
    pos = np.repeat(np.arange(len(groups)) * 700 + 100, [len(g) for g in groups]).astype('i4')
    nb = dict(n_reads=n, tid=0, pos=pos, flag=np.where(rng.random(n) < 0.5, 99, 147).astype('u2'),
              mapq=np.full(n, 60, 'u1'), tlen=np.concatenate(groups).astype('i4'), mtid=np.zeros(n, 'i4'),
              l_seq=np.full(n, L, 'i4'), cigar_off=np.arange(n + 1, dtype='i8'), cigar=np.full(n, L << 4, 'u4'),
              seq_off=np.arange(n + 1, dtype='i8') * (L // 2), seq4=np.full(n * L // 2, 0x11, 'u1'),
              qual=None, qual_off=None, name_prev=np.full(n, -1, 'i4'), endpos=pos.astype('i8') + L)
    reads = oracle_reads_from_arrays(nb)
    ref = np.frombuffer(b'A' * (int(pos.max()) + 200), dtype='u1')
    hi = int(pos.max()) + 100
    out, mask = _run_host(ctx, nb, ref, stat, 0, hi)
    pos_o, F = _oracle(stat, reads, ref, 0, hi)
    _compare(out, mask, pos_o, F, 0, stat)
    j = config.value_fields(stat).index('std_tlen')
    assert set(np.unique(out[j][mask.astype(bool)])) >= {0, 2}            # 0.5 -> 0 (half-even), 1.5 -> 2, ...
    # every tile through the replay path: same answer on the synthetic batch
    nbs, refs, rs = small
    a, ma = _run_host(ctx, nbs, refs, stat, 0, G)
    monkeypatch.setenv('PSS_TLEN_FORCE_REPLAY', '1')
    b, mb = _run_host(ctx, nbs, refs, stat, 0, G)
    assert np.array_equal(a, b) and np.array_equal(ma, mb)


@pytest.mark.parametrize('stat', ('coverage', 'variation_strand', 'baseq_ext', 'mapq', 'tlen'))
def test_long_reads_need_no_flag(stat, ctx):
    # reads of several kb do not fit the staging buffers of the short-read kernels; the host entry point repeats
    # the call on the long-read kernel by itself (the reference API has no way to pass PSS_FLAG_LONG_READS)
    rng = np.random.default_rng(5)
    G = 60000
    ref = rng.choice(list(b'ACGT'), G).astype('u1')
    code = np.zeros(256, 'u1')
    code[list(b'ACGT')] = (1, 2, 4, 8)
    lens = np.array([12000, 150, 7000, 150, 150, 9001, 150])
    pos = np.sort(rng.integers(0, G - 13000, len(lens))).astype('i4')
    n = len(lens)
    seqs, quals, seq_off, qual_off = [], [], [0], [0]
    for p, ln in zip(pos, lens):
        b = code[ref[p:p + ln]].copy()
        sub = rng.random(ln) < 0.02
        b[sub] = rng.choice([1, 2, 4, 8], int(sub.sum()))
        if ln % 2:
            b = np.append(b, 0).astype('u1')
        seqs.append(((b[0::2] << 4) | b[1::2]).astype('u1'))
        quals.append(rng.integers(2, 41, ln).astype('u1'))
        seq_off.append(seq_off[-1] + len(seqs[-1]))
        qual_off.append(qual_off[-1] + ln)
    nb = dict(n_reads=n, tid=0, pos=pos, flag=rng.choice([0, 16], n).astype('u2'), mapq=rng.choice([0, 30, 60], n).astype('u1'),
              tlen=np.zeros(n, 'i4'), mtid=np.full(n, -1, 'i4'), l_seq=lens.astype('i4'),
              cigar_off=np.arange(n + 1, dtype='i8'), cigar=(lens << 4).astype('u4'),
              seq_off=np.array(seq_off, 'i8'), seq4=np.concatenate(seqs), qual_off=np.array(qual_off, 'i8'),
              qual=np.concatenate(quals), name_prev=np.full(n, -1, 'i4'), endpos=pos.astype('i8') + lens)
    reads = oracle_reads_from_arrays(nb)
    out, mask = _run_host(ctx, nb, ref, stat, 0, G)
    pos_o, F = _oracle(stat, reads, ref, 0, G)
    _compare(out, mask, pos_o, F, 0, stat)


# ---- pss_pileup_host_rows: packed records == the int32 columns of pss_pileup_host -------------------

def _run_rows(ctx, nb, ref, stat, lo, hi, pad=None, capacity=None, flags=0, **kw):
    need_qual = bool(ctx.lib.pss_pileup_needs_qual(_cabi.PILEUP_STATS.index(stat), int(kw.get('min_baseq', 0))))
    d = dict(nb)
    if not need_qual:
        d['qual'] = d['qual_off'] = d['name_prev'] = None
    rs = _cabi.make_reads_struct(d)
    withref = stat in config.stats_types_withref
    p = ctx.pileup_params(stat, start=lo, end=hi, ref=ref if withref else None, ref_start=0,
                          ref_len=len(ref) if withref else 0, contig_len=len(ref), flags=flags, **kw)
    names = config.value_fields(stat)
    dt = np.dtype([('chrom', 'S5'), ('pos', '<i4')] + ([('ref', 'S1')] if withref else []) + [(f, '<i4') for f in names])
    layout = ctx.row_layout(stat, dt, names, chrom='chrS', pos_add=1, pad=pad)
    cap = hi - lo if capacity is None else capacity
    rows = np.full(cap + 2, 0x55, dtype='u1').repeat(dt.itemsize).view(dt)       # guard records around the result
    n, neg = ctx.pileup_host_rows(rs, p, layout, rows, 1, cap)
    guard = np.full(1, 0x55, dtype='u1').repeat(dt.itemsize).view(dt)
    # nothing outside rows[row0 : row0 + capacity] is touched (records beyond the n returned may have been written by an
    # attempt that had to be repeated with wider columns)
    assert rows[0].tobytes() == guard.tobytes() and rows[cap + 1].tobytes() == guard.tobytes()
    return rows[1:n + 1]


def _rows_from_columns(out, mask, lo, ref, stat, pad=None):
    emit = mask.astype(bool)
    if pad:
        emit[max(pad[0] - lo, 0):max(pad[1] - lo, 0)] = True
    sel = np.flatnonzero(emit)
    names = config.value_fields(stat)
    withref = stat in config.stats_types_withref
    dt = np.dtype([('chrom', 'S5'), ('pos', '<i4')] + ([('ref', 'S1')] if withref else []) + [(f, '<i4') for f in names])
    a = np.zeros(len(sel), dtype=dt)
    a['chrom'] = b'chrS'
    a['pos'] = sel + lo + 1
    if withref:
        p = sel + lo
        ok = p < len(ref)
        r = np.zeros(len(sel), 'u1')
        r[ok] = ref[p[ok]]
        lower = (r >= 97) & (r <= 122)
        a['ref'] = np.where(lower, r - 32, r).astype('u1').view('S1')
    for j, f in enumerate(names):
        a[f] = out[j][sel]
    return a


@pytest.mark.parametrize('stat', ['coverage', 'coverage_ext', 'variation', 'variation_strand', 'baseq_ext_strand', 'tlen',
                                  'mapq_strand', 'coverage_gc'])
def test_rows_entry_equals_columns(stat, ctx, small, monkeypatch):
    nb, ref, _ = small
    for chunk in (None, '2048'):
        if chunk:
            monkeypatch.setenv('PSS_HOST_CHUNK', chunk)
        else:
            monkeypatch.delenv('PSS_HOST_CHUNK', raising=False)
        for lo, hi, pad, kw in ((0, G, None, {}), (1000, 90001, (500, 3000), dict(no_dup=True, min_mapq=30)),
                                (G - 5000, G + 700, (G, G + 650), dict(min_baseq=20))):
            out, mask = _run_host(ctx, nb, ref, stat, lo, hi, **kw)
            exp = _rows_from_columns(out, mask, lo, ref, stat, pad)
            got = _run_rows(ctx, nb, ref, stat, lo, hi, pad=pad, **kw)
            assert got.shape == exp.shape and got.tobytes() == exp.tobytes(), (stat, lo, hi, pad, kw, chunk)
    with pytest.raises(ValueError):
        _run_rows(ctx, nb, ref, stat, 0, G, capacity=1000)


def test_rows_entry_beyond_16_bits(ctx):
    # 70 000 reads on one position: the 8-bit chunk is redone with 16 bits, those need several passes, so the records
    # come through the int32 columns of the same pipeline
    n, L = 70000, 100
    nb = dict(n_reads=n, tid=0,
              pos=np.full(n, 1000, 'i4'), flag=np.where(np.arange(n) % 2, 16, 0).astype('u2'),
              mapq=np.full(n, 60, 'u1'), tlen=np.zeros(n, 'i4'), mtid=np.full(n, -1, 'i4'),
              l_seq=np.full(n, L, 'i4'), cigar_off=np.arange(n + 1, dtype='i8'),
              cigar=np.full(n, (L << 4), 'u4'), seq_off=np.arange(n + 1, dtype='i8') * (L // 2),
              seq4=np.tile(np.array([0x12, 0x48], 'u1'), n * L // 4), qual=None, qual_off=None, name_prev=None)
    ref = np.frombuffer(b'ACGT' * 600, dtype='u1')
    for stat in ('variation_strand', 'coverage'):
        out, mask = _run_host(ctx, nb, ref, stat, 900, 1300, max_depth=1000000)
        exp = _rows_from_columns(out, mask, 900, ref, stat)
        got = _run_rows(ctx, nb, ref, stat, 900, 1300, max_depth=1000000)
        assert len(got) == 100 and got.tobytes() == exp.tobytes(), stat
        assert got['reads_all'].max() == n


def test_tlen_binned_small_and_huge_template_lengths(ctx, small):
    # k_binned_acc<tlen> sums |tlen| < 8192 with one 32-bit REDUX per sum and everything else in 64-bit pieces:
    # warps that mix both, the thresholds themselves, and the int32 extremes (sum of squares near 2^62 per read)
    nb, ref, _ = small
    nb = dict(nb)
    t = nb['tlen'].copy()
    rng = np.random.default_rng(7)
    idx = rng.choice(len(t), len(t) // 150, replace=False)
    t[idx] = rng.choice(np.array([8191, -8191, 8192, -8192, 2 ** 31 - 1, -2 ** 31 + 1, 100000, -70000, 0], 'i8'), len(idx)).astype('i4')
    nb['tlen'] = t
    reads = oracle_reads_from_arrays(nb)
    for W in (300, 7):
        p = ctx.binned_params('tlen', 0, G, W)
        d = dict(nb)
        d['qual'] = d['qual_off'] = d['name_prev'] = None
        out = np.empty((len(config.value_fields_binned('tlen')), int(p.n_bins)), dtype=np.int32)
        ctx.binned_host(_cabi.make_reads_struct(d), p, out)
        pos, F = orc.binned_packed('tlen', reads, 0, G, W, W // 2)
        assert np.array_equal(out.T, F), W


def test_binned_multi_equals_single_calls(ctx, small):
    # pss_binned_host_multi / pss_binned_multi: any set of the five statistics from ONE pass over the read records
    # (one accumulator layout per set) == the single-statistic calls, which are oracle-checked above
    import itertools
    import torch
    nb, ref, reads = small
    d = dict(nb)
    d['qual'] = d['qual_off'] = d['name_prev'] = None
    rs = _cabi.make_reads_struct(d)
    for (start, end, W, kw) in ((0, G, 300, dict()), (1234, G - 77, 77, dict(min_mapq=20, no_dup=True)), (5, 300000, 1, dict()),
                                (100000, 100100, 1000, dict())):
        single = {}
        for stat in config.binned_types:
            p = ctx.binned_params(stat, start, end, W, ref=ref, ref_start=0, ref_len=len(ref), **kw)
            single[stat] = np.empty((len(config.value_fields_binned(stat)), int(p.n_bins)), dtype=np.int32)
            ctx.binned_host(rs, p, single[stat])
            if (start, end) == (0, G):
                pos, F = orc.binned_packed(stat, reads, start, end, W, W // 2, ref=ref.tobytes(), **kw)
                assert np.array_equal(single[stat].T, F), stat
        subsets = [c for r in range(1, 6) for c in itertools.combinations(config.binned_types, r)]
        for sub in (subsets if W == 300 else subsets[-6:]):
            ps = [ctx.binned_params(t, start, end, W, ref=ref, ref_start=0, ref_len=len(ref), **kw) for t in sub]
            outs = [np.full_like(single[t], -9) for t in sub]
            ctx.binned_host_multi(rs, ps, outs)
            for t, o in zip(sub, outs):
                assert np.array_equal(o, single[t]), (sub, t, W)
    # device entry point, the bench's sweep
    dev = torch.device('cuda', ctx.device)
    db = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    db['flag'] = torch.from_numpy(nb['flag'].view(np.int16)).to(dev)
    db['cigar'] = torch.from_numpy(nb['cigar'].view(np.int32)).to(dev)
    rd = _cabi.make_reads_struct(db, n_reads=int(nb['n_reads']))
    dref = torch.from_numpy(ref).to(dev)
    sub = ('coverage', 'mapq', 'tlen')
    ps = [ctx.binned_params(t, 0, G, 300, ref=dref, ref_start=0, ref_len=len(ref)) for t in sub]
    outs = [torch.full((len(config.value_fields_binned(t)), int(ps[0].n_bins)), -9, dtype=torch.int32, device=dev) for t in sub]
    ctx.binned_multi(rd, ps, outs, stream=None)
    ctx.synchronize(None)
    for t, o in zip(sub, outs):
        p1 = ctx.binned_params(t, 0, G, 300, ref=ref, ref_start=0, ref_len=len(ref))
        e = np.empty((len(config.value_fields_binned(t)), int(p1.n_bins)), dtype=np.int32)
        ctx.binned_host(rs, p1, e)
        assert np.array_equal(o.cpu().numpy(), e), t
    # argument errors
    with pytest.raises(ValueError):
        ctx.binned_host_multi(rs, [ps[0], ps[0]], [np.empty((3, 1000), np.int32)] * 2)       # host/device mix aside: twice the same statistic
