"""GPU: the parity tests proper.  Everything goes through the public API (and therefore the
C ABI, ``pss_pileup_host`` / ``pss_binned_host``) and is compared bit-exactly with

* the reference's golden file ``regression.npz`` (test_regression.py:34-58),
* the reference's known-answer tests (test_pileup.py:939-1102, test_binned.py:583-626),
* the CPU oracle for everything the golden vectors do not pin (filters, steppers, rna,
  deep, one_based, upper-case FASTA -- the cases of test_pileup.py:25-61 / test_binned.py:21-53).
"""
import numpy as np
import pytest

import pysamstats_b200 as pysamstats
from pysamstats_b200 import config
from oracle import bamread, oracle as orc
from _util import fx, assert_recarrays_equal

pytestmark = pytest.mark.gpu

WITHREF = ('variation', 'variation_strand', 'baseq_ext', 'baseq_ext_strand', 'coverage_gc')


@pytest.fixture(scope='module')
def golden():
    return np.load(fx('regression.npz'))


@pytest.fixture(scope='module')
def obam():
    return {n: bamread.Bam(fx(n + '.bam')) for n in ('test', 'deep', 'rna', 'longcontignames')}


@pytest.fixture(scope='module')
def ofa():
    return bamread.Fasta(fx('ref.fa'))


def _load(stat, bam, **kw):
    fa = fx(kw.pop('fasta', 'ref.fa')) if stat in WITHREF else None
    return getattr(pysamstats, 'load_' + stat)(fx(bam + '.bam'), fafile=fa, **kw)


def _oracle(stat, obam, ofa, bam, **kw):
    kw.pop('fasta', None)
    return orc.load_pileup(stat, obam[bam], ofa if stat in WITHREF else None, **kw)


# ---- golden file -------------------------------------------------------------------

@pytest.mark.parametrize('stat', config.pileup_types)
def test_regression_pileup(stat, golden):
    assert_recarrays_equal(golden[stat], _load(stat, 'test'), stat)


@pytest.mark.parametrize('stat', config.binned_types)
def test_regression_binned(stat, golden):
    a = getattr(pysamstats, 'load_%s_binned' % stat)(fx('test.bam'), fafile=fx('ref.fa'))
    assert_recarrays_equal(golden[stat + '_binned'], a, stat + '_binned')


# ---- differential vs the oracle (the reference's compare_stats cases) -----------------

FILTERS = dict(min_mapq=1, min_baseq=17, no_del=True, no_dup=True)


@pytest.mark.parametrize('stat', config.pileup_types)
def test_region_and_filters(stat, obam, ofa):
    region = dict(chrom='Pf3D7_01_v3', start=0, end=2000)
    for kw in (dict(region, one_based=False), dict(region, start=1, one_based=True), dict(region, **FILTERS),
               dict(FILTERS), dict(min_baseq=30), dict(no_del=True), dict(min_mapq=30)):
        assert_recarrays_equal(_oracle(stat, obam, ofa, 'test', **kw), _load(stat, 'test', **kw),
                               '%s %r' % (stat, kw))


@pytest.mark.parametrize('stat', ['variation', 'variation_strand', 'baseq_ext', 'baseq_ext_strand', 'coverage',
                                  'coverage_ext', 'baseq', 'mapq', 'tlen'])
def test_rna_spliced(stat, obam, ofa):
    for kw in (dict(), dict(no_del=True), dict(min_baseq=20), dict(pad=True, chrom='Pf3D7_01_v3', start=900,
                                                                  end=1100)):
        assert_recarrays_equal(_oracle(stat, obam, ofa, 'rna', **kw), _load(stat, 'rna', **kw),
                               '%s rna %r' % (stat, kw))


@pytest.mark.parametrize('stat', ['variation', 'baseq_ext_strand', 'coverage_gc'])
def test_uppercase_fasta(stat, obam):
    ofa_up = bamread.Fasta(fx('ref.upper.fa'))
    e = orc.load_pileup(stat, obam['test'], ofa_up, chrom='Pf3D7_01_v3', start=0, end=2000)
    a = _load(stat, 'test', fasta='ref.upper.fa', chrom='Pf3D7_01_v3', start=0, end=2000)
    assert_recarrays_equal(e, a, stat)


@pytest.mark.parametrize('kw', [dict(window_size=1), dict(window_size=7, window_offset=0), dict(window_size=301, window_offset=300),
                                dict(window_size=64, window_offset=200), dict(window_size=1500), dict(window_size=9001),
                                dict(window_size=100000), dict(window_size=50, pad=True, chrom='Pf3D7_02_v3'),
                                dict(window_size=300, chrom='Pf3D7_01_v3', start=1001, end=7003, pad=True),
                                dict(window_size=1024, window_offset=1023, chrom='Pf3D7_03_v3', start=69000, pad=True)])
def test_coverage_gc_windows(kw, obam, ofa):
    # k_gc: table path (unclipped, window <= 1024), division path (clipped by the contig ends or a wider window),
    # per-position path (the window does not fit the CTA's reference span); gc_content opt.pyx:1527-1539
    assert_recarrays_equal(_oracle('coverage_gc', obam, ofa, 'test', **kw), _load('coverage_gc', 'test', **kw),
                           'coverage_gc %r' % (kw,))


@pytest.mark.parametrize('stat', ['coverage', 'coverage_ext_strand', 'variation_strand', 'baseq_strand', 'tlen_strand',
                                  'mapq_strand'])
@pytest.mark.parametrize('max_depth', [8000, 100, 1, 1000000])
def test_deep_vs_oracle(stat, max_depth, obam, ofa):
    for kw in (dict(max_depth=max_depth), dict(max_depth=max_depth, min_baseq=17, no_dup=True)):
        assert_recarrays_equal(_oracle(stat, obam, ofa, 'deep', **kw), _load(stat, 'deep', **kw),
                               '%s deep %r' % (stat, kw))


@pytest.mark.parametrize('stepper', ['all', 'nofilter', 'samtools'])
def test_steppers_vs_oracle(stepper, obam, ofa):
    lc = obam['longcontignames']
    # contigs with the most records exercise DUP / SECONDARY / unmapped handling
    counts = np.bincount(lc.tid[lc.tid >= 0], minlength=len(lc.references))
    for tid in np.argsort(counts)[::-1][:25]:
        c = lc.references[tid]
        for stat in ('coverage_ext', 'baseq', 'mapq_strand'):
            kw = dict(chrom=c, stepper=stepper)
            assert_recarrays_equal(_oracle(stat, obam, ofa, 'longcontignames', **kw),
                                   _load(stat, 'longcontignames', **kw), '%s %s %s' % (stat, c, stepper))
    kw = dict(stepper=stepper, no_dup=True)
    assert_recarrays_equal(_oracle('coverage', obam, ofa, 'longcontignames', **kw),
                           _load('coverage', 'longcontignames', **kw), 'wg coverage %s' % stepper)


# ---- the reference's known-answer tests ------------------------------------------------

def test_pileup_truncate_kat():
    for stat in config.pileup_types:
        a = _load(stat, 'test', chrom='Pf3D7_01_v3', start=2000, end=2100, one_based=False, truncate=False)
        assert (a['pos'][0], a['pos'][-1]) == (1952, 2154), stat
        a = _load(stat, 'test', chrom='Pf3D7_01_v3', start=2000, end=2100, one_based=False, truncate=True)
        assert (a['pos'][0], a['pos'][-1]) == (2000, 2099), stat


def test_pileup_pad_kat():
    for stat in config.pileup_types:
        a = _load(stat, 'test', chrom='Pf3D7_01_v3', start=0, end=20000, pad=False)
        assert (a['pos'][0], a['pos'][-1]) == (924, 9935), stat
        a = _load(stat, 'test', chrom='Pf3D7_01_v3', start=0, end=20000, pad=True)
        assert (a['pos'][0], a['pos'][-1]) == (0, 19999) and np.all(np.diff(a['pos']) == 1), stat


def test_pileup_pad_wg_kat():
    for stat in config.pileup_types:
        a = _load(stat, 'test', pad=False)
        assert sorted(set(a['chrom'])) == [b'Pf3D7_01_v3', b'Pf3D7_02_v3']
        x = a[a['chrom'] == b'Pf3D7_01_v3']
        assert (x['pos'][0], x['pos'][-1]) == (924, 9935)
        x = a[a['chrom'] == b'Pf3D7_02_v3']
        assert (x['pos'][0], x['pos'][-1]) == (926, 10074)
        a = _load(stat, 'test', pad=True)
        assert sorted(set(a['chrom'])) == [b'Pf3D7_01_v3', b'Pf3D7_02_v3', b'Pf3D7_03_v3']
        for c, last in ((b'Pf3D7_01_v3', 50000), (b'Pf3D7_02_v3', 60000), (b'Pf3D7_03_v3', 70000)):
            x = a[a['chrom'] == c]
            assert (x['pos'][0], x['pos'][-1]) == (0, last)


def test_pileup_limit_kat():
    for stat in config.pileup_types:
        assert _load(stat, 'deep', fields=['reads_all'], max_depth=1000000)[70] == 26169, stat
        assert _load(stat, 'deep', fields=['reads_all'], max_depth=12000)[70] == 12046, stat
        assert _load(stat, 'deep', fields=['reads_all'])[70] == 8052, stat


def test_long_contig_name_kat():
    label = 'AS2_scf7180000696055'
    x = pysamstats.load_coverage(fx('longcontignames.bam'), chrom=label)
    assert len(label) == x.dtype['chrom'].itemsize
    x = pysamstats.load_coverage(fx('longcontignames.bam'), chrom=label, dtype={'chrom': 'S10'})
    assert 10 == x.dtype['chrom'].itemsize


def test_steppers_kat():
    seq, pos = 'AS2_scf7180000695891', 14311
    for exp_all, exp_pp, step in zip([7, 8, 4], [4, 5, 4], ['all', 'nofilter', 'samtools']):
        a = pysamstats.load_coverage(fx('longcontignames.bam'), chrom=seq, stepper=step, pad=True)
        assert (a[pos]['reads_all'], a[pos]['reads_pp']) == (exp_all, exp_pp)
    with pytest.raises(ValueError):
        pysamstats.load_coverage(fx('longcontignames.bam'), chrom=seq, stepper='notastepper')


def test_binned_kats_and_filters(obam, ofa):
    for stat in config.binned_types:
        f = getattr(pysamstats, 'load_%s_binned' % stat)
        a = f(fx('test.bam'), fafile=fx('ref.fa'), chrom='Pf3D7_01_v3', start=1000, end=20000, window_size=200,
              window_offset=100)
        assert set(a['chrom']) == {b'Pf3D7_01_v3'} and (a['pos'][0], a['pos'][-1]) == (1100, 19900)
        a = f(fx('test.bam'), fafile=fx('ref.fa'), window_size=200, window_offset=100)
        for c, last in ((b'Pf3D7_01_v3', 50100), (b'Pf3D7_02_v3', 60100), (b'Pf3D7_03_v3', 70100)):
            x = a[a['chrom'] == c]
            assert (x['pos'][0], x['pos'][-1]) == (100, last)
        for kw in (dict(chrom='Pf3D7_01_v3', start=0, end=2000, window_size=50),
                   dict(chrom='Pf3D7_01_v3', start=1000, end=3000, window_size=77, window_offset=3, one_based=True),
                   dict(min_mapq=1, no_dup=True), dict(window_size=1000, min_mapq=30)):
            e = orc.load_binned(stat, obam['test'], ofa, **kw)
            assert_recarrays_equal(e, f(fx('test.bam'), fafile=fx('ref.fa'), **kw), '%s_binned %r' % (stat, kw))
        for bam in ('longcontignames', 'deep', 'rna'):
            c = obam[bam].references[int(np.bincount(obam[bam].tid[obam[bam].tid >= 0]).argmax())]
            kw = dict(chrom=c, window_size=100, no_dup=True)
            e = orc.load_binned(stat, obam[bam], ofa if bam != 'longcontignames' else None, **kw) \
                if not (bam == 'longcontignames' and stat in ('coverage', 'coverage_ext')) else None
            if e is not None:
                a = f(fx(bam + '.bam'), fafile=fx('ref.fa') if bam != 'longcontignames' else None, **kw)
                assert_recarrays_equal(e, a, '%s_binned %s' % (stat, bam))


def test_stat_generators_match_load():
    # stat_* yields dicts equal to the rows load_* returns (compare_iterators, test/util.py:29-51)
    a = pysamstats.load_variation(fx('test.bam'), fx('ref.fa'), chrom='Pf3D7_01_v3', start=1000, end=1200)
    recs = list(pysamstats.stat_variation(fx('test.bam'), fx('ref.fa'), chrom='Pf3D7_01_v3', start=1000, end=1200))
    assert len(recs) == len(a)
    for rec, row in zip(recs, a):
        assert set(rec) == set(a.dtype.names)
        for k in a.dtype.names:
            v = row[k]
            assert rec[k] == (v.decode() if isinstance(v, bytes) else v), k
    b = list(pysamstats.stat_tlen_binned(fx('test.bam'), chrom='Pf3D7_01_v3', start=0, end=3000))
    assert [r['pos'] for r in b] == list(range(150, 3150, 300))
    # a file object of another library: only .filename is used
    class Foreign(object):
        filename = fx('test.bam').encode()
    c = pysamstats.load_coverage(Foreign(), chrom='Pf3D7_01_v3', start=1000, end=1200)
    assert np.array_equal(c['reads_all'], a['reads_all'])


def test_writers_and_cli(tmp_path):
    # write_csv == csv.writer over the stat_* dicts (what the reference's io.write_csv does), and the
    # command line front end with the reference's flags (one-based by default)
    import csv
    import io
    import subprocess
    import sys
    from pysamstats_b200.io import write_csv
    from _util import ROOT
    kw = dict(chrom='Pf3D7_01_v3', start=1000, end=1300)
    for stat, withref in (('variation', True), ('coverage_ext_strand', False), ('tlen_binned', False),
                          ('coverage_binned', True)):
        fa = fx('ref.fa') if withref else None
        buf = io.StringIO()
        write_csv(stat, buf, fx('test.bam'), fafile=fa, dialect='excel-tab', **kw)
        exp = io.StringIO()
        w = csv.writer(exp, dialect='excel-tab')
        fields = [t[0] for t in getattr(config, 'dtype_' + stat)]
        w.writerow(fields)
        for rec in getattr(pysamstats, 'stat_' + stat)(fx('test.bam'), fafile=fa, **kw):
            w.writerow([rec[f] for f in fields])
        assert buf.getvalue() == exp.getvalue(), stat
    out = tmp_path / 'cov.tsv'
    cmd = [sys.executable, ROOT + '/scripts/pysamstats', '--type', 'variation', '-f', fx('ref.fa'), '-c', 'Pf3D7_01_v3',
           '-s', '1001', '-e', '1301', '--min-mapq', '1', '--no-dup', '--fields', 'chrom,pos,ref,reads_all,A',
           '--output', str(out), fx('test.bam')]
    subprocess.check_call(cmd)
    rows = [r.split('\t') for r in out.read_text().replace('\r', '').strip().split('\n')]
    a = pysamstats.load_variation(fx('test.bam'), fx('ref.fa'), chrom='Pf3D7_01_v3', start=1001, end=1301,
                                  one_based=True, min_mapq=1, no_dup=True)
    assert rows[0] == ['chrom', 'pos', 'ref', 'reads_all', 'A'] and len(rows) == len(a) + 1
    assert [int(r[1]) for r in rows[1:]] == a['pos'].tolist() and [int(r[4]) for r in rows[1:]] == a['A'].tolist()
    rc = subprocess.call([sys.executable, ROOT + '/scripts/pysamstats', '--type', 'variation', fx('test.bam')],
                         stderr=subprocess.DEVNULL)
    assert rc == 2                                            # missing --fasta option


def test_region_queries_through_the_bai_index(obam, ofa, monkeypatch):
    # same answers when the reads of a region come through the BAI index (only the listed BGZF blocks
    # are inflated) instead of a whole-file decode
    from pysamstats_b200 import bamio
    monkeypatch.setenv('PSS_BAI_MIN_BYTES', '0')
    monkeypatch.setattr(bamio, '_bam_cache', {})
    for stat, kw in (('variation_strand', dict(chrom='Pf3D7_02_v3', start=20000, end=24000)),
                     ('baseq_ext', dict(chrom='Pf3D7_01_v3', start=1000, end=3000, truncate=True)),
                     ('coverage_ext', dict(chrom='Pf3D7_03_v3', start=100, end=50000, pad=True)),
                     ('tlen', dict(chrom='Pf3D7_01_v3', start=5000, end=9000, one_based=True))):
        assert_recarrays_equal(_oracle(stat, obam, ofa, 'test', **kw), _load(stat, 'test', **kw), '%s %r' % (stat, kw))
    bam = bamio.open_bam(fx('test.bam'))
    assert bam._all is None                                       # nothing decoded the whole file
    a = pysamstats.load_mapq_binned(fx('test.bam'), chrom='Pf3D7_01_v3', start=2000, end=20000)
    assert_recarrays_equal(orc.load_binned('mapq', obam['test'], None, chrom='Pf3D7_01_v3', start=2000, end=20000), a, 'mapq_binned')


def test_chunked_tables_equal_single_call(monkeypatch):
    # a contig streams through pileup_tables in pieces of PSS_TABLE_CHUNK positions (bounded host memory for
    # stat_* / the writers); same rows as one call, with pad / truncate / one_based and on every kernel family
    cases = (('variation_strand', dict()), ('coverage_ext', dict(chrom='Pf3D7_02_v3', start=1000, end=40000, pad=True)),
             ('coverage', dict(chrom='Pf3D7_01_v3', start=500, end=30000, truncate=True, one_based=True)),
             ('mapq_strand', dict(chrom='Pf3D7_03_v3')), ('tlen', dict(chrom='Pf3D7_01_v3', pad=True, start=0, end=50000)),
             ('variation', dict(no_del=True, min_mapq=20)),
             # requests that read base qualities are cut too: the pairs of the mate-overlap rewrite come from one
             # pass over the contig's read metadata (pss_pairs_host), every piece gets them as its name chains
             ('baseq_strand', dict(chrom='Pf3D7_01_v3')), ('baseq_ext_strand', dict()),
             ('mapq', dict(chrom='Pf3D7_02_v3', min_baseq=17)), ('coverage_ext', dict(min_baseq=25, stepper='nofilter')),
             ('tlen', dict(chrom='Pf3D7_03_v3', min_baseq=20, no_del=True)))
    expect = [_load(stat, 'test', **kw) for stat, kw in cases]
    monkeypatch.setenv('PSS_TABLE_CHUNK', '4096')
    for (stat, kw), e in zip(cases, expect):
        assert_recarrays_equal(e, _load(stat, 'test', **kw), '%s %r chunked' % (stat, kw))
    rows = list(pysamstats.stat_coverage(fx('test.bam'), chrom='Pf3D7_01_v3', start=1000, end=20000))
    monkeypatch.delenv('PSS_TABLE_CHUNK')
    assert rows == list(pysamstats.stat_coverage(fx('test.bam'), chrom='Pf3D7_01_v3', start=1000, end=20000))


def _gpu_list():
    # every GPU of the box, or the one GPU twice (two shards at once, each on its own context)
    import torch
    n = torch.cuda.device_count()
    return ','.join(str(i) for i in range(n)) if n > 1 else '0,0'


def test_pieces_across_gpus(monkeypatch):
    # PSS_GPUS: the position pieces of a contig go to the GPUs of the box in turn (a thread per GPU, no collective),
    # the host concatenates them in order -- same rows as one GPU; through the row sink and through the column tables
    cases = (('variation_strand', dict()), ('coverage_ext', dict(chrom='Pf3D7_02_v3', start=1000, end=40000, pad=True)),
             ('mapq', dict(chrom='Pf3D7_01_v3')), ('tlen_strand', dict(chrom='Pf3D7_03_v3')),
             ('coverage', dict(chrom='Pf3D7_01_v3', start=3, end=30000, pad=True, truncate=True)),
             ('baseq_ext', dict(chrom='Pf3D7_01_v3')))                      # quality request: pieces with the contig's pairs
    expect = [_load(stat, 'test', **kw) for stat, kw in cases]
    monkeypatch.setenv('PSS_TABLE_CHUNK', '4096')
    monkeypatch.setenv('PSS_GPUS', _gpu_list())
    for no_rows in (False, True):
        if no_rows:
            monkeypatch.setenv('PSS_NO_ROWS', '1')
        for (stat, kw), e in zip(cases, expect):
            assert_recarrays_equal(e, _load(stat, 'test', **kw), '%s %r on %s (rows %r)' % (stat, kw, _gpu_list(), not no_rows))
    # stat_* (the generator of dicts) over the sharded tables
    rows = list(pysamstats.stat_coverage(fx('test.bam'), chrom='Pf3D7_01_v3', start=1000, end=20000))
    monkeypatch.delenv('PSS_GPUS')
    assert rows == list(pysamstats.stat_coverage(fx('test.bam'), chrom='Pf3D7_01_v3', start=1000, end=20000))


def test_binned_windows_across_gpus(obam, ofa, monkeypatch):
    # the *_binned family behind PSS_GPUS: contiguous window ranges, reads by start window, columns side by side
    cases = (dict(), dict(chrom='Pf3D7_01_v3', start=1000, end=20000, window_size=200, window_offset=100),
             dict(chrom='Pf3D7_02_v3', start=777, end=50001, window_size=77, min_mapq=1, no_dup=True),
             dict(window_size=1000, one_based=True), dict(chrom='Pf3D7_03_v3', start=10, end=40, window_size=300))
    for stat in config.binned_types:
        f = getattr(pysamstats, 'load_%s_binned' % stat)
        monkeypatch.delenv('PSS_GPUS', raising=False)
        expect = [f(fx('test.bam'), fafile=fx('ref.fa'), **kw) for kw in cases]
        assert_recarrays_equal(orc.load_binned(stat, obam['test'], ofa, **cases[2]), expect[2], stat)
        monkeypatch.setenv('PSS_GPUS', _gpu_list() + ',0')            # three shards
        for kw, e in zip(cases, expect):
            assert_recarrays_equal(e, f(fx('test.bam'), fafile=fx('ref.fa'), **kw), '%s_binned %r sharded' % (stat, kw))


# ---- load_*: records packed by the engine (pss_pileup_host_rows) vs the column tables --------------

@pytest.mark.parametrize('stat', ['coverage', 'variation_strand', 'baseq_ext', 'tlen_strand', 'mapq', 'coverage_gc'])
def test_rows_path_equals_column_path(stat, monkeypatch):
    from pysamstats_b200 import opt
    calls = []
    real = opt.pileup_rows
    monkeypatch.setattr(opt, 'pileup_rows', lambda *a, **k: calls.append(1) or real(*a, **k))
    cases = [dict(), dict(chrom='Pf3D7_01_v3', start=1000, end=20000, pad=True),
             dict(chrom='Pf3D7_02_v3', start=500, end=3000, truncate=True, one_based=True),
             dict(chrom='Pf3D7_03_v3', pad=True, min_mapq=20, no_dup=True),
             dict(fields=['pos', 'reads_all', 'chrom']), dict(chrom='Pf3D7_01_v3', start=49000, end=70000),
             dict(chrom='Pf3D7_01_v3', start=3, end=40, pad=True, truncate=True),          # (no read starts this early)
             dict(chrom='Pf3D7_02_v3', start=100, end=100, pad=True)]
    for chunk, table_chunk in ((None, None), ('1024', None), ('512', '4096')):
        for kw in cases:
            # PSS_HOST_CHUNK: result chunks inside one engine call; PSS_TABLE_CHUNK: position pieces = engine calls
            for var, val in (('PSS_HOST_CHUNK', chunk), ('PSS_TABLE_CHUNK', table_chunk)):
                if val:
                    monkeypatch.setenv(var, val)
                else:
                    monkeypatch.delenv(var, raising=False)
            monkeypatch.delenv('PSS_NO_ROWS', raising=False)
            n0 = len(calls)
            a = _load(stat, 'test', **kw)
            assert len(calls) == n0 + 1                       # the rows path ran ...
            monkeypatch.setenv('PSS_NO_ROWS', '1')
            b = _load(stat, 'test', **kw)
            assert len(calls) == n0 + 1                       # ... and the column path did not use it
            assert a.dtype == b.dtype and a.shape == b.shape, (stat, kw)
            assert a.tobytes() == b.tobytes(), (stat, kw, chunk)


def test_rows_path_declines_what_it_cannot_write():
    # user dtype overrides and single fields take the numpy path, with the reference's typing rules
    a = pysamstats.load_coverage(fx('test.bam'), chrom='Pf3D7_01_v3', start=0, end=2000, dtype={'reads_all': 'i8'})
    assert a.dtype['reads_all'] == np.dtype('i8')
    b = pysamstats.load_coverage(fx('test.bam'), chrom='Pf3D7_01_v3', start=0, end=2000)
    assert np.array_equal(a['reads_all'], b['reads_all']) and np.array_equal(a['pos'], b['pos'])
    c = pysamstats.load_coverage(fx('test.bam'), chrom='Pf3D7_01_v3', start=0, end=2000, fields='reads_all')
    assert c.dtype == np.dtype('i4') and np.array_equal(c, b['reads_all'])


# ---- round-2 boundary fixes ---------------------------------------------------------------------------

def test_chrom_none_keeps_start_end_for_pad_and_binned(obam, ofa):
    # iter_pileup_padded / iter_binned hand the caller's start / end to every contig (opt.pyx:1927-1932, :2009-2015);
    # only the unpadded pileup drops them with chrom=None (opt.pyx:2111-2112)
    for stat in ('coverage', 'variation'):
        kw = dict(pad=True, start=1000, end=3000)
        a = _load(stat, 'test', **kw)
        assert_recarrays_equal(_oracle(stat, obam, ofa, 'test', **kw), a, '%s %r' % (stat, kw))
        kw = dict(start=1000, end=3000)                    # no pad: whole contigs
        assert_recarrays_equal(_oracle(stat, obam, ofa, 'test', **kw), _load(stat, 'test', **kw), '%s %r' % (stat, kw))
    a = pysamstats.load_coverage(fx('test.bam'), pad=True, start=1000, end=3000, truncate=True)
    assert len(a) == 2000 * len(obam['test'].references) and a['pos'].min() == 1000 and a['pos'].max() == 2999
    for stat in config.binned_types:
        kw = dict(start=1000, end=5000, window_size=250)
        e = orc.load_binned(stat, obam['test'], ofa, **kw)
        a = getattr(pysamstats, 'load_%s_binned' % stat)(fx('test.bam'), fafile=fx('ref.fa'), **kw)
        assert_recarrays_equal(e, a, '%s_binned %r' % (stat, kw))
        assert len(a) == 16 * len(obam['test'].references)


def test_region_errors_of_parse_region():
    # pysam's parse_region raises for start > end (after normalise_coords clamped end to the contig) and for a
    # negative start -- one_based=True with start=0 normalises to -1 (opt.pyx:2116)
    for f, kw in ((pysamstats.load_coverage, {}), (pysamstats.load_mapq_binned, {})):
        with pytest.raises(ValueError, match='start out of range'):
            f(fx('test.bam'), chrom='Pf3D7_01_v3', start=0, end=3000, one_based=True, **kw)
        with pytest.raises(ValueError, match='invalid coordinates'):
            f(fx('test.bam'), chrom='Pf3D7_01_v3', start=60000, end=70000, **kw)


def test_unknown_fasta_contig_raises(tmp_path):
    fa = tmp_path / 'other.fa'
    fa.write_text('>someone_else\nACGTACGT\n')
    with pytest.raises(KeyError):
        pysamstats.load_variation(fx('test.bam'), str(fa), chrom='Pf3D7_01_v3', start=1000, end=1100)


def test_write_hdf5_through_the_engine(tmp_path, monkeypatch):
    # the reference's own test (pysamstats/test/test_io.py:14-40) with the stand-in for PyTables (tests/fake_tables.py),
    # plus the values: the dataset equals load_*
    import sys
    import fake_tables
    monkeypatch.setitem(sys.modules, 'tables', fake_tables)
    label = 'AS2_scf7180000696055'
    bam = fx('longcontignames.bam')
    class Foreign(object):
        filename = bam
    for dtype, alignment, expect in ((None, Foreign(), len(label)), ({'chrom': 'S20'}, Foreign(), 20), ({'chrom': 'S20'}, bam, 20)):
        out = str(tmp_path / 'c.h5')
        pysamstats.io.write_hdf5('coverage', out, alignment, chrom=label, dtype=dtype)
        with fake_tables.open_file(out, mode='r') as h5:
            assert h5.root.data.dtype['chrom'].itemsize == expect
            a = h5.root.data.read()
        e = pysamstats.load_coverage(bam, chrom=label, dtype=dtype)
        assert a.dtype == e.dtype and np.array_equal(a, np.asarray(e))
    out = str(tmp_path / 'v.h5')
    pysamstats.io.write_hdf5('variation_strand', out, fx('test.bam'), fafile=fx('ref.fa'), chrom='Pf3D7_02_v3', start=500, end=9000,
                          hdf5_chunksize=1 << 16, hdf5_dataset='vs')
    pysamstats.io.write_hdf5('tlen_binned', out, fx('test.bam'), hdf5_dataset='tb')
    with fake_tables.open_file(out, mode='r') as h5:
        e = pysamstats.load_variation_strand(fx('test.bam'), fx('ref.fa'), chrom='Pf3D7_02_v3', start=500, end=9000)
        assert np.array_equal(h5.root.vs.read(), np.asarray(e))
        assert np.array_equal(h5.root.tb.read(), np.asarray(pysamstats.load_tlen_binned(fx('test.bam'))))


# ---- decode -> upload -> compute pipeline (pieces decoded through the BAI while earlier pieces are on the GPU) ----------

@pytest.fixture(scope='module')
def synthetic_bam(tmp_path_factory):
    from pysamstats_b200 import bamio, synth
    G = 1_500_000
    batch, (ref, ref_lo) = synth.make_interval_batch(0, G, G, depth=12, block=500000, with_qual=True)
    nb = synth.batch_to_numpy(batch)
    d = tmp_path_factory.mktemp('stream')
    path, fa = str(d / 's.bam'), str(d / 's.fa')
    bamio.write_bam(path, ['chrS', 'chrT'], [G, 1000], nb, level=1)
    bamio.index_bam(path)
    with open(fa, 'wb') as fh:
        fh.write(b'>chrS\n' + ref.numpy()[:G].tobytes() + b'\n>chrT\n' + b'A' * 1000 + b'\n')
    return path, fa, G


def test_streamed_decode_equals_whole_file_decode(synthetic_bam, monkeypatch):
    from pysamstats_b200 import bamio, opt
    path, fa, G = synthetic_bam
    monkeypatch.setenv('PSS_BAI_MIN_BYTES', '0')
    monkeypatch.setenv('PSS_TABLE_CHUNK', '131072')
    cases = [('variation_strand', dict(chrom='chrS')), ('coverage_ext', dict(chrom='chrS', start=200_000, end=1_100_000)),
             ('coverage', dict(chrom='chrS', start=100_003, end=900_001, truncate=True)),
             ('coverage_strand', dict(chrom='chrS', start=5, end=1_400_000, pad=True)),
             ('variation', dict(chrom='chrS', start=300_000, end=1_000_000, pad=True, truncate=True, one_based=True)),
             ('mapq', dict(chrom='chrS', min_mapq=20, no_dup=True)), ('tlen', dict(chrom='chrS', start=1, end=700_000)),
             ('coverage', dict(chrom='chrS', max_depth=7)),               # the depth cap acts: no cut after the first piece
             ('coverage_ext', dict()),                                    # every contig
             ('baseq', dict(chrom='chrS', start=0, end=400_000))]          # reads qualities: never streamed
    streamed = []
    real = opt._streamed_pieces
    monkeypatch.setattr(opt, '_streamed_pieces', lambda *a, **k: streamed.append(1) or real(*a, **k))
    for no_rows in (False, True):
        for stat, kw in cases:
            f = getattr(pysamstats, 'load_' + stat)
            fasta = fa if stat in WITHREF else None
            monkeypatch.setenv('PSS_NO_STREAM', '1')
            bamio._bam_cache.clear()
            monkeypatch.delenv('PSS_NO_ROWS', raising=False)
            e = f(path, fafile=fasta, **kw)
            monkeypatch.delenv('PSS_NO_STREAM')
            if no_rows:
                monkeypatch.setenv('PSS_NO_ROWS', '1')
            bamio._bam_cache.clear()
            n0 = len(streamed)
            a = f(path, fafile=fasta, **kw)
            assert (len(streamed) > n0) == (stat != 'baseq'), (stat, kw)
            assert_recarrays_equal(e, a, 'streamed %s %r rows=%r' % (stat, kw, not no_rows))
            b = f(path, fafile=fasta, **kw)                        # second call: the cached pieces
            assert_recarrays_equal(e, b, 'cached pieces %s %r' % (stat, kw))
    # the generator of dicts over streamed tables
    bamio._bam_cache.clear()
    rows = list(pysamstats.stat_coverage(path, chrom='chrS', start=100_000, end=400_000))
    monkeypatch.setenv('PSS_NO_STREAM', '1')
    bamio._bam_cache.clear()
    assert rows == list(pysamstats.stat_coverage(path, chrom='chrS', start=100_000, end=400_000))
