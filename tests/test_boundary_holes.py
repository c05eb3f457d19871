"""CPU: the small holes at the python boundary that the round-1 review listed -- ``util.load_stats``
with the reference's ``statfun`` signature, ``write_hdf5`` against a stand-in for PyTables, shard
alignment / quality rule, FASTA errors, region decode with a negative start or a stale index."""
import os
import sys
import warnings

import numpy as np
import pytest

import pysamstats_b200 as pysamstats
from pysamstats_b200 import _cabi, bamio, config, io as pio, opt, shard, util
from _util import fx

import fake_tables


# ------------------------------------------------------------------------ util.load_stats

def _recs(alignmentfile, n=7, **kw):
    for i in range(n):
        yield {'chrom': 'Pf3D7_01_v3', 'pos': 10 + i, 'reads_all': i * 3, 'reads_pp': i}


def test_load_stats_takes_a_generator_of_dicts_like_the_reference():
    # /root/reference/pysamstats/util.py:31: load_stats(statfun, default_dtype, user_dtype, user_fields, **kwargs)
    a = util.load_stats(_recs, config.dtype_coverage, None, None, alignmentfile=fx('test.bam'))
    assert isinstance(a, np.recarray) and a.dtype.names == ('chrom', 'pos', 'reads_all', 'reads_pp')
    assert a.dtype['chrom'] == np.dtype('S11')                     # longest contig name of the file
    assert a.pos.tolist() == list(range(10, 17)) and a.reads_all.tolist() == [0, 3, 6, 9, 12, 15, 18]
    assert a.chrom[0] == b'Pf3D7_01_v3'
    b = util.load_stats(_recs, config.dtype_coverage, {'reads_all': 'i8'}, ['pos', 'reads_all'], alignmentfile=fx('test.bam'), n=3)
    assert b.dtype.names == ('pos', 'reads_all') and b.dtype['reads_all'] == np.dtype('i8') and len(b) == 3
    c = util.load_stats(_recs, config.dtype_coverage, None, ['reads_pp'], alignmentfile=fx('test.bam'))
    assert not isinstance(c, np.recarray) and c.dtype == np.dtype('i4') and c.tolist() == list(range(7))
    with pytest.raises(ValueError):
        util.load_stats(_recs, config.dtype_coverage, None, ['nope'], alignmentfile=fx('test.bam'))
    assert list(inspect_params(util.load_stats)) == ['statfun', 'default_dtype', 'user_dtype', 'user_fields', 'kwargs']


def inspect_params(f):
    import inspect
    return inspect.signature(f).parameters


# ------------------------------------------------------------------------------ write_hdf5

def _fake_tables_of(stats_type, alignmentfile, **kwargs):
    """Two contig tables with known values (no GPU): what opt.pileup_tables would yield."""
    names = config.value_fields(stats_type)
    for chrom, n, base in (('AS2_scf7180000696055', 1000, 0), ('AS2_scf7180000695891', 10, 5000)):
        pos = np.arange(n, dtype=np.int64) + base
        ref = np.full(n, b'A', dtype='S1')
        yield chrom, pos, ref, {f: (np.arange(n, dtype=np.int32) * (j + 1)) for j, f in enumerate(names)}


@pytest.fixture
def tables_standin(monkeypatch):
    monkeypatch.setitem(sys.modules, 'tables', fake_tables)
    monkeypatch.setattr(pio, '_tables', _fake_tables_of)
    return fake_tables


def test_write_hdf5_chrom_dtype_and_layout(tables_standin, tmp_path):
    # the reference's own test (pysamstats/test/test_io.py:14-40): chrom width from the header, or the override
    bam = fx('longcontignames.bam')
    label = 'AS2_scf7180000696055'
    for dtype, expect in ((None, len(label)), ({'chrom': 'a20'}, 20), ({'chrom': 'S20'}, 20)):
        out = str(tmp_path / ('c%d.h5' % expect))
        pio.write_hdf5('coverage', out, bam, chrom=label, dtype=dtype)
        with fake_tables.open_file(out, mode='r') as h5:
            assert h5.root.data.dtype['chrom'].itemsize == expect
    # layout: dataset under a created group, title, filters, chunk shape = chunksize // itemsize, chunked appends
    out = str(tmp_path / 'v.h5')
    pio.write_hdf5('variation', out, bam, hdf5_group='/grp/sub', hdf5_dataset='var', hdf5_chunksize=96 * 128,
                   hdf5_complevel=5, hdf5_complib='blosc', hdf5_shuffle=False, hdf5_fletcher32=True)
    with fake_tables.open_file(out, mode='r') as h5:
        t = h5.root.grp.sub.var
        it = t.dtype.itemsize
        assert t.dtype.names == tuple(n for n, _ in config.dtype_variation)
        assert t.title == 'variation' and t.chunkshape == (96 * 128 // it,)
        assert (t.filters.complevel, t.filters.complib, t.filters.shuffle, t.filters.fletcher32) == (5, 'blosc', False, True)
        a = t.read()
        assert len(a) == 1010 and a['pos'][-1] == 5009 and a['chrom'][-1] == b'AS2_scf7180000695891'
        assert np.array_equal(a['matches'][:1000], np.arange(1000) * 3)        # third value field
        assert (a['ref'] == b'A').all()
    # mode 'a': another dataset is kept, the same dataset is replaced (remove_node, io.py:183-186)
    pio.write_hdf5('coverage', out, bam)
    pio.write_hdf5('coverage', out, bam, fields=['pos', 'reads_all'])
    with fake_tables.open_file(out, mode='r') as h5:
        assert h5.root.data.dtype.names == ('pos', 'reads_all') and len(h5.root.data.read()) == 1010
        assert len(h5.root.grp.sub.var.read()) == 1010
    # a single field gives a plain column (io.py:163-164)
    one = str(tmp_path / 'one.h5')
    pio.write_hdf5('coverage', one, bam, fields=['reads_pp'])
    with fake_tables.open_file(one, mode='r') as h5:
        assert h5.root.data.dtype == np.dtype('i4') and h5.root.data.read()[:4].tolist() == [0, 2, 4, 6]


def test_write_hdf5_progress_lines(tables_standin, tmp_path, capsys):
    pio.write_hdf5('coverage', str(tmp_path / 'p.h5'), fx('longcontignames.bam'), progress=100, hdf5_chunksize=28 * 100)
    err = capsys.readouterr().err
    assert err.count('[pysamstats]') >= 2 and '1010 rows in' in err


# ----------------------------------------------------------------------------------- shards

def test_shard_tile_is_the_library_tile():
    lib = _cabi.load_library()
    assert shard.TILE == lib.pss_tile_positions() == 512
    for (lo, hi) in shard.plan_intervals(7, 250_000_007, 8)[:-1]:
        assert (hi - 7) % 512 == 0


def test_quality_requests_are_not_cut():
    parts = shard.plan_intervals(100, 10_000_000, 4, need_qual=True)
    assert parts[0] == (100, 10_000_000) and all(lo == hi for lo, hi in parts[1:])
    assert sum(hi - lo for lo, hi in parts) == 10_000_000 - 100


# ------------------------------------------------------------------------------ FASTA / BAI

def test_unknown_fasta_contig_raises_keyerror():
    fa = bamio.open_fasta(fx('ref.fa'))
    assert len(fa.contig('Pf3D7_01_v3')) > 0
    with pytest.raises(KeyError):
        fa.contig('not_a_contig')


def _indexed_copy(tmp_path, name='test.bam'):
    import shutil
    dst = str(tmp_path / name)
    shutil.copy(fx(name), dst)
    shutil.copy(fx(name + '.bai'), dst + '.bai')
    return dst


@pytest.mark.skipif(bamio.host_lib() is None, reason='libpss_host.so is not built')
def test_region_decode_with_negative_start(tmp_path, monkeypatch):
    # one_based=True, start=0 normalises to start = -1 (opt.pyx:2116); the index query must treat it as 0
    monkeypatch.setenv('PSS_BAI_MIN_BYTES', '0')
    path = _indexed_copy(tmp_path)
    a = bamio.BamFile(path).region_batch('Pf3D7_01_v3', -1, 5000)
    b = bamio.BamFile(path).region_batch('Pf3D7_01_v3', 0, 5000)
    c = bamio.BamFile(path).load().region_batch('Pf3D7_01_v3', -1, 5000)      # whole-file decode, no index
    assert a.n_reads == b.n_reads == c.n_reads > 0
    for k in ('pos', 'flag', 'mapq', 'cigar', 'seq4', 'qual'):
        assert np.array_equal(getattr(a, k), getattr(c, k)), k


@pytest.mark.skipif(bamio.host_lib() is None, reason='libpss_host.so is not built')
def test_stale_index_is_ignored_with_a_warning(tmp_path, monkeypatch):
    import shutil
    monkeypatch.setenv('PSS_BAI_MIN_BYTES', '0')
    path = _indexed_copy(tmp_path)
    shutil.copy(fx('longcontignames.bam.bai') if os.path.exists(fx('longcontignames.bam.bai')) else fx('deep.bam.bai'),
                path + '.bai')                                    # an index of another file (other number of references)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        a = bamio.BamFile(path).region_batch('Pf3D7_01_v3', 1000, 5000)
    assert any('ignoring BAM index' in str(x.message) for x in w)
    c = bamio.BamFile(fx('test.bam')).load().region_batch('Pf3D7_01_v3', 1000, 5000)
    assert a.n_reads == c.n_reads and np.array_equal(a.pos, c.pos)


# ------------------------------------------------------------------- chrom=None start / end

def test_chrom_none_start_end_semantics_in_the_plan(monkeypatch):
    # iter_pileup_default drops start / end when chrom is None (opt.pyx:2111-2112); iter_pileup_padded and
    # iter_binned forward them to every contig (opt.pyx:1927-1932, :2009-2015).  Checked on the plan only (no GPU).
    class _Ctx(object):
        class lib(object):
            pss_pileup_needs_qual = staticmethod(lambda s, q: 0)
            pss_pileup_needs_ref = staticmethod(lambda s: 0)
    monkeypatch.setattr(_cabi, 'get_context', lambda device=None: _Ctx())
    kw = dict(one_based=False, truncate=True, stepper='all', max_depth=8000, min_mapq=0, min_baseq=0, no_del=False,
              no_dup=False)
    plain = list(opt._pileup_plan(opt.Coverage(), fx('test.bam'), None, False, None, 1000, 2000, **kw))
    padded = list(opt._pileup_plan(opt.Coverage(), fx('test.bam'), None, True, None, 1000, 2000, **kw))
    assert len(padded) == len(bamio.open_bam(fx('test.bam')).references)
    assert all((p[1], p[2]) == (1000, 2000) for p in padded)
    lens = dict(zip(bamio.open_bam(fx('test.bam')).references, bamio.open_bam(fx('test.bam')).lengths))
    assert all((p[1], p[2]) == (0, lens[p[0]]) for p in plain)
    with pytest.raises(ValueError, match='invalid coordinates'):
        list(opt._pileup_plan(opt.Coverage(), fx('test.bam'), None, True, 'Pf3D7_01_v3', 3000, 2000, **kw))
    kw['one_based'] = True
    with pytest.raises(ValueError, match='start out of range'):
        list(opt._pileup_plan(opt.Coverage(), fx('test.bam'), None, False, 'Pf3D7_01_v3', 0, 2000, **kw))


# ------------------------------------------------------------------------ streamed piece decode

@pytest.mark.skipif(bamio.host_lib() is None, reason='libpss_host.so is not built')
def test_piece_batches_equal_region_queries(tmp_path, monkeypatch):
    from pysamstats_b200 import synth
    G = 600_000
    batch, _ = synth.make_interval_batch(0, G, G, depth=8, block=200000, with_qual=True)
    nb = synth.batch_to_numpy(batch)
    path = str(tmp_path / 'p.bam')
    bamio.write_bam(path, ['chrS'], [G], nb, level=1)
    bamio.index_bam(path)
    monkeypatch.setenv('PSS_BAI_MIN_BYTES', '0')
    bam = bamio.BamFile(path)
    assert bam.can_stream('chrS')
    cuts = [1000, 150_000, 150_512, 400_000, 599_000]
    whole = bamio.BamFile(path).load()
    pieces = list(bam.piece_batches('chrS', cuts))
    assert len(pieces) == len(cuts) - 1
    for k, b in enumerate(pieces):
        e = whole.region_batch('chrS', cuts[k], cuts[k + 1])
        assert b.n_reads == e.n_reads > 0
        for f in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'endpos', 'cigar', 'seq4'):
            assert np.array_equal(getattr(b, f), getattr(e, f)), (k, f)
        assert np.array_equal(np.diff(b.cigar_off), np.diff(e.cigar_off)) and np.array_equal(np.diff(b.seq_off), np.diff(e.seq_off))
        assert b.names is None and b.qual_off is None          # streamed pieces are decoded without qualities / names
    again = list(bam.piece_batches('chrS', cuts))                 # served from the cache: the same objects
    assert all(x is y for x, y in zip(pieces, again))
    monkeypatch.setenv('PSS_NO_STREAM', '1')
    assert not bam.can_stream('chrS')
    assert not whole.can_stream('chrS')                             # already decoded in this process
