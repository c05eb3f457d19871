"""CPU: host-side logic of the product (decoder, registry, API surface, C-ABI export list)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import pysamstats_b200 as pysamstats
from pysamstats_b200 import _cabi, bamio, config, opt
from oracle import bamread
from _util import fx, ROOT


@pytest.fixture(params=['native', 'numpy'])
def decoder(request, monkeypatch):
    """Both host decoders: csrc/pss_host.cpp (libpss_host.so) and the numpy fallback."""
    monkeypatch.setattr(bamio, '_host_lib', None)
    if request.param == 'numpy':
        monkeypatch.setenv('PSS_NUMPY_DECODER', '1')
    else:
        monkeypatch.delenv('PSS_NUMPY_DECODER', raising=False)
        assert bamio.host_lib() is not None, 'libpss_host.so is not built'
    yield request.param
    bamio._host_lib = None


@pytest.mark.parametrize('name', ['test.bam', 'deep.bam', 'rna.bam', 'longcontignames.bam'])
def test_decoder_matches_oracle_reader(name, decoder):
    mine = bamio.BamFile(fx(name)).load()
    assert (mine.decode_seconds is not None) == (decoder == 'native')
    ref = bamread.Bam(fx(name))
    assert list(mine.references) == ref.references and list(mine.lengths) == ref.lengths
    A = mine._all
    assert np.array_equal(mine._tid, ref.tid)
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq'):
        assert np.array_equal(getattr(A, k), getattr(ref, k)), k
    assert np.array_equal(A.endpos, ref.endpos)
    assert np.array_equal(A.cigar, np.concatenate(ref.cigar))
    assert np.array_equal(A.seq4, np.concatenate(ref.seq4))
    assert np.array_equal(A.qual, np.concatenate(ref.qual))
    assert [bytes(x) for x in A.names] == ref.qname


def test_bam_writer_roundtrip(tmp_path):
    from pysamstats_b200 import synth
    G = 200000
    batch, _ = synth.make_interval_batch(0, G, G, block=100000, with_qual=True)
    nb = synth.batch_to_numpy(batch)
    path = str(tmp_path / 'synthetic.bam')
    bamio.write_bam(path, ['chrS', 'chrT'], [G, 1000], nb, level=1, threads=2)
    ref = bamread.Bam(path)                                  # independent (gzip module) reader
    assert ref.references == ['chrS', 'chrT'] and len(ref.pos) == nb['n_reads']
    assert np.array_equal(ref.pos, nb['pos']) and np.array_equal(ref.flag, nb['flag'])
    assert np.array_equal(np.concatenate(ref.seq4), nb['seq4']) and np.array_equal(np.concatenate(ref.qual), nb['qual'])
    mine = bamio.BamFile(path).contig_batch('chrS')
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'cigar', 'seq_off', 'seq4', 'qual_off',
              'qual', 'name_prev'):
        assert np.array_equal(getattr(mine, k), nb[k]), k
    assert np.array_equal(mine.endpos, nb['endpos'])
    with pytest.raises(ValueError):
        bamio.BamFile(fx('ref.fa'))                          # not BGZF


def test_name_prev_native_equals_numpy(monkeypatch):
    names = np.array([b'a', b'bb', b'a', b'c', b'bb', b'a', b'', b''], dtype='S4')
    expect = [-1, -1, 0, -1, 1, 2, -1, 6]
    assert bamio._name_prev(names).tolist() == expect
    monkeypatch.setattr(bamio, '_host_lib', None)
    monkeypatch.setenv('PSS_NUMPY_DECODER', '1')
    assert bamio._name_prev(names).tolist() == expect
    bamio._host_lib = None


def test_region_batch_rule():
    mine = bamio.BamFile(fx('test.bam'))
    ref = bamread.Bam(fx('test.bam'))
    for (c, s, e) in (('Pf3D7_01_v3', 2000, 2100), ('Pf3D7_02_v3', 0, 60001), ('Pf3D7_03_v3', 0, 70001),
                      ('Pf3D7_01_v3', 924, 925), ('Pf3D7_01_v3', 9999, 12000)):
        b = mine.region_batch(c, s, e)
        idx = ref.region_reads(c, s, e)
        assert b.n_reads == len(idx)
        assert np.array_equal(b.pos, ref.pos[idx])
        assert np.array_equal(b.cigar_off[1:] - b.cigar_off[:-1], [len(ref.cigar[i]) for i in idx])
        # name_prev links the previous record of the same QNAME inside the batch
        names = [ref.qname[i] for i in idx]
        last = {}
        for j, nm in enumerate(names):
            assert b.name_prev[j] == last.get(nm, -1)
            last[nm] = j


BATCH_KEYS = ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'endpos', 'cigar_off', 'cigar', 'seq_off', 'seq4', 'qual_off',
              'qual', 'names', 'name_prev')


def _same_batch(a, b, label):
    assert a.n_reads == b.n_reads, (label, a.n_reads, b.n_reads)
    for k in BATCH_KEYS:
        x, y = getattr(a, k), getattr(b, k)
        if k in ('cigar', 'seq4', 'qual') and a.n_reads == 0:
            continue
        assert np.array_equal(x, y), (label, k)


@pytest.mark.parametrize('name', ['test.bam', 'deep.bam', 'rna.bam', 'longcontignames.bam'])
def test_indexed_region_decode_equals_whole_file_decode(name, monkeypatch):
    # the BAI path (pss_bam_open_region: only the BGZF blocks the index lists) against the in-memory filter
    monkeypatch.setenv('PSS_BAI_MIN_BYTES', '0')
    whole = bamio.BamFile(fx(name)).load()
    rng = np.random.default_rng(3)
    refs = list(zip(whole.references, whole.lengths))
    picks = [refs[i] for i in rng.choice(len(refs), min(len(refs), 6), replace=False)]
    for chrom, ln in picks:
        for _ in range(4):
            s = int(rng.integers(0, max(1, ln)))
            e = s + int(rng.choice([1, 50, 1000, 40000]))
            via_index = bamio.BamFile(fx(name))
            assert via_index.index_path() is not None
            got = via_index.region_batch(chrom, s, e)
            assert via_index._all is None                     # the whole file was never decoded
            _same_batch(got, whole.region_batch(chrom, s, e), (name, chrom, s, e))


def test_index_build_and_region_decode_on_a_multi_block_bam(tmp_path, monkeypatch):
    from pysamstats_b200 import synth
    monkeypatch.setenv('PSS_BAI_MIN_BYTES', '0')
    G = 400000
    batch, _ = synth.make_interval_batch(0, G, G, block=100000, with_qual=True)
    nb = synth.batch_to_numpy(batch)
    path = str(tmp_path / 'synthetic.bam')
    bamio.write_bam(path, ['chrS', 'chrT'], [G, 1000], nb, level=1, threads=2)
    assert bamio.index_bam(path, threads=2) == path + '.bai'
    assert os.path.getsize(path) > 40 * 65536                  # many BGZF blocks, chunks that span them
    whole = bamio.BamFile(path).load()
    rng = np.random.default_rng(5)
    for _ in range(12):
        s = int(rng.integers(0, G))
        e = s + int(rng.choice([1, 300, 20000, 150000]))
        f = bamio.BamFile(path)
        got = f.region_batch('chrS', s, e)
        assert f._all is None
        _same_batch(got, whole.region_batch('chrS', s, e), (s, e))
        if e - s <= 300:                                       # a small region touches a small part of the file
            assert f.decode_seconds['raw_bytes'] < 0.2 * whole.decode_seconds['raw_bytes']
    assert bamio.BamFile(path).region_batch('chrT', 0, 1000).n_reads == 0
    # the index this package writes and the one samtools wrote select the same records
    own = str(tmp_path / 'test.own.bai')
    bamio.index_bam(fx('test.bam'), index_path=own)
    a = bamio.BamFile(fx('test.bam'))
    b = bamio.BamFile(fx('test.bam'))
    b.index_path = lambda: own
    for (c, s, e) in (('Pf3D7_01_v3', 2000, 2100), ('Pf3D7_02_v3', 0, 60001), ('Pf3D7_03_v3', 100, 70001)):
        _same_batch(a.region_batch(c, s, e), b.region_batch(c, s, e), (c, s, e))
    with pytest.raises(ValueError):
        bamio.index_bam(fx('ref.fa'), index_path=str(tmp_path / 'x.bai'))


def test_fasta_fetch():
    fa = bamio.FastaFile(fx('ref.fa'))
    ofa = bamread.Fasta(fx('ref.fa'))
    for c in ('Pf3D7_01_v3', 'Pf3D7_02_v3', 'Pf3D7_03_v3'):
        assert fa.contig(c).tobytes() == ofa.contig(c)
    assert fa.fetch('Pf3D7_01_v3', 640850, 640900) == ofa.fetch('Pf3D7_01_v3', 640850, 640900)
    assert fa.fetch('Pf3D7_01_v3', 700000, 700010) == b''


def test_public_api_surface():
    for t in config.pileup_types:
        assert getattr(pysamstats, 'stat_' + t).__name__ == 'stat_' + t
        assert getattr(pysamstats, 'load_' + t).__name__ == 'load_' + t
        assert hasattr(config, 'dtype_' + t)
    for t in config.binned_types:
        assert callable(getattr(pysamstats, 'stat_%s_binned' % t))
        assert callable(getattr(pysamstats, 'load_%s_binned' % t))
    sig = inspect.signature(pysamstats.stat_pileup)
    assert list(sig.parameters) == ['type', 'alignmentfile', 'fafile', 'chrom', 'start', 'end', 'one_based',
                                    'truncate', 'stepper', 'pad', 'max_depth', 'window_size', 'window_offset',
                                    'min_mapq', 'min_baseq', 'no_del', 'no_dup']
    assert sig.parameters['max_depth'].default == 8000 and sig.parameters['stepper'].default == 'all'
    sig = inspect.signature(pysamstats.load_binned)
    assert list(sig.parameters) == ['type', 'alignmentfile', 'fafile', 'chrom', 'start', 'end', 'one_based',
                                    'window_size', 'window_offset', 'min_mapq', 'no_dup', 'dtype', 'fields']


def test_errors_without_gpu():
    # argument errors are raised before the engine is touched (pileup.py:80-89, binned.py:63-66)
    with pytest.raises(ValueError):
        pysamstats.stat_pileup('variation', fx('test.bam'))
    with pytest.raises(ValueError):
        pysamstats.stat_pileup('nosuchstat', fx('test.bam'))
    with pytest.raises(ValueError):
        pysamstats.load_pileup('nosuchstat', fx('test.bam'))
    with pytest.raises(ValueError):
        pysamstats.stat_binned('nosuchstat', fx('test.bam'))
    with pytest.raises(AttributeError):
        pysamstats.load_binned('nosuchstat', fx('test.bam'))
    with pytest.raises(ValueError):
        pysamstats.load_coverage(fx('test.bam'), fields=['nosuchfield'])


def test_normalise_coords():
    bam = bamio.open_bam(fx('test.bam'))
    assert opt.normalise_coords(bam, None, 5, 10, False) == (None, None)
    assert opt.normalise_coords(bam, 'Pf3D7_01_v3', None, None, False) == (0, 50001)
    assert opt.normalise_coords(bam, 'Pf3D7_01_v3', 10, 99999, False) == (10, 50001)
    assert opt.normalise_coords(bam, 'Pf3D7_01_v3', 10, 20, True) == (9, 19)
    with pytest.raises(AssertionError):
        opt.normalise_coords(bam, 'nochrom', None, None, False)


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'pss_b200.h')).read()
    declared = set(re.findall(r'\b(pss_[a-z_0-9]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)      # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    # shape helpers are pure host code
    lib.pss_pileup_nfields.restype = ctypes.c_int32
    for i, t in enumerate(_cabi.PILEUP_STATS):
        assert lib.pss_pileup_nfields(i) == len(config.value_fields(t))
    for i, t in enumerate(_cabi.BINNED_STATS):
        assert lib.pss_binned_nfields(i) == len(config.value_fields_binned(t))
    lib.pss_binned_nbins.restype = ctypes.c_int64
    lib.pss_binned_nbins.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int32]
    assert lib.pss_binned_nbins(0, 50001, 300) == 167
    assert lib.pss_binned_nbins(1000, 20000, 200) == 95
    assert lib.pss_binned_nbins(5, 5, 300) == 1


def test_native_row_formatter_equals_csv_writer(monkeypatch):
    # pss_format_rows (libpss_host): the path write_csv takes when no cell needs quoting -- int32 extremes, int64
    # positions, empty ref cells, a chrom that csv has to quote, several 4096-row blocks
    import csv
    import io
    from pysamstats_b200 import io as pio
    calls = []
    real = bamio.format_rows_native
    monkeypatch.setattr(bamio, 'format_rows_native', lambda *a: calls.append(1) or real(*a))
    n = 9001
    rng = np.random.default_rng(1)
    pos = np.arange(10 ** 6 - 5, 10 ** 6 - 5 + n, dtype=np.int64)
    pos[:3] = [0, 2 ** 40, 9999999999]
    ref = rng.choice(np.array([b'A', b'c', b'', b'N'], dtype='S1'), n)
    fields = ['chrom', 'pos', 'ref'] + ['f%d' % i for i in range(20)]
    cols = {f: rng.integers(0, 10 ** int(rng.integers(1, 10)), n).astype('i4') for f in fields[3:]}
    cols['f0'][:8] = [-2 ** 31, 2 ** 31 - 1, 0, -1, 9, 10, 99, 100]
    for chrom in ('Pf3D7_01_v3', 'odd,na\tme'):
        for dialect in ('excel-tab', 'excel', 'unix'):
            buf = io.StringIO()
            w = csv.writer(buf, dialect=dialect)
            for i in range(n):
                w.writerow([chrom, int(pos[i]), ref[i].decode()] + [int(cols[f][i]) for f in fields[3:]])
            got, exp = pio.format_rows(fields, chrom, pos, ref, cols, dialect), buf.getvalue()
            same = got == exp                      # (no pytest diff of megabyte strings)
            assert same, (chrom, dialect, next((i for i, (a, b) in enumerate(zip(got, exp)) if a != b), min(len(got), len(exp))))
    assert len(calls) == 4          # 'unix' quotes every cell (QUOTE_ALL): csv.writer itself


def test_host_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'pss_host.h')).read()
    declared = set(re.findall(r'\b(pss_[a-z_0-9]+)\s*\(', header))
    assert {'pss_bam_open', 'pss_bam_fill', 'pss_name_prev', 'pss_bam_write'} <= declared
    lib = ctypes.CDLL(bamio.HOST_LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert ctypes.sizeof(bamio.HostArrays) == 16 * 8


def test_struct_layouts_match_header():
    # sizes the C side expects (checked again on the GPU by every call)
    assert ctypes.sizeof(_cabi.PssReads) == 4 * 8 + 8 + 13 * 8
    assert ctypes.sizeof(_cabi.PssPileupParams) == 10 * 4 + 6 * 8
    assert ctypes.sizeof(_cabi.PssBinnedParams) == 4 * 4 + 6 * 8


def test_csv_block_formatter_equals_csv_writer():
    import csv
    import io
    from pysamstats_b200.io import format_rows
    n = 50
    rng = np.random.default_rng(0)
    pos = np.arange(1000, 1000 + n)
    ref = np.array(list('ACGTN",\t') * 8, dtype='S1')[:n]
    cols = {'reads_all': rng.integers(0, 5000, n).astype('i4'), 'gc': rng.integers(0, 100, n).astype('u1')}
    for chrom in ('Pf3D7_01_v3', 'odd,name', 'tab\tname', 'quo"te'):
        for dialect in ('excel-tab', 'excel'):
            fields = ['chrom', 'pos', 'ref', 'reads_all', 'gc']
            buf = io.StringIO()
            w = csv.writer(buf, dialect=dialect)
            for i in range(n):
                w.writerow([chrom, int(pos[i]), ref[i].decode(), int(cols['reads_all'][i]), int(cols['gc'][i])])
            assert format_rows(fields, chrom, pos, ref, cols, dialect) == buf.getvalue(), (chrom, dialect)
    assert format_rows(['pos'], 'x', pos[:0], None, {}, 'excel') == ''


def test_native_row_packer_and_column_gather_equal_numpy():
    # load_* assembles its packed structured array through libpss_host (pss_pack_rows) and compacts the
    # engine's per-position columns with pss_take_cols_i32: same bytes as the numpy field-by-field path
    from pysamstats_b200 import util
    rng = np.random.default_rng(11)
    names = config.value_fields('variation_strand')
    n = 70001
    dt = np.dtype([('chrom', 'S11'), ('pos', 'i4'), ('ref', 'S1')] + [(f, 'i4') for f in names])
    assert dt.itemsize == 256                                   # the packed layout of regression.npz
    cols = {f: rng.integers(0, 1 << 20, n).astype('i4') for f in names}
    pos = rng.integers(0, 1 << 30, n).astype('i8')
    ref = rng.choice([b'A', b'C', b'', b'N'], n).astype('S1')
    a = np.empty(n + 5, dtype=dt)
    a[:] = a[0]
    assert util._pack_native(a, 3, n, list(dt.names), 'Pf3D7_01_v3', pos, ref, cols)
    b = np.zeros(n, dtype=dt)
    b['chrom'], b['pos'], b['ref'] = b'Pf3D7_01_v3', pos, ref
    for f in names:
        b[f] = cols[f]
    assert a[3:3 + n].tobytes() == b.tobytes()
    # a user dtype the packer cannot copy byte for byte falls back to numpy
    dt2 = np.dtype([('chrom', 'S4'), ('pos', 'i8'), ('reads_all', 'i4')])
    assert not util._pack_native(np.empty(n, dtype=dt2), 0, n, list(dt2.names), 'c', pos, None, cols)
    out = rng.integers(0, 100, (7, 1 << 20)).astype('i4')
    sel = np.flatnonzero(rng.random(out.shape[1]) < 0.7)
    assert np.array_equal(bamio.take_cols(out, sel), out[:, sel])
    big = bamio.alloc_big((3, 12_000_000), np.int32)            # huge-page backed
    big[:] = 5
    assert big.shape == (3, 12_000_000) and int(big.sum()) == 5 * big.size


def test_position_chunks_cover_every_overlapping_read(monkeypatch):
    # pileup_tables cuts a contig into position pieces with zero-copy sub-batch views: every read overlapping a piece
    # must be inside its view, pieces must tile [lo, hi), and the range must not be cut when max_depth could bite or
    # base qualities are read
    import ctypes as C
    from pysamstats_b200 import synth
    G = 300000
    batch, _ = synth.make_interval_batch(0, G, G, block=100000, with_qual=True)
    nb = synth.batch_to_numpy(batch)
    b = bamio.ReadBatch()
    for k in bamio.ReadBatch.__slots__:
        if k != '_cache':
            setattr(b, k.lstrip('_'), nb.get(k.lstrip('_')))  # (name_prev is a lazily hashed property over a slot)
    b.tid = 0
    b.names = None
    monkeypatch.setenv('PSS_TABLE_CHUNK', '20000')
    lo, hi = 1000, G - 500
    pieces = list(opt._position_chunks(b, lo, hi, False, 8000))
    assert len(pieces) > 10 and pieces[0][0] == lo and pieces[-1][1] == hi
    assert all(p[1] == q[0] for p, q in zip(pieces, pieces[1:]))
    addr0 = b.pos.ctypes.data
    for c_lo, c_hi, rs in pieces:
        i0 = (C.cast(rs.pos, C.c_void_p).value - addr0) // 4
        i1 = i0 + rs.n_reads
        want = np.flatnonzero((b.pos < c_hi) & (b.endpos > c_lo))
        assert len(want) == 0 or (i0 <= want[0] and want[-1] < i1), (c_lo, c_hi)
        assert rs.qual is None or C.cast(rs.qual, C.c_void_p).value is None      # no qualities requested
    # max_depth within reach of the local read density: one piece, with or without qualities
    assert len(list(opt._position_chunks(b, lo, hi, False, 20))) == 1
    assert len(list(opt._position_chunks(b, lo, hi, True, 20, pair_links=lambda *a: None))) == 1
    # qualities wanted: the same pieces, each with its qualities and with the batch-wide pairs of the mate-overlap
    # rewrite as two-read name chains in the piece's own indices (here a stand-in for pss_pairs_host: every read
    # paired with the previous read of its name)
    links = np.asarray(b.name_prev, dtype=np.int32)
    qpieces = list(opt._position_chunks(b, lo, hi, True, 8000, pair_links=lambda *a: links))
    assert [(p[0], p[1]) for p in qpieces] == [(p[0], p[1]) for p in pieces]
    n_cross = 0
    for (c_lo, c_hi, rs), (_, _, rs0) in zip(qpieces, pieces):
        i0 = (C.cast(rs.pos, C.c_void_p).value - addr0) // 4
        i1 = i0 + rs.n_reads
        assert (i0, rs.n_reads) == ((C.cast(rs0.pos, C.c_void_p).value - addr0) // 4, rs0.n_reads)
        assert C.cast(rs.qual, C.c_void_p).value == b.qual.ctypes.data
        assert C.cast(rs.qual_off, C.c_void_p).value == b.qual_off.ctypes.data + 8 * i0
        assert rs.n_qual_bytes == int(b.qual_off[i1] - b.qual_off[i0])
        prev = np.ctypeslib.as_array(C.cast(rs.name_prev, C.POINTER(C.c_int32)), shape=(rs.n_reads,))
        want = np.where(links[i0:i1] >= i0, links[i0:i1] - i0, -1)
        assert np.array_equal(prev, want)
        n_cross += int(((links[i0:i1] >= 0) & (links[i0:i1] < i0)).sum())
        assert prev.max(initial=-1) < rs.n_reads and np.all(prev < np.arange(rs.n_reads))
    assert n_cross > 0          # some pairs do straddle a piece's first read: their links are dropped there


def test_row_sink_packer_cpp(tmp_path):
    # the packer behind pss_pileup_host_rows, without a GPU: tests/cpp/pack_rows_test.cpp against csrc/pss_expand.cpp
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, 'pysamstats_b200', 'csrc')
    exe = str(tmp_path / 'pack_rows_test')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-I', csrc, '-o', exe, os.path.join(root, 'tests', 'cpp', 'pack_rows_test.cpp'),
                           os.path.join(csrc, 'pss_expand.cpp'), '-lpthread'])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count('bad 0') == 21


def test_interval_maximum_structure_cpp(tmp_path):
    # the max(mapq) structure of k_tilem (csrc/pss_imax.cuh: at most four atomic maxima per interval, resolved per
    # block) and the kept-run enumeration under min_baseq (csrc/pss_lowq.cuh), compiled for the host and run against
    # brute force: tests/cpp/imax_test.cpp
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, 'pysamstats_b200', 'csrc')
    exe = str(tmp_path / 'imax_test')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-I', csrc, '-o', exe, os.path.join(root, 'tests', 'cpp', 'imax_test.cpp')])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and 'imax_test: ok' in out.stdout, out.stdout


def test_own_inflate_equals_zlib():
    # pss_inflate.cpp (the BGZF decode path): every stream zlib writes -- stored / fixed / dynamic blocks, flush points,
    # long distances, short runs -- inflates to the same bytes; streams it accepts after corruption are streams zlib
    # accepts with the same bytes; wrong output sizes are refused
    import zlib
    lib = ctypes.CDLL(bamio.HOST_LIB_PATH)
    lib.pss_inflate_raw.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    rng = np.random.default_rng(5)

    def own(comp, n, mode=1):
        out = ctypes.create_string_buffer(max(n, 1))
        rc = lib.pss_inflate_raw(comp, len(comp), out, n, mode)
        return rc, out.raw[:n]

    datasets = [b'', b'a', b'ab' * 40000, bytes(65280), rng.integers(0, 256, 65280, dtype=np.uint8).tobytes(),
                rng.choice(np.frombuffer(b'ACGT', dtype=np.uint8), 65280).tobytes(),
                np.clip(rng.normal(35, 5, 65280), 2, 41).astype(np.uint8).tobytes()]
    base = rng.integers(0, 256, 20000, dtype=np.uint8).tobytes()
    datasets.append(base + base[5000:15000] + base[:12345] + base)
    for n in (1, 2, 3, 8, 9, 257, 258, 259, 32768, 32769):
        datasets.append(rng.integers(0, 7, n, dtype=np.uint8).tobytes())
    n_corrupt = 0
    for data in datasets:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
                half = len(data) // 2
                comp = co.compress(data[:half]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(data[half:]) + co.flush()
                assert own(comp, len(data)) == (0, data), (len(data), level, strategy)
                assert own(comp, len(data), 0) == (0, data)
                assert own(comp, len(data) + 1)[0] != 0 and (not data or own(comp, len(data) - 1)[0] != 0)
                if len(comp) > 8:
                    bad = bytearray(comp)
                    bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
                    bad = bytes(bad[:int(rng.integers(len(bad) // 2, len(bad) + 1))])
                    rc, got = own(bad, len(data))
                    if rc == 0:
                        n_corrupt += 1
                        d = zlib.decompressobj(-15)
                        assert d.decompress(bad) == got and d.eof      # (a flipped literal bit is still a valid stream)
    # fixture BAMs: every BGZF block through both decoders
    import struct
    for name in ('test.bam', 'rna.bam'):
        raw = open(fx(name), 'rb').read()
        p = 0
        while p < len(raw):
            xlen, = struct.unpack_from('<H', raw, p + 10)
            bsize = struct.unpack_from('<H', raw, p + 16)[0] + 1
            pay, isize = raw[p + 12 + xlen:p + bsize - 8], struct.unpack_from('<I', raw, p + bsize - 4)[0]
            assert own(pay, isize) == (0, zlib.decompress(pay, -15))
            p += bsize


def test_region_batch_equals_brute_force_overlap():
    # BamFile.region_batch narrows the start-sorted contig by binary search (longest span behind the region start)
    # and gathers only the selected records: same records, same ragged payloads as a mask over the whole contig
    bam = bamio.BamFile(fx('test.bam'))
    rng = np.random.default_rng(3)
    for chrom in bam.references[:3]:
        whole = bam.contig_batch(chrom)
        clen = bam.lengths[bam.references.index(chrom)]
        for _ in range(25):
            start = int(rng.integers(-50, clen + 50))
            end = start + int(rng.integers(0, 3000))
            sub = bam.region_batch(chrom, start, end)
            want = np.flatnonzero((whole.pos < end) & (whole.endpos > start))
            assert sub.n_reads == len(want), (chrom, start, end)
            assert np.array_equal(sub.pos, whole.pos[want]) and np.array_equal(sub.flag, whole.flag[want])
            assert np.array_equal(sub.names, whole.names[want])
            for data, off in (('cigar', 'cigar_off'), ('seq4', 'seq_off'), ('qual', 'qual_off')):
                o, so = getattr(whole, off), getattr(sub, off)
                if len(want):
                    exp = np.concatenate([getattr(whole, data)[o[i]:o[i + 1]] for i in want.tolist()])
                    assert np.array_equal(getattr(sub, data)[so[0]:so[-1]], exp), (chrom, start, end, data)
                    assert np.array_equal(np.diff(so), (o[want + 1] - o[want]))
            if sub.n_reads:
                assert np.array_equal(sub.name_prev >= 0, np.array([n in set(sub.names[:j].tolist()) for j, n in enumerate(sub.names.tolist())]))


def test_alloc_rows_is_lazy_address_space():
    # the record buffer load_* fills from the front: writable, right dtype, and a big request costs no memory until
    # it is touched (an untouched 8 GiB reservation must not fail or commit on this machine)
    dt = np.dtype([('chrom', 'S12'), ('pos', '<i4'), ('x', '<i4')])
    small = bamio.alloc_rows(1000, dt)
    assert small.dtype == dt and len(small) == 1000 and small.flags.writeable
    big = bamio.alloc_rows((8 << 30) // dt.itemsize, dt)
    big[:3]['pos'] = [1, 2, 3]
    part = big[:3]
    assert part['pos'].tolist() == [1, 2, 3] and part.base is not None
    del part, big
