"""TEST INFRASTRUCTURE ONLY -- python front end of the CPU parity oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this.  Wraps ``oracle/pileup_oracle.c`` (built by ``oracle/Makefile`` into
``oracle/_build/liboracle.so``) with the calling conventions of the reference's python
layer so that parity tests read like the reference's own:

* ``stat_pileup`` / ``load_pileup``  ~ reference ``pysamstats/pileup.py:53-144``
* ``stat_binned`` / ``load_binned``  ~ reference ``pysamstats/binned.py:41-115``
* coordinate handling              ~ reference ``pysamstats/opt.pyx:2105-2130``
* array assembly                   ~ reference ``pysamstats/util.py:31-73``

It has its own field tables and its own BAM reader (``oracle/bamread.py``): nothing here
is shared with the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import bamread

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_build', 'liboracle.so')

PILEUP_TYPES = ['coverage', 'coverage_strand', 'coverage_ext', 'coverage_ext_strand', 'variation',
                'variation_strand', 'tlen', 'tlen_strand', 'mapq', 'mapq_strand', 'baseq',
                'baseq_strand', 'baseq_ext', 'baseq_ext_strand', 'coverage_gc']
BINNED_TYPES = ['coverage', 'coverage_ext', 'mapq', 'alignment', 'tlen']
WITHREF = ('variation', 'variation_strand', 'baseq_ext', 'baseq_ext_strand', 'coverage_gc')
STEPPERS = {'all': 0, 'nofilter': 1, 'samtools': 2}

_S6 = ['', '_fwd', '_rev', '_pp', '_pp_fwd', '_pp_rev']
_S2 = ['', '_pp']
_S3 = ['', '_fwd', '_rev']
_EXT = ['reads_all', 'reads_pp', 'reads_mate_unmapped', 'reads_mate_other_chr',
        'reads_mate_same_strand', 'reads_faceaway', 'reads_softclipped', 'reads_duplicate']
_VAR = ['reads', 'matches', 'mismatches', 'deletions', 'insertions', 'A', 'C', 'T', 'G', 'N']


def _sfx(base, sfx):
    # 'reads' + '_pp' -> 'reads_pp'; the bare total of the read counter is 'reads_all'
    if base == 'reads' and sfx in ('', ):
        return 'reads_all'
    return base + sfx


def _fields(t):
    if t == 'coverage':
        return ['reads_all', 'reads_pp']
    if t == 'coverage_strand':
        return [_sfx('reads', s) for s in _S6]
    if t == 'coverage_ext':
        return list(_EXT)
    if t == 'coverage_ext_strand':
        out = []
        for b in _EXT:
            out += [b, b + '_fwd', b + '_rev'] if b != 'reads_all' else ['reads_all', 'reads_fwd', 'reads_rev']
        return out
    if t == 'variation':
        return [_sfx(b, s) for b in _VAR for s in _S2]
    if t == 'variation_strand':
        return [_sfx(b, s) for b in _VAR for s in _S6]
    if t == 'tlen':
        return ['reads_all', 'reads_paired', 'reads_pp', 'mean_tlen', 'mean_tlen_pp', 'rms_tlen',
                'rms_tlen_pp', 'std_tlen', 'std_tlen_pp']
    if t == 'tlen_strand':
        out = ['reads_all', 'reads_fwd', 'reads_rev']
        out += ['reads_paired' + s for s in _S3] + ['reads_pp' + s for s in _S3]
        for b in ('mean_tlen', 'rms_tlen', 'std_tlen'):
            out += [b + s for s in _S6]
        return out
    if t == 'mapq':
        return ['reads_all', 'reads_pp', 'reads_mapq0', 'reads_mapq0_pp', 'rms_mapq', 'rms_mapq_pp',
                'max_mapq', 'max_mapq_pp']
    if t == 'mapq_strand':
        return ([_sfx('reads', s) for s in _S6] + ['reads_mapq0' + s for s in _S6] +
                ['rms_mapq' + s for s in _S6] + ['max_mapq' + s for s in _S6])
    if t == 'baseq':
        return ['reads_all', 'reads_pp', 'rms_baseq', 'rms_baseq_pp']
    if t == 'baseq_strand':
        return [_sfx('reads', s) for s in _S6] + ['rms_baseq' + s for s in _S6]
    if t in ('baseq_ext', 'baseq_ext_strand'):
        sf = _S2 if t == 'baseq_ext' else _S6
        out = [_sfx(b, s) for b in ('reads', 'matches', 'mismatches') for s in sf]
        out += ['rms_baseq' + g + s for g in ('', '_matches', '_mismatches') for s in sf]
        return out
    if t == 'coverage_gc':
        return ['gc', 'reads_all', 'reads_pp']
    raise KeyError(t)


def _fields_binned(t):
    if t == 'coverage':
        return ['gc', 'reads_all', 'reads_pp']
    if t == 'coverage_ext':
        return ['gc'] + list(_EXT)
    if t == 'mapq':
        return ['reads_all', 'reads_mapq0', 'rms_mapq']
    if t == 'alignment':
        return ['reads_all', 'bases_all', 'M', 'I', 'D', 'N', 'S', 'H', 'P', '=', 'X']
    if t == 'tlen':
        return ['reads_all', 'reads_pp', 'mean_tlen', 'mean_tlen_pp', 'rms_tlen', 'rms_tlen_pp']
    raise KeyError(t)


class _Reads(C.Structure):
    _fields_ = [('n', C.c_int64), ('pos', C.c_void_p), ('tlen', C.c_void_p), ('mtid', C.c_void_p),
                ('l_seq', C.c_void_p), ('flag', C.c_void_p), ('mapq', C.c_void_p),
                ('cigar', C.c_void_p), ('cigar_off', C.c_void_p), ('seq4', C.c_void_p),
                ('seq_off', C.c_void_p), ('qual', C.c_void_p), ('qual_off', C.c_void_p),
                ('name_id', C.c_void_p), ('n_names', C.c_int64), ('tid', C.c_int32)]


class _Params(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ('stat', 'stepper', 'max_depth', 'truncate', 'pad', 'min_mapq',
                                          'min_baseq', 'no_del', 'no_dup', 'window_size',
                                          'window_offset', 'has_ref')] + \
               [('start', C.c_int64), ('end', C.c_int64)]


def build(force=False):
    """Compile the C restatement (gcc).  Building the checker is not using it."""
    src = os.path.join(HERE, 'pileup_oracle.c')
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src)):
        return LIB_PATH
    subprocess.check_call(['make', '-s', '-C', HERE])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_pileup.restype = C.c_int64
        _lib.orc_pileup.argtypes = [C.POINTER(_Reads), C.POINTER(_Params), C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        _lib.orc_binned.restype = C.c_int64
        _lib.orc_binned.argtypes = [C.POINTER(_Reads), C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_int64]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PackedReads(object):
    """SoA arrays for a list of records (keeps the numpy arrays alive)."""

    def __init__(self, tid, pos, flag, mapq, tlen, mtid, l_seq, cigars, seqs, quals, names):
        n = len(pos)
        self.n = n
        self.pos = np.ascontiguousarray(pos, dtype='i4')
        self.flag = np.ascontiguousarray(flag, dtype='u2')
        self.mapq = np.ascontiguousarray(mapq, dtype='u1')
        self.tlen = np.ascontiguousarray(tlen, dtype='i4')
        self.mtid = np.ascontiguousarray(mtid, dtype='i4')
        self.l_seq = np.ascontiguousarray(l_seq, dtype='i4')

        def cat(parts, dtype):
            off = np.zeros(n + 1, dtype='i8')
            if n:
                off[1:] = np.cumsum([len(p) for p in parts])
            data = np.concatenate(parts).astype(dtype) if n and off[-1] else np.zeros(1, dtype=dtype)
            return np.ascontiguousarray(data), off

        self.cigar, self.cigar_off = cat(cigars, 'u4')
        self.seq4, self.seq_off = cat(seqs, 'u1')
        self.qual, self.qual_off = cat(quals, 'u1')
        ids = {}
        self.name_id = np.array([ids.setdefault(q, len(ids)) for q in names], dtype='i8')
        self.n_names = len(ids)
        self.tid = tid

    def cstruct(self):
        r = _Reads()
        r.n = self.n
        for k in ('pos', 'tlen', 'mtid', 'l_seq', 'flag', 'mapq', 'cigar', 'cigar_off', 'seq4',
                  'seq_off', 'qual', 'qual_off', 'name_id'):
            setattr(r, k, _ptr(getattr(self, k)))
        r.n_names = self.n_names
        r.tid = self.tid
        return r


def pack_region(bam, chrom, start, end):
    idx = bam.region_reads(chrom, start, end)
    return PackedReads(bam.references.index(chrom), bam.pos[idx], bam.flag[idx], bam.mapq[idx],
                       bam.tlen[idx], bam.mtid[idx], bam.l_seq[idx],
                       [bam.cigar[i] for i in idx], [bam.seq4[i] for i in idx],
                       [bam.qual[i] for i in idx], [bam.qname[i] for i in idx])


def normalise_coords(bam, chrom, start, end, one_based):
    """reference opt.pyx:2105-2130."""
    if chrom is None:
        return None, None
    assert chrom in bam.references, 'chromosome not in SAM references: %s' % chrom
    if one_based:
        start = start - 1 if start is not None else None
        end = end - 1 if end is not None else None
    chrlen = bam.lengths[bam.references.index(chrom)]
    if start is None:
        start = 0
    if end is None:
        end = chrlen
    if end > chrlen:
        end = chrlen
    return start, end


def _open(bam, fafile):
    if isinstance(bam, str):
        bam = bamread.Bam(bam)
    if isinstance(fafile, str):
        fafile = bamread.Fasta(fafile)
    return bam, fafile


def pileup_packed(stat, reads, start, end, ref=None, stepper='all', max_depth=8000, truncate=False,
                  pad=False, min_mapq=0, min_baseq=0, no_del=False, no_dup=False, window_size=300,
                  window_offset=None):
    """Run the C oracle on packed reads; returns (pos[int64], fields[int32 n x nf], ref[u1])."""
    L = lib()
    p = _Params()
    p.stat = PILEUP_TYPES.index(stat)
    if stepper not in STEPPERS:
        raise ValueError('unknown stepper: %r' % stepper)   # pysam raises ValueError (A.0)
    p.stepper = STEPPERS[stepper]
    p.max_depth = int(max_depth)
    p.truncate = int(bool(truncate))
    p.pad = int(bool(pad))
    p.min_mapq, p.min_baseq = int(min_mapq), int(min_baseq)
    p.no_del, p.no_dup = int(bool(no_del)), int(bool(no_dup))
    p.window_size = int(window_size)
    p.window_offset = int(window_size // 2 if window_offset is None else window_offset)  # opt.pyx:1479-1480
    p.has_ref = int(ref is not None)
    p.start, p.end = int(start), int(end)
    refa = np.frombuffer(ref, dtype='u1') if ref is not None and len(ref) else np.zeros(1, 'u1')
    ref_len = len(ref) if ref is not None else 0
    nf = L.orc_nfields(p.stat)
    cap = 1 << 16
    rs = reads.cstruct()
    while True:
        out_pos = np.empty(cap, dtype='i8')
        out_f = np.empty((cap, nf), dtype='i4')
        out_ref = np.zeros(cap, dtype='u1')
        n = L.orc_pileup(C.byref(rs), C.byref(p), _ptr(refa), ref_len, _ptr(out_pos), _ptr(out_f),
                         _ptr(out_ref), cap)
        if n <= cap:
            return out_pos[:n], out_f[:n], out_ref[:n]
        cap = int(n)


def binned_packed(stat, reads, start, end, window_size, window_offset, min_mapq=0, no_dup=False,
                  ref=None):
    L = lib()
    sid = BINNED_TYPES.index(stat)
    nf = L.orc_nfields_binned(sid)
    refa = np.frombuffer(ref, dtype='u1') if ref is not None and len(ref) else np.zeros(1, 'u1')
    ref_len = len(ref) if ref is not None else 0
    cap = max(16, (int(end) - int(start)) // max(1, int(window_size)) + 16)
    rs = reads.cstruct()
    while True:
        out_pos = np.empty(cap, dtype='i8')
        out_f = np.empty((cap, nf), dtype='i4')
        n = L.orc_binned(C.byref(rs), sid, int(start), int(end), int(window_size), int(window_offset),
                         int(min_mapq), int(bool(no_dup)), _ptr(refa), ref_len, _ptr(out_pos),
                         _ptr(out_f), cap)
        if n <= cap:
            return out_pos[:n], out_f[:n]
        cap = int(n)


def _assemble(names, chrom_width, parts, with_ref, fields, gc_col):
    """util.py:31-73 -- packed structured array; parts = [(chrom, pos, F, ref)]."""
    dt = [('chrom', 'S%d' % chrom_width), ('pos', 'i4')]
    if with_ref:
        dt.append(('ref', 'S1'))
    for f in names:
        dt.append((f, 'u1' if f == gc_col else 'i4'))
    n = sum(len(p[1]) for p in parts)
    a = np.zeros(n, dtype=dt)
    o = 0
    for chrom, pos, F, ref in parts:
        m = len(pos)
        a['chrom'][o:o + m] = chrom.encode('ascii')
        a['pos'][o:o + m] = pos
        if with_ref:
            a['ref'][o:o + m] = ref.view('S1')
        for j, f in enumerate(names):
            a[f][o:o + m] = F[:, j]
        o += m
    if fields is not None:
        if any(f not in a.dtype.names for f in fields):
            raise ValueError('invalid fields: %r' % fields)
        if len(fields) == 1:
            return a[fields[0]].copy()
        a = a[list(fields)]
    return a.view(np.recarray)


def load_pileup(type, alignmentfile, fafile=None, chrom=None, start=None, end=None, one_based=False,
                truncate=False, stepper='all', pad=False, max_depth=8000, window_size=300,
                window_offset=None, min_mapq=0, min_baseq=0, no_del=False, no_dup=False, fields=None):
    if type in WITHREF and fafile is None:
        raise ValueError('reference sequence is required; please provide fafile argument')
    if type not in PILEUP_TYPES:
        raise ValueError('unsupported statistics type: %r' % type)
    bam, fa = _open(alignmentfile, fafile)
    chroms = [chrom] if chrom is not None else list(bam.references)
    parts = []
    for c in chroms:
        # chrom=None: iter_pileup_default drops start / end (opt.pyx:2111-2112), iter_pileup_padded hands them to
        # every contig (opt.pyx:1927-1932)
        keep = chrom is not None or pad
        s, e = normalise_coords(bam, c, start if keep else None, end if keep else None, one_based)
        reads = pack_region(bam, c, s, e)
        ref = fa.contig(c) if fa is not None else None
        pos, F, refcol = pileup_packed(type, reads, s, e, ref=ref, stepper=stepper, max_depth=max_depth,
                                       truncate=truncate, pad=pad, min_mapq=min_mapq,
                                       min_baseq=min_baseq, no_del=no_del, no_dup=no_dup,
                                       window_size=window_size, window_offset=window_offset)
        if one_based:
            pos = pos + 1
        parts.append((c, pos, F, refcol))
    width = max(len(x) for x in bam.references)
    return _assemble(_fields(type), width, parts, type in WITHREF and type != 'coverage_gc', fields,
                     'gc')


def load_binned(type, alignmentfile, fafile=None, chrom=None, start=None, end=None, one_based=False,
                window_size=300, window_offset=None, min_mapq=0, no_dup=False, fields=None):
    if type not in BINNED_TYPES:
        raise ValueError('unsupported statistics type: %r' % type)
    bam, fa = _open(alignmentfile, fafile)
    if window_offset is None:
        window_offset = window_size / 2       # opt.pyx:2006-2007, truncated by the typed int arg
    window_offset = int(window_offset)
    chroms = [chrom] if chrom is not None else list(bam.references)
    parts = []
    for c in chroms:
        s, e = normalise_coords(bam, c, start, end, one_based)    # iter_binned keeps start / end per contig (opt.pyx:2009-2015)
        reads = pack_region(bam, c, s, e)
        ref = fa.contig(c) if fa is not None else None
        pos, F = binned_packed(type, reads, s, e, window_size, window_offset, min_mapq, no_dup, ref)
        if one_based:
            pos = pos + 1
        parts.append((c, pos, F, None))
    width = max(len(x) for x in bam.references)
    return _assemble(_fields_binned(type), width, parts, False, fields, 'gc')
