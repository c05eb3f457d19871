"""Host-side BAM / FASTA decode into structure-of-arrays read batches.

The reference reads alignments through pysam / htslib (``AlignmentFile.pileup`` at
``pysamstats/opt.pyx:1853``, ``IteratorRowRegion`` at ``opt.pyx:2040``, ``FastaFile.fetch``
at ``opt.pyx:1989``); neither is available to this package, so it carries its own reader:

* BGZF: a series of gzip members, each at most 64 KiB, inflated independently (``zlib``
  releases the GIL, so blocks are inflated on a small thread pool).
* BAM records are parsed with numpy gathers over the inflated buffer -- only the walk over
  the ``block_size`` chain is a python loop.
* A batch is the set of records of one contig overlapping ``[start, end)`` in file order:
  the rule htslib's region iterator applies (``pos < end and endpos > start``), which is
  what both reference call sites consume.

File objects of other libraries (e.g. ``pysam.AlignmentFile`` / ``pysam.FastaFile``) are
accepted wherever a path is: only their ``.filename`` is used.
"""
import ctypes as C
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, 'csrc', 'libpss_host.so')


class HostArrays(C.Structure):
    """struct pss_host_arrays (include/pss_host.h)"""
    FIELDS = ('tid', 'pos', 'flag', 'mapq', 'tlen', 'mtid', 'mpos', 'l_seq', 'endpos', 'cigar_off', 'cigar',
              'seq_off', 'seq4', 'qual_off', 'qual', 'names')
    _fields_ = [(k, C.c_void_p) for k in FIELDS]


_host_lib = None


def host_lib():
    """libpss_host.so (C++ decoder, include/pss_host.h) or None when it is not built; the numpy
    decoder below is the (slow) fallback for the HOST side only -- the statistics themselves have
    no CPU path."""
    global _host_lib
    if _host_lib is None:
        if os.environ.get('PSS_NUMPY_DECODER') or not os.path.exists(HOST_LIB_PATH):
            _host_lib = False
        else:
            lib = C.CDLL(HOST_LIB_PATH)
            P = C.c_void_p
            lib.pss_host_error.restype = C.c_char_p
            lib.pss_bam_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(P)]
            lib.pss_bam_open_region.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_int,
                                                C.POINTER(P)]
            lib.pss_bam_index_build.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
            lib.pss_bam_close.argtypes = [P]
            lib.pss_bam_close.restype = None
            lib.pss_bam_n_ref.argtypes = [P]
            lib.pss_bam_n_ref.restype = C.c_int32
            lib.pss_bam_ref_name.argtypes = [P, C.c_int32]
            lib.pss_bam_ref_name.restype = C.c_char_p
            lib.pss_bam_ref_len.argtypes = [P, C.c_int32]
            lib.pss_bam_ref_len.restype = C.c_int64
            lib.pss_bam_sizes.argtypes = [P] + [C.POINTER(C.c_int64)] * 4 + [C.POINTER(C.c_int32)]
            lib.pss_bam_fill.argtypes = [P, C.POINTER(HostArrays)]
            lib.pss_bam_timing.argtypes = [P] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int64)]
            lib.pss_name_prev.argtypes = [P, C.c_int64, C.c_int32, P]
            lib.pss_take_cols_i32.argtypes = [P, C.c_int64, C.c_int32, P, C.c_int64, P, C.c_int64, C.c_int]
            lib.pss_pack_rows.argtypes = [P, C.c_int64, C.c_int32, C.c_int32, C.POINTER(P), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
            lib.pss_format_rows.argtypes = [C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(P), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32), C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int,
                                            C.POINTER(P), C.POINTER(C.c_int64)]
            lib.pss_host_free.argtypes = [P]
            lib.pss_host_free.restype = None
            lib.pss_host_bandwidth.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int]
            lib.pss_host_bandwidth.restype = C.c_double
            lib.pss_bam_write.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int64,
                                          C.POINTER(HostArrays), C.c_int32, C.c_int, C.c_int]
            _host_lib = lib
    return _host_lib or None


def alloc_big(shape, dtype):
    """np.empty for large results: anonymous memory advised to use transparent huge pages, so the first
    touch by the worker threads costs one fault per 2 MiB instead of one per 4 KiB (3-4x on a cold array)."""
    import mmap
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    count = int(np.prod(shape, dtype=np.int64))
    nbytes = count * dtype.itemsize
    if nbytes < (32 << 20) or not hasattr(mmap, 'MADV_HUGEPAGE'):
        return np.empty(shape, dtype=dtype)
    mm = mmap.mmap(-1, (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20))
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)
    except OSError:
        pass
    return np.frombuffer(mm, dtype=dtype, count=count).reshape(shape)


def alloc_rows(n, dtype):
    """Address space for up to ``n`` records that are filled from the front (``opt.pileup_rows``): an anonymous
    mapping that is not charged before it is touched, so a sparse contig costs what its rows cost."""
    import mmap
    dtype = np.dtype(dtype)
    nbytes = int(n) * dtype.itemsize
    if nbytes < (32 << 20):
        return np.empty(int(n), dtype=dtype)
    flags = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | getattr(mmap, 'MAP_NORESERVE', 0)
    mm = mmap.mmap(-1, (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20), flags=flags)
    if hasattr(mmap, 'MADV_HUGEPAGE'):
        try:
            mm.madvise(mmap.MADV_HUGEPAGE)
        except OSError:
            pass
    return np.frombuffer(mm, dtype=dtype, count=int(n))


def format_rows_native(columns, delimiter, terminator):
    """Text (str) of the rows made of ``columns`` -- a list of str (constant cell), int32 / int64 arrays or S1 arrays
    -- formatted on all host threads by libpss_host; None when the library or a column type is not available."""
    lib = host_lib()
    if lib is None:
        return None
    n = None
    kinds, srcs, strides, widths, keep = [], [], [], [], []
    for col in columns:
        if isinstance(col, str):
            b = col.encode('utf-8')
            buf = C.create_string_buffer(b, len(b) + 1)
            keep.append(buf)
            kinds.append(0); srcs.append(C.addressof(buf)); strides.append(0); widths.append(len(b))
            continue
        col = np.asarray(col)
        if col.ndim != 1 or (n is not None and len(col) != n):
            return None
        n = len(col)
        if col.dtype == np.dtype('<i4'):
            kinds.append(1)
        elif col.dtype == np.dtype('<i8'):
            kinds.append(2)
        elif col.dtype == np.dtype('S1'):
            kinds.append(3)
        else:
            return None
        keep.append(col)
        srcs.append(col.ctypes.data if n else C.addressof(C.create_string_buffer(8)))
        strides.append(col.strides[0]); widths.append(0)
    if n is None:
        return None
    k = len(columns)
    out, out_len = C.c_void_p(), C.c_int64(0)
    d, t = delimiter.encode('utf-8'), terminator.encode('utf-8')
    _host_check(lib, lib.pss_format_rows(n, k, (C.c_int32 * k)(*kinds), (C.c_void_p * k)(*srcs), (C.c_int32 * k)(*strides),
                                         (C.c_int32 * k)(*widths), d, len(d), t, len(t), 0, C.byref(out), C.byref(out_len)))
    try:
        # decoded straight from the library's buffer (no intermediate bytes object)
        return str((C.c_char * out_len.value).from_address(out.value), 'utf-8') if out_len.value else ''
    finally:
        lib.pss_host_free(out)


def take_cols(out, sel):
    """out[:, sel] for a C-contiguous int32 [n_cols][n] array, threaded (libpss_host) when it is large."""
    lib = host_lib()
    if lib is None or out.dtype != np.int32 or not out.flags.c_contiguous or out.size < (1 << 22):
        return np.ascontiguousarray(out[:, sel])
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    dst = alloc_big((out.shape[0], len(sel)), np.int32)
    _host_check(lib, lib.pss_take_cols_i32(out.ctypes.data, out.shape[1], out.shape[0], sel.ctypes.data, len(sel),
                                           dst.ctypes.data, len(sel), 0))
    return dst


def _host_check(lib, rc):
    if rc != 0:
        raise ValueError(lib.pss_host_error().decode('ascii', 'replace'))

_REF_CONSUMING = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=bool)  # MIDNSHP=X


def _path_of(f):
    if isinstance(f, (BamFile, FastaFile)):
        return f.filename
    if isinstance(f, bytes):
        return f.decode()
    if isinstance(f, (str, os.PathLike)):
        return os.fspath(f)
    name = getattr(f, 'filename', None)
    if name is None:
        raise TypeError('expected a path or an object with a .filename, got %r' % type(f))
    return name.decode() if isinstance(name, bytes) else str(name)


# ------------------------------------------------------------------------------ BGZF

def _bgzf_blocks(data):
    """Yield (payload_start, payload_end) of every BGZF block's raw deflate stream."""
    off, n = 0, len(data)
    while off < n:
        if data[off:off + 4] != b'\x1f\x8b\x08\x04':
            raise ValueError('not a BGZF file (bad block header at %d)' % off)
        xlen, = struct.unpack_from('<H', data, off + 10)
        x, xend, bsize = off + 12, off + 12 + xlen, None
        while x < xend:
            si1, si2, slen = data[x], data[x + 1], struct.unpack_from('<H', data, x + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize, = struct.unpack_from('<H', data, x + 4)
            x += 4 + slen
        if bsize is None:
            raise ValueError('BGZF block without BC subfield at %d' % off)
        yield xend, off + bsize + 1 - 8
        off += bsize + 1


def bgzf_inflate(path, threads=4):
    with open(path, 'rb') as fh:
        data = fh.read()
    spans = list(_bgzf_blocks(data))

    def inflate(span):
        return zlib.decompress(data[span[0]:span[1]], -15)

    if len(spans) > 8 and threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(inflate, spans, chunksize=32))
    else:
        parts = [inflate(s) for s in spans]
    return b''.join(parts)


# ---------------------------------------------------------------------- read batches

def _ragged_gather(buf, starts, lens):
    """Concatenate buf[starts[i]:starts[i]+lens[i]] for all i; returns (data, offsets)."""
    lens = lens.astype(np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    if total == 0:
        return np.zeros(0, dtype=buf.dtype), off
    idx = np.repeat(starts.astype(np.int64) - off[:-1], lens) + np.arange(total, dtype=np.int64)
    return buf[idx], off


class ReadBatch(object):
    """SoA arrays of the records of one contig, file order (see include/pss_b200.h
    ``struct pss_reads``)."""

    __slots__ = ('tid', 'pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'cigar', 'seq_off',
                 'seq4', 'qual_off', 'qual', '_name_prev', 'endpos', 'names', '_cache')

    @property
    def cache(self):
        """Derived facts about the batch that several calls share (``opt._position_chunks``)."""
        if getattr(self, '_cache', None) is None:
            self._cache = {}
        return self._cache

    @property
    def n_reads(self):
        return len(self.pos)

    @property
    def name_prev(self):
        """Index of the previous record of the same QNAME (mate-overlap rewrite); hashed on first use -- only
        requests that read base qualities need it."""
        if getattr(self, '_name_prev', None) is None:
            self._name_prev = _name_prev(self.names)
        return self._name_prev

    @name_prev.setter
    def name_prev(self, v):
        self._name_prev = v

    def subset(self, idx):
        """Records idx (increasing) as a new contiguous batch."""
        b = ReadBatch()
        b.tid = self.tid
        for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'endpos', 'names'):
            setattr(b, k, np.ascontiguousarray(getattr(self, k)[idx]))
        for data, off in (('cigar', 'cigar_off'), ('seq4', 'seq_off'), ('qual', 'qual_off')):
            o = getattr(self, off)
            d, no = _ragged_gather(getattr(self, data), o[idx], o[idx + 1] - o[idx])    # (no pass over the whole batch)
            setattr(b, data, d)
            setattr(b, off, no)
        b.name_prev = None
        return b

    def padded(self):
        """Arrays safe to hand to the engine (no zero-length buffers)."""
        for k in ('cigar', 'seq4', 'qual'):
            if len(getattr(self, k)) == 0:
                setattr(self, k, np.zeros(1, dtype=getattr(self, k).dtype))
        return self


def _name_prev(names):
    """Index of the previous record with the same QNAME (or -1), file order."""
    n = len(names)
    prev = np.full(n, -1, dtype=np.int32)
    lib = host_lib()
    if lib is not None and n:
        names = np.ascontiguousarray(names)
        _host_check(lib, lib.pss_name_prev(names.ctypes.data, n, names.dtype.itemsize, prev.ctypes.data))
        return prev
    if n:
        _, inv = np.unique(names, return_inverse=True)
        order = np.argsort(inv, kind='stable')
        same = inv[order][1:] == inv[order][:-1]
        prev[order[1:][same]] = order[:-1][same]
    return prev


class BamFile(object):
    """A BAM file.  With the native decoder the file is read lazily: the header at once, the
    records either whole (first use of ``contig_batch`` / a region query without an index) or,
    when a ``.bai`` index sits next to a file of at least ``PSS_BAI_MIN_BYTES`` (default 32 MiB),
    only the BGZF blocks the index lists for the region asked (``pss_bam_open_region``)."""

    def __init__(self, path, threads=0):
        self.filename = _path_of(path)
        self.decode_seconds = None
        self._threads = int(threads)
        self._all = None
        self._tid = None
        lib = host_lib()
        if lib is not None:
            h = C.c_void_p()
            _host_check(lib, lib.pss_bam_open_region(self.filename.encode(), None, -1, 0, 0, 1, C.byref(h)))
            try:
                self._read_header(lib, h)
            finally:
                lib.pss_bam_close(h)
        else:
            self._init_numpy()
        self._tid_of = {r: i for i, r in enumerate(self.references)}
        self._contig_cache = {}
        self._piece_cache = {}

    def _read_header(self, lib, h):
        nref = lib.pss_bam_n_ref(h)
        self.references = tuple(lib.pss_bam_ref_name(h, i).decode('ascii') for i in range(nref))
        self.lengths = tuple(int(lib.pss_bam_ref_len(h, i)) for i in range(nref))

    def index_path(self):
        """Path of the BAI index next to the file (``x.bam.bai`` or ``x.bai``) or None."""
        for cand in (self.filename + '.bai', os.path.splitext(self.filename)[0] + '.bai'):
            if os.path.exists(cand):
                return cand
        return None

    def _decode(self, lib, h, light=False):
        """SoA arrays of the records behind an open native handle -> (ReadBatch, tid array).  ``light``: without base
        qualities and read names (150 + ~10 of the ~300 bytes of a record that requests which do not read qualities never
        look at: not allocated, not copied)."""
        n, nc, ns, nq, w = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        _host_check(lib, lib.pss_bam_sizes(h, C.byref(n), C.byref(nc), C.byref(ns), C.byref(nq), C.byref(w)))
        n, nc, ns, nq, w = n.value, nc.value, ns.value, nq.value, w.value
        A = ReadBatch()
        A.tid = -2
        tid = np.empty(n, 'i4')
        A.pos, A.flag, A.mapq = np.empty(n, 'i4'), np.empty(n, 'u2'), np.empty(n, 'u1')
        A.tlen, A.mtid, A.l_seq = np.empty(n, 'i4'), np.empty(n, 'i4'), np.empty(n, 'i4')
        A.endpos = np.empty(n, 'i8')
        A.cigar_off, A.cigar = np.empty(n + 1, 'i8'), np.empty(max(nc, 1), 'u4')
        A.seq_off, A.seq4 = np.empty(n + 1, 'i8'), np.empty(max(ns, 1), 'u1')
        if light:
            A.qual_off, A.qual, names = None, np.zeros(1, 'u1'), None
        else:
            A.qual_off, A.qual = np.empty(n + 1, 'i8'), np.empty(max(nq, 1), 'u1')
            names = np.zeros(n, dtype='S%d' % w)
        ha = HostArrays()
        ha.tid = tid.ctypes.data
        for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'endpos', 'cigar_off', 'cigar', 'seq_off',
                  'seq4') + (() if light else ('qual_off', 'qual')):
            setattr(ha, k, getattr(A, k).ctypes.data)
        if not light:
            ha.names = names.ctypes.data
        _host_check(lib, lib.pss_bam_fill(h, C.byref(ha)))
        A.cigar, A.seq4 = A.cigar[:nc], A.seq4[:ns]
        if not light:
            A.qual = A.qual[:nq]
        A.names = names
        A.name_prev = None
        tr, ti, tx, rb = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        lib.pss_bam_timing(h, C.byref(tr), C.byref(ti), C.byref(tx), C.byref(rb))
        self.decode_seconds = dict(read=tr.value, inflate=ti.value, index=tx.value, raw_bytes=rb.value)
        return A, tid

    def _load_all(self):
        """Whole-file decode (C++ decoder, csrc/pss_host.cpp: threaded BGZF inflate + parallel SoA fill)."""
        if self._all is None:
            lib = host_lib()
            h = C.c_void_p()
            _host_check(lib, lib.pss_bam_open(self.filename.encode(), self._threads, C.byref(h)))
            try:
                self._all, self._tid = self._decode(lib, h)
            finally:
                lib.pss_bam_close(h)
        return self._all

    def _region_native(self, tid, start, end, bai, light=False):
        lib = host_lib()
        h = C.c_void_p()
        rc = lib.pss_bam_open_region(self.filename.encode(), bai.encode(), int(tid), int(start), int(end),
                                     self._threads, C.byref(h))
        if rc == -6:
            # the index does not belong to this file (stale / truncated .bai): decode the whole file instead
            import warnings
            warnings.warn('ignoring BAM index %s: %s' % (bai, lib.pss_host_error().decode('ascii', 'replace')))
            return None
        _host_check(lib, rc)
        try:
            b, _ = self._decode(lib, h, light)
        finally:
            lib.pss_bam_close(h)
        b.tid = tid
        b.name_prev = None                      # hashed on first use (only requests that read base qualities need it)
        if b.n_reads > 1 and np.any(np.diff(b.pos) < 0):
            raise ValueError('BAM file is not coordinate sorted: %r' % self.filename)
        return b

    def _init_numpy(self):
        raw = bgzf_inflate(self.filename)
        if raw[:4] != b'BAM\x01':
            raise ValueError('not a BAM file: %r' % self.filename)
        l_text, = struct.unpack_from('<i', raw, 4)
        off = 8 + l_text
        n_ref, = struct.unpack_from('<i', raw, off)
        off += 4
        refs, lens = [], []
        for _ in range(n_ref):
            l_name, = struct.unpack_from('<i', raw, off)
            refs.append(raw[off + 4:off + 4 + l_name - 1].decode('ascii'))
            lens.append(struct.unpack_from('<i', raw, off + 4 + l_name)[0])
            off += 8 + l_name
        self.references = tuple(refs)
        self.lengths = tuple(lens)
        # walk the block_size chain
        starts = []
        n = len(raw)
        unpack = struct.Struct('<i').unpack_from
        while off < n:
            starts.append(off + 4)
            off += 4 + unpack(raw, off)[0]
        buf = np.frombuffer(raw, dtype=np.uint8)
        st = np.array(starts, dtype=np.int64)
        fixed = buf[st[:, None] + np.arange(32)] if len(st) else np.zeros((0, 32), np.uint8)
        fixed = np.ascontiguousarray(fixed)

        def col(a, b, dt):
            return np.ascontiguousarray(fixed[:, a:b]).view(dt).ravel()

        self._tid = col(0, 4, '<i4')
        A = ReadBatch()
        A.tid = -2
        A.pos = col(4, 8, '<i4')
        l_name = fixed[:, 8].astype(np.int64)
        A.mapq = np.ascontiguousarray(fixed[:, 9])
        n_cig = col(12, 14, '<u2').astype(np.int64)
        A.flag = col(14, 16, '<u2')
        A.l_seq = col(16, 20, '<i4')
        A.mtid = col(20, 24, '<i4')
        A.tlen = col(28, 32, '<i4')
        lseq = A.l_seq.astype(np.int64)
        name_start = st + 32
        cig_start = name_start + l_name
        seq_start = cig_start + 4 * n_cig
        qual_start = seq_start + (lseq + 1) // 2
        cig_bytes, coff = _ragged_gather(buf, cig_start, 4 * n_cig)
        A.cigar = np.ascontiguousarray(cig_bytes).view('<u4')
        A.cigar_off = coff // 4
        A.seq4, A.seq_off = _ragged_gather(buf, seq_start, (lseq + 1) // 2)
        A.qual, A.qual_off = _ragged_gather(buf, qual_start, lseq)
        # QNAMEs as a fixed-width bytes column (NUL padded)
        width = int(l_name.max()) if len(l_name) else 1
        nm = np.zeros((len(st), width), dtype=np.uint8)
        nb, noff = _ragged_gather(buf, name_start, l_name)
        if len(nb):
            rows = np.repeat(np.arange(len(st)), l_name)
            cols = np.arange(len(nb)) - np.repeat(noff[:-1], l_name)
            nm[rows, cols] = nb
        A.names = nm.view('S%d' % width).ravel()
        # htslib bam_endpos(): reference length of the CIGAR, 1 if unmapped or zero
        op = (A.cigar & 15).astype(np.int64)
        ln = (A.cigar >> 4).astype(np.int64)
        reflen = np.zeros(len(st), dtype=np.int64)
        if len(op):
            rid = np.repeat(np.arange(len(st)), n_cig)
            np.add.at(reflen, rid, np.where(_REF_CONSUMING[op], ln, 0))
        reflen[(A.flag & 4) != 0] = 0
        A.endpos = A.pos.astype(np.int64) + np.where(reflen > 0, reflen, 1)
        A.name_prev = None
        self._all = A

    def _load_all_any(self):
        return self._all if self._all is not None else self._load_all()

    def load(self):
        """Decode the whole file now (it is otherwise decoded on first use)."""
        self._load_all_any()
        return self

    def get_tid(self, chrom):
        return self._tid_of[chrom]

    def contig_batch(self, chrom):
        """All records placed on `chrom`, file order."""
        if chrom not in self._contig_cache:
            self._load_all_any()
            tid = self._tid_of[chrom]
            idx = np.flatnonzero(self._tid == tid)
            if len(idx) == len(self._tid):        # single-contig file: no gather needed
                b = self._all
            else:
                b = self._all.subset(idx)
            b.tid = tid
            if len(idx) > 1 and np.any(np.diff(b.pos) < 0):
                raise ValueError('BAM file is not coordinate sorted: %r' % self.filename)
            self._contig_cache[chrom] = b
        return self._contig_cache[chrom]

    def can_stream(self, chrom):
        """True when the records of ``chrom`` can be decoded piece by piece through the BAI index (native decoder, an
        index next to a file of at least PSS_BAI_MIN_BYTES) and are not already decoded in this process."""
        if os.environ.get('PSS_NO_STREAM') or self._all is not None or chrom in self._contig_cache or host_lib() is None:
            return False
        return self.index_path() is not None and \
            os.path.getsize(self.filename) >= int(os.environ.get('PSS_BAI_MIN_BYTES', 32 << 20))

    def piece_batches(self, chrom, cuts, ahead=2):
        """The records overlapping each of the consecutive position pieces ``[cuts[k], cuts[k + 1])`` of ``chrom``, as a
        generator of ReadBatch -- the host half of the decode -> upload -> compute pipeline (SURVEY.md section 8f N3;
        the reference streams reads as htslib inflates them, ``opt.pyx:1853-1854``, ``:2038-2048``): piece k + 1 and
        k + 2 are inflated and parsed through the BAI index (``pss_bam_open_region``, all host threads) on a background
        threads while the caller uploads / computes / packs piece k.  The decoded pieces are kept (``_piece_cache``), so a
        second statistic over the same pieces does not decode again.  Yields None (once, and stops) if the index turns
        out to be unusable."""
        key = (chrom, tuple(int(c) for c in cuts))
        cached = self._piece_cache.get(key)
        if cached is not None:
            for b in cached:
                yield b
            return
        tid, bai = self._tid_of[chrom], self.index_path()
        n = len(cuts) - 1
        done = []
        # (three decodes at a time: the serial parts of one -- index lookup, record walk -- run beside the threaded
        # inflate / fill of the others)
        with ThreadPoolExecutor(1 + ahead) as ex:
            futs = {}

            def want(k):
                if k < n and k not in futs:
                    futs[k] = ex.submit(self._region_native, tid, cuts[k], cuts[k + 1], bai, True)   # (streamed requests never read qualities)
            for k in range(min(n, 1 + ahead)):
                want(k)
            for k in range(n):
                b = futs.pop(k).result()
                want(k + 1 + ahead)
                if b is None:
                    for f in futs.values():
                        f.cancel()
                    yield None
                    return
                done.append(b)
                if k == n - 1:           # (before the last yield: the consumer may not come back for a StopIteration)
                    if len(self._piece_cache) >= 2:
                        self._piece_cache.pop(next(iter(self._piece_cache)))
                    self._piece_cache[key] = done
                yield b

    def region_batch(self, chrom, start, end):
        """Records the region iterator yields for chrom:[start, end)."""
        if self._all is None and chrom not in self._contig_cache and host_lib() is not None:
            bai = self.index_path()
            full = start <= 0 and end >= self.lengths[self._tid_of[chrom]] and len(self.references) == 1
            if bai is not None and not full and \
                    os.path.getsize(self.filename) >= int(os.environ.get('PSS_BAI_MIN_BYTES', 32 << 20)):
                b = self._region_native(self._tid_of[chrom], start, end, bai)
                if b is not None:
                    return b
        b = self.contig_batch(chrom)
        n = b.n_reads
        if n == 0:
            return b
        # the records are start sorted: only those starting in [start - longest span, end) can overlap the region
        if 'maxspan' not in b.cache:
            b.cache['maxspan'] = int((b.endpos - b.pos).max())
        i0 = int(np.searchsorted(b.pos, start - b.cache['maxspan'], side='left'))
        i1 = int(np.searchsorted(b.pos, end, side='left'))
        idx = i0 + np.flatnonzero(b.endpos[i0:i1] > start)
        if len(idx) == n:
            return b
        sub = b.subset(idx)
        sub.tid = b.tid
        return sub


def write_bam(path, references, lengths, batch, names=None, level=1, threads=0):
    """Write the records of ONE contig batch (numpy SoA arrays as in ``struct pss_reads`` plus
    ``endpos``; ``tid`` scalar) as a coordinate-sorted BAM.  ``names``: fixed-width bytes array,
    default ``r<k>`` with k shared by a record and its ``name_prev`` chain.  Needs
    libpss_host.so."""
    lib = host_lib()
    if lib is None:
        raise RuntimeError('libpss_host.so is not built (make -C pysamstats_b200/csrc)')
    n = int(batch['n_reads'])
    if names is None:
        prev = np.asarray(batch['name_prev'])
        ids = np.arange(n, dtype=np.int64)
        has = prev >= 0
        # chains are short (mates): resolve to the first record of the chain
        for _ in range(64):
            nxt = np.where(has, ids[np.where(has, prev, 0)], ids)
            if np.array_equal(nxt, ids):
                break
            ids = nxt
        names = np.char.add(b'r', ids.astype('S20')).astype('S21')
    names = np.ascontiguousarray(names)
    keep = []

    def arr(k, dt, size=None):
        a = np.ascontiguousarray(np.asarray(batch[k]).astype(dt, copy=False))
        keep.append(a)
        return a.ctypes.data

    ha = HostArrays()
    tid = np.full(n, int(batch.get('tid', 0)), 'i4')
    mpos = np.full(n, -1, 'i4')
    keep += [tid, mpos, names]
    ha.tid, ha.mpos, ha.names = tid.ctypes.data, mpos.ctypes.data, names.ctypes.data
    for k, dt in (('pos', 'i4'), ('flag', 'u2'), ('mapq', 'u1'), ('tlen', 'i4'), ('mtid', 'i4'), ('l_seq', 'i4'),
                  ('endpos', 'i8'), ('cigar_off', 'i8'), ('cigar', 'u4'), ('seq_off', 'i8'), ('seq4', 'u1')):
        setattr(ha, k, arr(k, dt))
    if batch.get('qual') is not None:
        ha.qual_off, ha.qual = arr('qual_off', 'i8'), arr('qual', 'u1')
    rn = (C.c_char_p * len(references))(*[r.encode() for r in references])
    rl = (C.c_int64 * len(references))(*[int(x) for x in lengths])
    _host_check(lib, lib.pss_bam_write(os.fspath(path).encode(), len(references), rn, rl, n, C.byref(ha),
                                       names.dtype.itemsize, int(level), int(threads)))


_bam_cache = {}


def index_bam(path, index_path=None, threads=0):
    """Write the BAI index of a coordinate-sorted BAM (what ``samtools index`` / ``pysam.index`` do for the
    reference's users); returns the index path."""
    lib = host_lib()
    if lib is None:
        raise RuntimeError('libpss_host.so is not built')
    path = _path_of(path)
    out = index_path or path + '.bai'
    _host_check(lib, lib.pss_bam_index_build(path.encode(), out.encode(), int(threads)))
    return out


def open_bam(f):
    """BamFile for a path / BamFile / foreign object with .filename (small in-process cache)."""
    if isinstance(f, BamFile):
        return f
    path = _path_of(f)
    key = (os.path.abspath(path), os.path.getmtime(path), os.path.getsize(path))
    bam = _bam_cache.get(key)
    if bam is None:
        if len(_bam_cache) >= 4:
            _bam_cache.pop(next(iter(_bam_cache)))
        bam = _bam_cache[key] = BamFile(path)
    return bam


# ------------------------------------------------------------------------------ FASTA

class FastaFile(object):
    """Indexed FASTA access: uses ``<path>.fai`` when present, else scans the file once."""

    def __init__(self, path):
        self.filename = _path_of(path)
        self._index = {}
        self.references = []
        fai = self.filename + '.fai'
        if os.path.exists(fai):
            with open(fai) as fh:
                for line in fh:
                    f = line.rstrip('\n').split('\t')
                    if len(f) >= 5:
                        self._index[f[0]] = tuple(int(x) for x in f[1:5])
                        self.references.append(f[0])
        else:
            self._scan()
        self._cache = (None, None)

    def _scan(self):
        with open(self.filename, 'rb') as fh:
            off = 0
            name, length, seq_off, lb, lw = None, 0, 0, 0, 0
            for line in fh:
                if line.startswith(b'>'):
                    if name is not None:
                        self._index[name] = (length, seq_off, lb, lw)
                    name = line[1:].split()[0].decode('ascii')
                    self.references.append(name)
                    length, seq_off, lb, lw = 0, off + len(line), 0, 0
                else:
                    s = line.rstrip(b'\r\n')
                    if lb == 0:
                        lb, lw = len(s), len(line)
                    length += len(s)
                off += len(line)
            if name is not None:
                self._index[name] = (length, seq_off, lb, lw)

    def get_reference_length(self, chrom):
        return self._index[chrom][0] if chrom in self._index else 0

    def contig(self, chrom):
        """Whole contig as a uint8 array, case preserved.  An unknown contig raises KeyError, as
        ``pysam.FastaFile.fetch`` does where the reference asks for its bases (``opt.pyx:1989``)."""
        if self._cache[0] == chrom:
            return self._cache[1]
        if chrom not in self._index:
            raise KeyError("sequence '%s' not present" % (chrom,))
        else:
            length, off, lb, lw = self._index[chrom]
            nlines = (length + lb - 1) // lb if lb else 0
            with open(self.filename, 'rb') as fh:
                fh.seek(off)
                raw = fh.read(length + nlines * (lw - lb) if lb else 0)
            a = np.frombuffer(raw, dtype=np.uint8)
            a = a[(a != 10) & (a != 13)][:length].copy()
        self._cache = (chrom, a)
        return a

    def fetch(self, chrom, start, end):
        """bytes of chrom[start:end], clipped to the contig like faidx."""
        a = self.contig(chrom)
        start = max(0, int(start))
        end = min(len(a), int(end))
        return a[start:end].tobytes() if end > start else b''


_fa_cache = {}


def open_fasta(f):
    if f is None or isinstance(f, FastaFile):
        return f
    path = _path_of(f)
    key = (os.path.abspath(path), os.path.getmtime(path))
    fa = _fa_cache.get(key)
    if fa is None:
        if len(_fa_cache) >= 4:
            _fa_cache.pop(next(iter(_fa_cache)))
        fa = _fa_cache[key] = FastaFile(path)
    return fa
