"""ctypes binding of ``libpss_b200.so`` (the C ABI declared in ``include/pss_b200.h``).

The shared library is built in-tree by ``pysamstats_b200/csrc/Makefile`` (see
``__graft_entry__.build``).  There is NO CPU fallback: if the library is missing or no
CUDA device is usable, importing the engine fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PSS_LIB') or os.path.join(HERE, 'csrc', 'libpss_b200.so')   # PSS_LIB: kernel-variant experiments

PSS_OK, PSS_EINVAL, PSS_ECUDA, PSS_ENOMEM, PSS_EUNSORTED, PSS_ETOOLONG = 0, -1, -2, -3, -4, -5
PSS_FLAG_GENERIC_KERNEL = 1
PSS_FLAG_LONG_READS = 2
KTAG_GC, KTAG_PREP, KTAG_BINNED_ACC, KTAG_BINNED_FIN, KTAG_BINNED_GC = 100, 101, 110, 111, 112

PILEUP_STATS = ('coverage', 'coverage_strand', 'coverage_ext', 'coverage_ext_strand', 'variation',
                'variation_strand', 'tlen', 'tlen_strand', 'mapq', 'mapq_strand', 'baseq',
                'baseq_strand', 'baseq_ext', 'baseq_ext_strand', 'coverage_gc')
BINNED_STATS = ('coverage', 'coverage_ext', 'mapq', 'alignment', 'tlen')
STEPPERS = {'all': 0, 'nofilter': 1, 'samtools': 2}


class PssReads(C.Structure):
    """struct pss_reads"""
    _fields_ = [('n_reads', C.c_int64), ('n_cigar_words', C.c_int64), ('n_seq_bytes', C.c_int64),
                ('n_qual_bytes', C.c_int64), ('tid', C.c_int32), ('reserved', C.c_int32),
                ('pos', C.c_void_p), ('flag', C.c_void_p), ('mapq', C.c_void_p), ('tlen', C.c_void_p),
                ('mtid', C.c_void_p), ('l_seq', C.c_void_p), ('cigar_off', C.c_void_p),
                ('cigar', C.c_void_p), ('seq_off', C.c_void_p), ('seq4', C.c_void_p),
                ('qual_off', C.c_void_p), ('qual', C.c_void_p), ('name_prev', C.c_void_p)]


class PssPileupParams(C.Structure):
    """struct pss_pileup_params"""
    _fields_ = [('stat', C.c_int32), ('stepper', C.c_int32), ('max_depth', C.c_int32),
                ('min_mapq', C.c_int32), ('min_baseq', C.c_int32), ('no_del', C.c_int32),
                ('no_dup', C.c_int32), ('window_size', C.c_int32), ('window_offset', C.c_int32),
                ('flags', C.c_int32), ('start', C.c_int64), ('end', C.c_int64),
                ('ref', C.c_void_p), ('ref_start', C.c_int64), ('ref_len', C.c_int64),
                ('contig_len', C.c_int64)]


class PssBinnedParams(C.Structure):
    """struct pss_binned_params"""
    _fields_ = [('stat', C.c_int32), ('window_size', C.c_int32), ('min_mapq', C.c_int32),
                ('no_dup', C.c_int32), ('start', C.c_int64), ('end', C.c_int64), ('n_bins', C.c_int64),
                ('ref', C.c_void_p), ('ref_start', C.c_int64), ('ref_len', C.c_int64)]


class PssRowLayout(C.Structure):
    """struct pss_row_layout"""
    _fields_ = [('row_bytes', C.c_int32), ('n_fields', C.c_int32), ('field_off', C.c_void_p),
                ('field_bytes', C.c_void_p), ('chrom', C.c_void_p), ('chrom_off', C.c_int32),
                ('chrom_bytes', C.c_int32), ('pos_off', C.c_int32), ('ref_off', C.c_int32),
                ('check_field', C.c_int32), ('reserved', C.c_int32), ('pos_add', C.c_int64),
                ('pad_start', C.c_int64), ('pad_end', C.c_int64)]


# every symbol include/pss_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    'pss_create': (C.c_int, [C.c_int, C.POINTER(_P)]),
    'pss_destroy': (None, [_P]),
    'pss_last_error': (C.c_char_p, [_P]),
    'pss_version': (C.c_char_p, []),
    'pss_synchronize': (C.c_int, [_P, _P]),
    'pss_launch_count': (C.c_int64, [_P]),
    'pss_pileup_kernel_ms': (C.c_double, [_P, C.c_int32, C.POINTER(C.c_int64)]),
    'pss_transfer_bytes': (None, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'pss_set_kernel_timing': (None, [_P, C.c_int]),
    'pss_pileup_nfields': (C.c_int32, [C.c_int32]),
    'pss_binned_nfields': (C.c_int32, [C.c_int32]),
    'pss_pileup_needs_ref': (C.c_int32, [C.c_int32]),
    'pss_pileup_needs_qual': (C.c_int32, [C.c_int32, C.c_int32]),
    'pss_tile_positions': (C.c_int32, []),
    'pss_binned_nbins': (C.c_int64, [C.c_int64, C.c_int64, C.c_int32]),
    'pss_pileup': (C.c_int, [_P, C.POINTER(PssReads), C.POINTER(PssPileupParams), _P, _P, _P]),
    'pss_prepare': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    'pss_pileup_prepared': (C.c_int, [_P, C.POINTER(PssPileupParams), _P, _P, _P]),
    'pss_pileup_host': (C.c_int, [_P, C.POINTER(PssReads), C.POINTER(PssPileupParams), _P, _P]),
    'pss_pileup_host_multi': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.POINTER(PssPileupParams),
                                        C.POINTER(_P), C.POINTER(_P)]),
    'pss_pileup_host_rows': (C.c_int, [_P, C.POINTER(PssReads), C.POINTER(PssPileupParams), C.POINTER(PssRowLayout), _P,
                                       C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'pss_pileup_host_stream': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.POINTER(PssPileupParams),
                                         C.POINTER(_P), C.POINTER(_P), C.c_int64, C.c_int64]),
    'pss_admit': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.c_int32, _P, _P]),
    'pss_overlap_tweak': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.c_int32, _P, _P, _P]),
    'pss_pairs_host': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.c_int32, _P]),
    'pss_binned': (C.c_int, [_P, C.POINTER(PssReads), C.POINTER(PssBinnedParams), _P, _P]),
    'pss_binned_host': (C.c_int, [_P, C.POINTER(PssReads), C.POINTER(PssBinnedParams), _P]),
    'pss_binned_multi': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.POINTER(PssBinnedParams), C.POINTER(_P), _P]),
    'pss_binned_host_multi': (C.c_int, [_P, C.POINTER(PssReads), C.c_int32, C.POINTER(PssBinnedParams), C.POINTER(_P)]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load_library():
    """dlopen libpss_b200.so and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError('%s is missing: build it with `make -C pysamstats_b200/csrc` '
                              '(or __graft_entry__.build()); there is no CPU fallback' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def _addr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, 'data_ptr'):
        return a.data_ptr()
    raise TypeError('cannot take the address of %r' % type(a))


READ_FIELDS = ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'cigar', 'seq_off', 'seq4',
               'qual_off', 'qual', 'name_prev')
READ_DTYPES = {'pos': 'i4', 'flag': 'u2', 'mapq': 'u1', 'tlen': 'i4', 'mtid': 'i4', 'l_seq': 'i4',
               'cigar_off': 'i8', 'cigar': 'u4', 'seq_off': 'i8', 'seq4': 'u1', 'qual_off': 'i8',
               'qual': 'u1', 'name_prev': 'i4'}


def make_reads_struct(batch, n_reads=None):
    """Build a struct pss_reads from an object/dict holding the SoA arrays (numpy arrays for
    the *_host entry points, torch CUDA tensors or raw addresses for the device ones).
    The sizes come from ``batch['n_reads']`` etc. when given, else from the array lengths."""
    get = batch.get if isinstance(batch, dict) else (lambda k, d=None: getattr(batch, k, d))
    r = PssReads()

    def length(x):
        return int(x.shape[0]) if hasattr(x, 'shape') else None

    n = get('n_reads') if n_reads is None else n_reads
    if n is None:
        n = length(get('pos'))
    r.n_reads = int(n)
    for key, arr in (('n_cigar_words', 'cigar'), ('n_seq_bytes', 'seq4'), ('n_qual_bytes', 'qual')):
        v = get(key)
        if v is None:
            a = get(arr)
            v = length(a) if a is not None else 0
        setattr(r, key, int(v or 0))
    r.tid = int(get('tid', 0))
    for f in READ_FIELDS:
        setattr(r, f, _addr(get(f)))
    return r


class Context(object):
    """One pss_ctx (one GPU)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _P()
        rc = self.lib.pss_create(int(device), C.byref(h))
        if rc != PSS_OK:
            raise EngineError('pss_create(device=%d) failed (%d): no usable CUDA device; this '
                              'engine has no CPU fallback' % (device, rc))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.pss_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc == PSS_OK:
            return
        msg = self.lib.pss_last_error(self.h).decode('ascii', 'replace')
        if rc in (PSS_EINVAL, PSS_EUNSORTED, PSS_ETOOLONG):
            raise ValueError(msg)
        if rc == PSS_ENOMEM:
            raise MemoryError(msg)
        raise EngineError(msg)

    # -- thin wrappers -------------------------------------------------------------
    def pileup_params(self, stat, stepper='all', max_depth=8000, min_mapq=0, min_baseq=0, no_del=False,
                      no_dup=False, window_size=300, window_offset=None, start=0, end=0, ref=None,
                      ref_start=0, ref_len=0, contig_len=0, flags=0):
        p = PssPileupParams()
        p.stat = PILEUP_STATS.index(stat) if isinstance(stat, str) else int(stat)
        if isinstance(stepper, str):
            if stepper not in STEPPERS:
                raise ValueError('unknown stepper: %r' % stepper)
            stepper = STEPPERS[stepper]
        p.stepper = stepper
        p.max_depth = int(max_depth or 0)
        p.min_mapq, p.min_baseq = int(min_mapq), int(min_baseq)
        p.no_del, p.no_dup = int(bool(no_del)), int(bool(no_dup))
        p.window_size = int(window_size)
        p.window_offset = int(window_size // 2 if window_offset is None else window_offset)
        p.flags = int(flags)
        p.start, p.end = int(start), int(end)
        p.ref = _addr(ref)
        p.ref_start, p.ref_len, p.contig_len = int(ref_start), int(ref_len), int(contig_len)
        return p

    def pileup_host(self, reads_struct, params, out, rowmask):
        self.check(self.lib.pss_pileup_host(self.h, C.byref(reads_struct), C.byref(params), _addr(out),
                                            _addr(rowmask)))

    def pileup_host_multi(self, reads_struct, params_list, outs, rowmasks):
        n = len(params_list)
        pa = (PssPileupParams * n)(*params_list)
        oa = (_P * n)(*[_addr(o) for o in outs])
        ma = (_P * n)(*[_addr(m) for m in rowmasks])
        self.check(self.lib.pss_pileup_host_multi(self.h, C.byref(reads_struct), n, pa, oa, ma))

    def row_layout(self, stat, dtype, value_fields, chrom=None, pos_add=0, pad=None, check_field=None):
        """struct pss_row_layout for records of ``dtype`` (a numpy structured dtype): the value fields of the
        statistic that ``dtype`` names (int32, or one byte), plus ``chrom`` / ``pos`` / ``ref`` when present.
        None when a field's type is not one the row sink writes."""
        names = dtype.names or ()
        nf = len(value_fields)
        off = np.full(nf, -1, dtype=np.int32)
        nby = np.full(nf, 4, dtype=np.int32)
        for j, f in enumerate(value_fields):
            if f in names:
                ft = dtype.fields[f][0]
                if ft == np.dtype('<i4'):
                    nby[j] = 4
                elif ft == np.dtype('u1'):
                    nby[j] = 1
                else:
                    return None
                off[j] = dtype.fields[f][1]
        L = PssRowLayout()
        L.row_bytes, L.n_fields = dtype.itemsize, nf
        L.chrom_off = L.pos_off = L.ref_off = L.check_field = -1
        keep = [off, nby]
        if 'chrom' in names:
            ft = dtype.fields['chrom'][0]
            if ft.kind != 'S':
                return None
            cb = np.zeros(1, dtype=ft)
            cb[0] = chrom.encode('ascii')
            if cb[0] != chrom.encode('ascii'):
                return None
            keep.append(cb)
            L.chrom, L.chrom_off, L.chrom_bytes = cb.ctypes.data, dtype.fields['chrom'][1], ft.itemsize
        if 'pos' in names:
            if dtype.fields['pos'][0] != np.dtype('<i4'):
                return None
            L.pos_off = dtype.fields['pos'][1]
        if 'ref' in names:
            if dtype.fields['ref'][0] != np.dtype('S1'):
                return None
            L.ref_off = dtype.fields['ref'][1]
        if check_field is not None and check_field in value_fields and check_field in names:
            L.check_field = list(value_fields).index(check_field)
        L.field_off, L.field_bytes = off.ctypes.data, nby.ctypes.data
        L.pos_add = int(pos_add)
        L.pad_start, L.pad_end = (int(pad[0]), int(pad[1])) if pad else (0, 0)
        L._keep = keep
        return L

    def pileup_host_rows(self, reads_struct, params, layout, rows, row0, capacity):
        """Records of one statistic into rows[row0:row0 + capacity] (a C-contiguous structured array);
        returns (rows written, negative values seen in the layout's check_field)."""
        n, neg = C.c_int64(0), C.c_int64(0)
        self.check(self.lib.pss_pileup_host_rows(self.h, C.byref(reads_struct), C.byref(params), C.byref(layout),
                                                 rows.ctypes.data + int(row0) * rows.dtype.itemsize, int(capacity),
                                                 C.byref(n), C.byref(neg)))
        return int(n.value), int(neg.value)

    def pileup_host_stream(self, reads_struct, params_list, outs, rowmasks, max_span, chunk_positions=0):
        n = len(params_list)
        pa = (PssPileupParams * n)(*params_list)
        oa = (_P * n)(*[_addr(o) for o in outs])
        ma = (_P * n)(*[_addr(m) for m in rowmasks])
        self.check(self.lib.pss_pileup_host_stream(self.h, C.byref(reads_struct), n, pa, oa, ma, int(max_span),
                                                   int(chunk_positions)))

    def pileup(self, reads_struct, params, out, rowmask, stream=None):
        self.check(self.lib.pss_pileup(self.h, C.byref(reads_struct), C.byref(params), _addr(out),
                                       _addr(rowmask), stream))

    def prepare(self, reads_struct, stepper='all', max_depth=8000, min_mapq=0, no_dup=False, with_qual=False,
                stream=None):
        if isinstance(stepper, str):
            if stepper not in STEPPERS:
                raise ValueError('unknown stepper: %r' % stepper)
            stepper = STEPPERS[stepper]
        self.check(self.lib.pss_prepare(self.h, C.byref(reads_struct), int(stepper), int(max_depth or 0), int(min_mapq),
                                        int(bool(no_dup)), int(bool(with_qual)), stream))

    def pileup_prepared(self, params, out, rowmask, stream=None):
        self.check(self.lib.pss_pileup_prepared(self.h, C.byref(params), _addr(out), _addr(rowmask), stream))

    def binned_params(self, stat, start, end, window_size, min_mapq=0, no_dup=False, ref=None, ref_start=0,
                      ref_len=0):
        p = PssBinnedParams()
        p.stat = BINNED_STATS.index(stat) if isinstance(stat, str) else int(stat)
        p.window_size = int(window_size)
        p.min_mapq, p.no_dup = int(min_mapq), int(bool(no_dup))
        p.start, p.end = int(start), int(end)
        p.n_bins = self.lib.pss_binned_nbins(int(start), int(end), int(window_size))
        p.ref = _addr(ref)
        p.ref_start, p.ref_len = int(ref_start), int(ref_len)
        return p

    def binned_host(self, reads_struct, params, out):
        self.check(self.lib.pss_binned_host(self.h, C.byref(reads_struct), C.byref(params), _addr(out)))

    def binned(self, reads_struct, params, out, stream=None):
        self.check(self.lib.pss_binned(self.h, C.byref(reads_struct), C.byref(params), _addr(out), stream))

    def binned_multi(self, reads_struct, params_list, outs, stream=None):
        n = len(params_list)
        pa = (PssBinnedParams * n)(*params_list)
        oa = (_P * n)(*[_addr(o) for o in outs])
        self.check(self.lib.pss_binned_multi(self.h, C.byref(reads_struct), n, pa, oa, stream))

    def binned_host_multi(self, reads_struct, params_list, outs):
        n = len(params_list)
        pa = (PssBinnedParams * n)(*params_list)
        oa = (_P * n)(*[_addr(o) for o in outs])
        self.check(self.lib.pss_binned_host_multi(self.h, C.byref(reads_struct), n, pa, oa))

    def admit(self, reads_struct, stepper, max_depth, admitted, stream=None):
        self.check(self.lib.pss_admit(self.h, C.byref(reads_struct), int(stepper), int(max_depth),
                                      _addr(admitted), stream))

    def overlap_tweak(self, reads_struct, stepper, max_depth, qual_out, pair_first=None, stream=None):
        self.check(self.lib.pss_overlap_tweak(self.h, C.byref(reads_struct), int(stepper), int(max_depth),
                                              _addr(qual_out), _addr(pair_first), stream))

    def pairs_host(self, reads_struct, stepper, max_depth):
        """pair_first (int32 per read: index of the earlier mate the read's qualities are rewritten against, or -1) of a
        whole host batch from its metadata alone -- no sequences or qualities cross the bus (pss_pairs_host)."""
        pair = np.full(int(reads_struct.n_reads), -1, dtype=np.int32)
        step = STEPPERS[stepper] if isinstance(stepper, str) else int(stepper)
        self.check(self.lib.pss_pairs_host(self.h, C.byref(reads_struct), step, int(max_depth or 0), _addr(pair)))
        return pair

    def synchronize(self, stream=None):
        self.check(self.lib.pss_synchronize(self.h, stream))

    def launch_count(self):
        return int(self.lib.pss_launch_count(self.h))

    def transfer_bytes(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self.lib.pss_transfer_bytes(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def set_kernel_timing(self, on):
        self.lib.pss_set_kernel_timing(self.h, int(bool(on)))

    def pileup_kernel_ms(self, stat=-1):
        if isinstance(stat, str):
            stat = PILEUP_STATS.index(stat)
        n = C.c_int64(0)
        ms = self.lib.pss_pileup_kernel_ms(self.h, int(stat), C.byref(n))
        return float(ms), int(n.value)


_contexts = {}


def default_device():
    for k in ('PSS_DEVICE', 'LOCAL_RANK'):
        if os.environ.get(k, '') != '':
            return int(os.environ[k])
    return 0


def get_context(device=None, slot=0):
    """Process-wide context per (device, slot) (created on first use; fails loudly without a GPU).  A context serves
    one host thread at a time; ``slot`` > 0 gives the further threads of a multi-shard call their own context on a
    device that is listed more than once (``PSS_GPUS=0,0``)."""
    if device is None:
        device = default_device()
    key = (int(device), int(slot))
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = _contexts[key] = Context(device)
    return ctx
