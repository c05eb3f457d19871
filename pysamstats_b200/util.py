"""Array assembly for ``load_*`` (reference: ``pysamstats/util.py:31-81``).

The reference turns a generator of dicts into a structured array with ``np.fromiter``; here
the engine already returns columns, so the array is filled column-wise.  The typing rules
are the reference's: default dtype list per statistic, ``chrom`` widened to the longest
contig name of the file, user dtype overrides, optional field subset (a single field
gives a plain ndarray, unknown fields raise ``ValueError``).
"""
import ctypes as C
import os

import numpy as np

from . import bamio


def _pack_native(a, o, m, fields, chrom, pos, ref, cols):
    """Fill records a[o:o + m] through libpss_host's row packer; False when a column's type is not a plain
    byte copy of what the engine produced (user dtype overrides), so the numpy path has to convert."""
    lib = bamio.host_lib()
    if lib is None or m == 0:
        return False
    dt = a.dtype
    keep, src, stride, nbytes, off = [], [], [], [], []
    for f in fields:
        ft, fo = dt.fields[f][0], dt.fields[f][1]
        if f == 'chrom':
            if ft.kind != 'S':
                return False
            v = np.zeros(1, dtype=ft)
            v[0] = chrom.encode('ascii')
            col, st = v, 0
        elif f == 'pos':
            if ft != np.dtype('<i4') or pos.dtype != np.dtype('<i8'):
                return False
            col, st = np.ascontiguousarray(pos), 8            # low half of the little-endian int64
        elif f == 'ref':
            if ft != np.dtype('S1'):
                return False
            col, st = np.ascontiguousarray(ref), 1
        else:
            col = cols[f]
            if ft != col.dtype or not col.flags.c_contiguous:
                return False
            st = col.dtype.itemsize
        keep.append(col)
        src.append(col.ctypes.data)
        stride.append(st)
        nbytes.append(ft.itemsize)
        off.append(fo)
    n = len(fields)
    rc = lib.pss_pack_rows(a.ctypes.data + o * dt.itemsize, m, dt.itemsize, n, (C.c_void_p * n)(*src),
                           (C.c_int32 * n)(*stride), (C.c_int32 * n)(*nbytes), (C.c_int32 * n)(*off), 0)
    if rc != 0:
        raise ValueError(lib.pss_host_error().decode('ascii', 'replace'))
    return True


def flatten(recs, *fields):
    """Convert a record (dict) iterator to a row (tuple) iterator (``util.py:7-28``)."""
    from operator import itemgetter
    getter = itemgetter(*fields)
    return (getter(rec) for rec in recs)


def determine_max_seqid(alignmentfile):
    """Longest contig name in the header (``util.py:76-81``)."""
    return max(len(x) for x in bamio.open_bam(alignmentfile).references)


def load_stats(statfun, default_dtype, user_dtype, user_fields, **kwargs):
    """Assemble the structured array ``load_*`` returns (``util.py:31-73``, same signature).

    ``statfun(**kwargs)`` is a generator of dicts, as in the reference.  The engine's own
    ``stat_*`` partials carry two attributes, ``_pss_tables`` (a generator of per-contig column
    tables, ``opt.pileup_tables``) and optionally ``_pss_rows`` (``opt.pileup_rows``: records packed
    by the engine's host pipeline); for them the array is filled column-wise / by the engine instead
    of row by row.  Any other ``statfun`` takes the reference's path (``flatten`` + ``np.fromiter``)."""
    default_fields = [t[0] for t in default_dtype]
    if user_fields is None:
        fields = default_fields
    else:
        fields = [user_fields] if isinstance(user_fields, str) else list(user_fields)
        if any(f not in default_fields for f in fields):
            raise ValueError('invalid fields: %r' % (user_fields,))
    dtype = dict(default_dtype)
    dtype['chrom'] = 'S{0}'.format(determine_max_seqid(kwargs['alignmentfile']))
    if user_dtype is not None:
        dtype.update(dict(user_dtype))
    single = len(fields) == 1
    dt = np.dtype(dtype[fields[0]] if single else [(f, dtype[f]) for f in fields])

    tablefun = getattr(statfun, '_pss_tables', None)
    rowsfun = getattr(statfun, '_pss_rows', None)
    if tablefun is None:
        # a caller-supplied generator of dicts: util.py:61-71
        a = np.fromiter(flatten(statfun(**kwargs), *fields), dtype=dt)
        return a if single else a.view(np.recarray)
    if rowsfun is not None and not single and os.environ.get('PSS_NO_ROWS') is None:
        recs = rowsfun(dt, **kwargs)
        if recs is not None:
            a = recs[0] if len(recs) == 1 else (np.concatenate(recs) if recs else np.zeros(0, dtype=dt))
            return a.view(np.recarray)
    parts = list(tablefun(**kwargs))
    n = sum(len(p[1]) for p in parts)
    a = np.zeros(n, dtype=dt) if single or bamio.host_lib() is None else bamio.alloc_big(n, dt)
    o = 0
    for chrom, pos, ref, cols in parts:
        m = len(pos)
        if not single and _pack_native(a, o, m, fields, chrom, pos, ref, cols):
            o += m
            continue
        for f in fields:
            dst = a[o:o + m] if single else a[f][o:o + m]
            if f == 'chrom':
                dst[...] = chrom.encode('ascii')
            elif f == 'pos':
                dst[...] = pos
            elif f == 'ref':
                dst[...] = ref
            else:
                dst[...] = cols[f]
        o += m
    if not single:
        a = a.view(np.recarray)
    return a
