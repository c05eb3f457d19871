"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, 'tests', 'golden', 'fixture')


def fx(name):
    return os.path.join(FIX, name)


def assert_recarrays_equal(expect, actual, label=''):
    """Field-by-field exact comparison of two structured arrays (test_regression.py:49-57)."""
    assert expect.dtype.names == actual.dtype.names, (label, expect.dtype.names, actual.dtype.names)
    assert expect.shape == actual.shape, (label, expect.shape, actual.shape)
    for k in expect.dtype.names:
        e, a = expect[k], actual[k]
        if not np.array_equal(e, a):
            bad = np.flatnonzero(e != a)
            raise AssertionError('%s field %r: %d mismatches, first at row %d: expect %r actual %r (pos %r)'
                                 % (label, k, len(bad), bad[0], e[bad[0]], a[bad[0]],
                                    expect['pos'][bad[0]] if 'pos' in expect.dtype.names else None))


def oracle_reads_from_arrays(nb):
    """oracle.PackedReads from the numpy SoA arrays of a synthetic batch."""
    from oracle import oracle as orc
    n = int(nb['n_reads'])
    r = orc.PackedReads.__new__(orc.PackedReads)
    r.n = n
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar', 'cigar_off', 'seq4', 'seq_off'):
        setattr(r, k, np.ascontiguousarray(nb[k]))
    if nb.get('qual') is not None:
        r.qual, r.qual_off = np.ascontiguousarray(nb['qual']), np.ascontiguousarray(nb['qual_off'])
    else:
        r.qual_off = np.ascontiguousarray(nb['seq_off'] * 0)
        r.qual = np.zeros(1, dtype='u1')
    prev = nb['name_prev']
    ids = np.empty(n, dtype='i8')
    nxt = 0
    for i in range(n):
        if prev[i] >= 0:
            ids[i] = ids[prev[i]]
        else:
            ids[i] = nxt
            nxt += 1
    r.name_id = ids
    r.n_names = nxt
    r.tid = int(nb.get('tid', 0))
    return r
