"""Per-window statistics keyed by read start: ``stat_binned`` / ``load_binned`` and the five
named pairs.  Signatures, defaults and names follow the reference's
``pysamstats/binned.py:41-156``.
"""
import functools

from . import config, opt, util

_doc_params = """
    Parameters
    ----------
    type : string
        Statistics type. One of "coverage", "coverage_ext", "mapq", "alignment", "tlen".
    alignmentfile : path, or an alignment file object exposing ``.filename``
    fafile : path or FASTA file object
        Required for "coverage" and "coverage_ext" (window GC content).
    chrom, start, end : optional
        Region; the whole genome when ``chrom`` is None.
    one_based : bool
        Coordinate system, False if zero-based (default), True if one-based.
    window_size : int
        Window size to use.
    window_offset : int
        Distance from window start to record position (default: half the window).
    min_mapq : int, optional
        Only reads with at least this mapping quality are counted (0 by default).
    no_dup : bool, optional
        If True, don't count reads flagged as duplicate."""

stats_classes_binned = {
    'coverage': opt.CoverageBinned,
    'coverage_ext': opt.CoverageExtBinned,
    'mapq': opt.MapqBinned,
    'alignment': opt.AlignmentBinned,
    'tlen': opt.TlenBinned,
}


def _make_stat(type):
    try:
        return stats_classes_binned[type]()
    except (KeyError, TypeError):
        raise ValueError('unsupported statistics type: %r' % (type,))


# noinspection PyShadowingBuiltins
def stat_binned(type, alignmentfile, fafile=None, chrom=None, start=None, end=None, one_based=False,
                window_size=300, window_offset=None, min_mapq=0, no_dup=False):
    """Generate statistics per genome window, based on all reads whose alignment starts within
    the window.
    {params}

    Returns
    -------
    recs : iterator
        An iterator yielding dict objects, one per window.
    """
    return opt.iter_binned(_make_stat(type), alignmentfile=alignmentfile, fafile=fafile, chrom=chrom,
                           start=start, end=end, one_based=one_based, window_size=window_size,
                           window_offset=window_offset, min_mapq=min_mapq, no_dup=no_dup)


def _binned_tables(type, **kwargs):
    return opt.binned_tables(_make_stat(type), **kwargs)


# noinspection PyShadowingBuiltins
def load_binned(type, alignmentfile, fafile=None, chrom=None, start=None, end=None, one_based=False,
                window_size=300, window_offset=None, min_mapq=0, no_dup=False, dtype=None, fields=None):
    """Load statistics per genome window, based on all reads whose alignment starts within
    the window.
    {params}
    dtype : dtype
        Override default dtype.
    fields : string or list of strings
        Select a subset of fields to load.

    Returns
    -------
    ra : numpy structured array
    """
    # like the reference (binned.py:106-109) an unknown type surfaces as AttributeError here
    default_dtype = getattr(config, 'dtype_' + type + '_binned')
    statfun = functools.partial(stat_binned, type)
    statfun._pss_tables = functools.partial(_binned_tables, type)
    return util.load_stats(statfun, user_dtype=dtype, default_dtype=default_dtype, user_fields=fields,
                           alignmentfile=alignmentfile, fafile=fafile, chrom=chrom, start=start, end=end,
                           one_based=one_based, window_size=window_size, window_offset=window_offset,
                           min_mapq=min_mapq, no_dup=no_dup)


stat_binned.__doc__ = stat_binned.__doc__.format(params=_doc_params)
load_binned.__doc__ = load_binned.__doc__.format(params=_doc_params)

__all__ = ['stat_binned', 'load_binned', 'stats_classes_binned']


def _specialize(type):
    statfun = functools.partial(stat_binned, type)
    statfun.__doc__ = stat_binned.__doc__
    statfun.__name__ = 'stat_' + type
    loadfun = functools.partial(load_binned, type)
    loadfun.__doc__ = load_binned.__doc__
    loadfun.__name__ = 'load_' + type
    return statfun, loadfun


for _t in stats_classes_binned:
    _s, _l = _specialize(_t)
    globals()['stat_' + _t + '_binned'] = _s
    globals()['load_' + _t + '_binned'] = _l
    __all__ += ['stat_' + _t + '_binned', 'load_' + _t + '_binned']
del _t, _s, _l
