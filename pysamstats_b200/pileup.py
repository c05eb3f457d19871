"""Per-position statistics: ``stat_pileup`` / ``load_pileup`` and the 15 named pairs.

Public signatures, defaults, error behaviour and function names are those of the
reference's ``pysamstats/pileup.py:53-205``; the work happens in ``opt`` (CUDA engine).
"""
import functools

from . import config, opt, util

_doc_params = """
    Parameters
    ----------
    type : string
        Statistics type. One of "coverage", "coverage_strand", "coverage_ext",
        "coverage_ext_strand", "variation", "variation_strand", "tlen", "tlen_strand", "mapq",
        "mapq_strand", "baseq", "baseq_strand", "baseq_ext", "baseq_ext_strand", "coverage_gc".
    alignmentfile : path, or an alignment file object exposing ``.filename``
        BAM file (e.g. a ``pysam.AlignmentFile``).
    fafile : path or FASTA file object, optional
        Only required for the statistics that report the reference base or GC content.
    chrom, start, end : optional
        Region; the whole genome when ``chrom`` is None.
    one_based : bool, optional
        Coordinate system, False if zero-based (default), True if one-based.
    truncate : bool, optional
        If True, truncate output to the selected region.
    stepper : {"all", "nofilter", "samtools"}, optional
        Which reads reach the pileup (flag filter).
    pad : bool, optional
        If True, emit records for every position, even if no reads are aligned.
    max_depth : int, optional
        Maximum depth to allow in a pileup column.
    window_size, window_offset : int, optional
        Window for the percent GC calculation (coverage_gc only).
    min_mapq, min_baseq : int, optional
        Only reads / bases with at least this quality are counted (0 by default).
    no_del : bool, optional
        If True, don't count reads aligned with a deletion at the current position.
    no_dup : bool, optional
        If True, don't count reads flagged as duplicate."""

stats_classes_pileup = {
    'coverage': opt.Coverage,
    'coverage_strand': opt.CoverageStrand,
    'coverage_ext': opt.CoverageExt,
    'coverage_ext_strand': opt.CoverageExtStrand,
    'variation': opt.Variation,
    'variation_strand': opt.VariationStrand,
    'tlen': opt.Tlen,
    'tlen_strand': opt.TlenStrand,
    'mapq': opt.Mapq,
    'mapq_strand': opt.MapqStrand,
    'baseq': opt.Baseq,
    'baseq_strand': opt.BaseqStrand,
    'baseq_ext': opt.BaseqExt,
    'baseq_ext_strand': opt.BaseqExtStrand,
    'coverage_gc': opt.CoverageGC,
}


def _make_stat(type, fafile, window_size, window_offset):
    if type in config.stats_types_withref and fafile is None:
        raise ValueError('reference sequence is required; please provide fafile argument')
    try:
        cls = stats_classes_pileup[type]
    except (KeyError, TypeError):
        raise ValueError('unsupported statistics type: %r' % (type,))
    if type == 'coverage_gc':
        return cls(window_size=window_size, window_offset=window_offset)
    return cls()


# noinspection PyShadowingBuiltins
def stat_pileup(type, alignmentfile, fafile=None, chrom=None, start=None, end=None, one_based=False,
                truncate=False, stepper="all", pad=False, max_depth=8000, window_size=300,
                window_offset=None, min_mapq=0, min_baseq=0, no_del=False, no_dup=False):
    """Generate statistics per genome position, based on read pileups.
    {params}

    Returns
    -------
    recs : iterator
        An iterator yielding dict objects, one per genome position.
    """
    stat = _make_stat(type, fafile, window_size, window_offset)
    return opt.iter_pileup(stat, alignmentfile=alignmentfile, fafile=fafile, chrom=chrom, start=start,
                           end=end, one_based=one_based, truncate=truncate, stepper=stepper, pad=pad,
                           max_depth=max_depth, min_mapq=min_mapq, min_baseq=min_baseq, no_del=no_del,
                           no_dup=no_dup)


def _pileup_tables(type, fafile=None, window_size=300, window_offset=None, **kwargs):
    stat = _make_stat(type, fafile, window_size, window_offset)
    return opt.pileup_tables(stat, fafile=fafile, **kwargs)


def _pileup_rows(type, dtype, fafile=None, window_size=300, window_offset=None, **kwargs):
    stat = _make_stat(type, fafile, window_size, window_offset)
    return opt.pileup_rows(stat, dtype, fafile=fafile, **kwargs)


# noinspection PyShadowingBuiltins
def load_pileup(type, alignmentfile, fafile=None, chrom=None, start=None, end=None, one_based=False,
                truncate=False, stepper="all", pad=False, max_depth=8000, window_size=300,
                window_offset=None, min_mapq=0, min_baseq=0, no_del=False, no_dup=False, dtype=None,
                fields=None):
    """Load statistics per genome position, based on read pileups.
    {params}
    dtype : dtype
        Override default dtype.
    fields : string or list of strings
        Select a subset of fields to load.

    Returns
    -------
    ra : numpy structured array
    """
    try:
        default_dtype = getattr(config, 'dtype_' + type)
    except (AttributeError, TypeError):
        raise ValueError('unsupported statistics type: %r' % (type,))
    statfun = functools.partial(stat_pileup, type)
    statfun._pss_tables = functools.partial(_pileup_tables, type)     # column tables / packed records of the same
    statfun._pss_rows = functools.partial(_pileup_rows, type)         # statistic straight from the engine
    return util.load_stats(statfun, user_dtype=dtype, default_dtype=default_dtype, user_fields=fields,
                           alignmentfile=alignmentfile, fafile=fafile, chrom=chrom, start=start, end=end,
                           one_based=one_based, truncate=truncate, stepper=stepper, pad=pad, max_depth=max_depth,
                           window_size=window_size, window_offset=window_offset, min_mapq=min_mapq,
                           min_baseq=min_baseq, no_del=no_del, no_dup=no_dup)


stat_pileup.__doc__ = stat_pileup.__doc__.format(params=_doc_params)
load_pileup.__doc__ = load_pileup.__doc__.format(params=_doc_params)

__all__ = ['stat_pileup', 'load_pileup', 'stats_classes_pileup']


def _specialize(type):
    stat = functools.partial(stat_pileup, type)
    stat.__doc__ = stat_pileup.__doc__
    stat.__name__ = 'stat_' + type
    load = functools.partial(load_pileup, type)
    load.__doc__ = load_pileup.__doc__
    load.__name__ = 'load_' + type
    return stat, load


for _t in stats_classes_pileup:
    _s, _l = _specialize(_t)
    globals()['stat_' + _t] = _s
    globals()['load_' + _t] = _l
    __all__ += ['stat_' + _t, 'load_' + _t]
del _t, _s, _l
