"""Statistics types and output record layouts.

Same public names and the same field order / dtypes as the reference's registry
(``pysamstats/config.py:5-402``): the field lists ARE the output contract of ``load_*``
and of the tsv/csv/hdf5 writers.  They are generated from the counter families instead
of being spelled out (``S<n>`` is numpy's current spelling of the reference's ``a<n>``).
"""

stats_types_noref = ('coverage', 'coverage_strand', 'coverage_ext', 'coverage_ext_strand', 'tlen',
                     'tlen_strand', 'mapq', 'mapq_strand', 'baseq', 'baseq_strand', 'mapq_binned',
                     'alignment_binned', 'tlen_binned')
stats_types_withref = ('variation', 'variation_strand', 'baseq_ext', 'baseq_ext_strand', 'coverage_gc',
                       'coverage_binned', 'coverage_ext_binned')
stats_types = sorted(stats_types_noref + stats_types_withref)
stepper_types = ('nofilter', 'samtools', 'all')

_PP = ('', '_pp')
_STRAND = ('', '_fwd', '_rev')
_PP_STRAND = ('', '_fwd', '_rev', '_pp', '_pp_fwd', '_pp_rev')
_MATE_CLASSES = ('reads_all', 'reads_pp', 'reads_mate_unmapped', 'reads_mate_other_chr',
                 'reads_mate_same_strand', 'reads_faceaway', 'reads_softclipped', 'reads_duplicate')
_EVENTS = ('reads', 'matches', 'mismatches', 'deletions', 'insertions', 'A', 'C', 'T', 'G', 'N')


def _n(base, suffix):
    return 'reads_all' if (base, suffix) == ('reads', '') else base + suffix


def _cross(bases, suffixes):
    return [_n(b, s) for b in bases for s in suffixes]


def value_fields(stat):
    """Names of the int32 value columns of a pileup statistic, engine order."""
    if stat == 'coverage':
        return _cross(['reads'], _PP)
    if stat == 'coverage_strand':
        return _cross(['reads'], _PP_STRAND)
    if stat == 'coverage_ext':
        return list(_MATE_CLASSES)
    if stat == 'coverage_ext_strand':
        return [_n('reads' if c == 'reads_all' else c, s) for c in _MATE_CLASSES for s in _STRAND]
    if stat == 'variation':
        return _cross(_EVENTS, _PP)
    if stat == 'variation_strand':
        return _cross(_EVENTS, _PP_STRAND)
    if stat == 'tlen':
        return ['reads_all', 'reads_paired', 'reads_pp'] + _cross(['mean_tlen', 'rms_tlen', 'std_tlen'], _PP)
    if stat == 'tlen_strand':
        return (['reads_all', 'reads_fwd', 'reads_rev'] + _cross(['reads_paired', 'reads_pp'], _STRAND) +
                _cross(['mean_tlen', 'rms_tlen', 'std_tlen'], _PP_STRAND))
    if stat == 'mapq':
        return _cross(['reads', 'reads_mapq0', 'rms_mapq', 'max_mapq'], _PP)
    if stat == 'mapq_strand':
        return _cross(['reads', 'reads_mapq0', 'rms_mapq', 'max_mapq'], _PP_STRAND)
    if stat == 'baseq':
        return _cross(['reads', 'rms_baseq'], _PP)
    if stat == 'baseq_strand':
        return _cross(['reads', 'rms_baseq'], _PP_STRAND)
    if stat in ('baseq_ext', 'baseq_ext_strand'):
        sfx = _PP if stat == 'baseq_ext' else _PP_STRAND
        return _cross(['reads', 'matches', 'mismatches', 'rms_baseq', 'rms_baseq_matches',
                       'rms_baseq_mismatches'], sfx)
    if stat == 'coverage_gc':
        return ['gc', 'reads_all', 'reads_pp']
    raise KeyError(stat)


def value_fields_binned(stat):
    """Names of the value columns of a binned statistic, engine order."""
    if stat == 'coverage':
        return ['gc', 'reads_all', 'reads_pp']
    if stat == 'coverage_ext':
        return ['gc'] + list(_MATE_CLASSES)
    if stat == 'mapq':
        return ['reads_all', 'reads_mapq0', 'rms_mapq']
    if stat == 'alignment':
        return ['reads_all', 'bases_all', 'M', 'I', 'D', 'N', 'S', 'H', 'P', '=', 'X']
    if stat == 'tlen':
        return ['reads_all', 'reads_pp'] + _cross(['mean_tlen', 'rms_tlen'], _PP)
    raise KeyError(stat)


def _dtype(fields, with_ref):
    head = [('chrom', 'S12'), ('pos', 'i4')] + ([('ref', 'S1')] if with_ref else [])
    return head + [(f, 'u1' if f == 'gc' else 'i4') for f in fields]


pileup_types = ('coverage', 'coverage_strand', 'coverage_ext', 'coverage_ext_strand', 'variation',
                'variation_strand', 'tlen', 'tlen_strand', 'mapq', 'mapq_strand', 'baseq', 'baseq_strand',
                'baseq_ext', 'baseq_ext_strand', 'coverage_gc')
binned_types = ('coverage', 'coverage_ext', 'mapq', 'alignment', 'tlen')
ref_column_types = ('variation', 'variation_strand', 'baseq_ext', 'baseq_ext_strand')

for _t in pileup_types:
    globals()['dtype_' + _t] = _dtype(value_fields(_t), _t in ref_column_types)
for _t in binned_types:
    globals()['dtype_' + _t + '_binned'] = _dtype(value_fields_binned(_t), False)
del _t
