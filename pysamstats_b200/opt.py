"""The engine seam: drop-in for the reference's compiled module ``pysamstats/opt.pyx``.

The reference's python layer calls ``opt.iter_pileup(stat, ...)`` (``pileup.py:91``) and
``opt.iter_binned(stat, ...)`` (``binned.py:68``) with an accumulator object taken from its
class tables and consumes a generator of dicts.  The same names live here: the "stat"
objects only describe WHICH statistic to compute (the per-read ``recv`` protocol of the
reference is ``cdef`` and never was a python-level plugin point), the
work is done by CUDA kernels behind the C ABI (``include/pss_b200.h``).

Two levels are offered:

* ``pileup_tables`` / ``binned_tables`` -- per-contig column arrays straight from the
  engine (what ``load_*`` assembles recarrays from, no per-row python);
* ``iter_pileup`` / ``iter_binned`` -- the reference's generator-of-dicts interface, built
  lazily on top of the tables.

There is no CPU path: everything below ends in ``_cabi.get_context()``, which raises when
the CUDA library or a GPU is missing.
"""
import numpy as np

from . import _cabi, bamio, config


# --------------------------------------------------------------------- stat descriptors

class PileupStat(object):
    """Which per-position statistic to compute (reference: ``opt.pyx:75-81``)."""
    name = None

    def __init__(self):
        self.window_size = 300
        self.window_offset = None


class BinnedStat(object):
    """Which per-window statistic to compute (reference: ``opt.pyx:1542-1548``)."""
    name = None


def _pileup_cls(clsname, stat):
    return type(clsname, (PileupStat,), {'name': stat, '__doc__': 'reference: opt.pyx class %s' % clsname})


Coverage = _pileup_cls('Coverage', 'coverage')
CoverageStrand = _pileup_cls('CoverageStrand', 'coverage_strand')
CoverageExt = _pileup_cls('CoverageExt', 'coverage_ext')
CoverageExtStrand = _pileup_cls('CoverageExtStrand', 'coverage_ext_strand')
Variation = _pileup_cls('Variation', 'variation')
VariationStrand = _pileup_cls('VariationStrand', 'variation_strand')
Tlen = _pileup_cls('Tlen', 'tlen')
TlenStrand = _pileup_cls('TlenStrand', 'tlen_strand')
Mapq = _pileup_cls('Mapq', 'mapq')
MapqStrand = _pileup_cls('MapqStrand', 'mapq_strand')
Baseq = _pileup_cls('Baseq', 'baseq')
BaseqStrand = _pileup_cls('BaseqStrand', 'baseq_strand')
BaseqExt = _pileup_cls('BaseqExt', 'baseq_ext')
BaseqExtStrand = _pileup_cls('BaseqExtStrand', 'baseq_ext_strand')


class CoverageGC(PileupStat):
    """reference: ``opt.pyx:1471-1481`` (window_offset defaults to window_size // 2)."""
    name = 'coverage_gc'

    def __init__(self, window_size=300, window_offset=None):
        self.window_size = window_size
        self.window_offset = window_size // 2 if window_offset is None else window_offset


def _binned_cls(clsname, stat):
    return type(clsname, (BinnedStat,), {'name': stat, '__doc__': 'reference: opt.pyx class %s' % clsname})


CoverageBinned = _binned_cls('CoverageBinned', 'coverage')
CoverageExtBinned = _binned_cls('CoverageExtBinned', 'coverage_ext')
MapqBinned = _binned_cls('MapqBinned', 'mapq')
AlignmentBinned = _binned_cls('AlignmentBinned', 'alignment')
TlenBinned = _binned_cls('TlenBinned', 'tlen')


# ------------------------------------------------------------------------------ helpers

def normalise_coords(alignmentfile, chrom, start, end, one_based):
    """Zero-based half-open coordinates with defaults filled in (``opt.pyx:2105-2130``,
    including its quirk that a one-based ``end`` is decremented and stays exclusive)."""
    if chrom is None:
        return None, None
    assert chrom in alignmentfile.references, 'chromosome not in SAM references: %s' % chrom
    if one_based:
        start = start - 1 if start is not None else None
        end = end - 1 if end is not None else None
    chrlen = alignmentfile.lengths[alignmentfile.references.index(chrom)]
    if start is None:
        start = 0
    if end is None:
        end = chrlen
    if end > chrlen:
        end = chrlen
    return start, end


def _check_region(start, end):
    """The checks of pysam's ``AlignmentFile.parse_region``, which both reference call sites reach with the
    normalised coordinates (``opt.pyx:1853``, ``:2036``): same exception type and message."""
    if start > end:
        raise ValueError('invalid coordinates: start (%i) > stop (%i)' % (start, end))
    if start < 0:
        raise ValueError('start out of range (%i)' % start)          # e.g. one_based=True with start=0
    if end < 0:
        raise ValueError('end out of range (%i)' % end)


def _reads_struct(batch, need_qual):
    batch.padded()
    d = {k: getattr(batch, k) for k in _cabi.READ_FIELDS if need_qual or k != 'name_prev'}
    d['tid'] = batch.tid
    d['n_reads'] = batch.n_reads
    if not need_qual:
        d['qual'] = d['qual_off'] = d['name_prev'] = None
        d['n_qual_bytes'] = 0
    r = _cabi.make_reads_struct(d)
    r._keep = batch             # the struct holds raw addresses: the arrays live as long as it does
    return r


def _ref_column(refseq, pos):
    """Upper-cased reference base per position, '' beyond the FASTA (``opt.pyx:1987-1994``)."""
    col = np.zeros(len(pos), dtype='S1')
    ok = (pos >= 0) & (pos < len(refseq))
    b = refseq[pos[ok]]
    lower = (b >= 97) & (b <= 122)
    b = np.where(lower, b - 32, b).astype(np.uint8)
    col[ok] = b.view('S1')
    return col


# ------------------------------------------------------------------------ pileup tables

def _table_chunk():
    """Positions per piece of a contig (PSS_TABLE_CHUNK, default 16 Mi), a multiple of the scatter kernel's tile."""
    import os
    from . import shard
    return max(shard.TILE * 2, int(os.environ.get('PSS_TABLE_CHUNK', 16 << 20)) // shard.TILE * shard.TILE)


def _refusable(batch, max_depth):
    """True when the order-dependent ``max_depth`` rule could refuse a read of this batch at all: some read has
    ``max_depth - 1`` earlier reads starting within one maximal read span before it (start sorted: one subtraction
    over the positions).  Cached per batch."""
    k = (max_depth or 8000) - 1
    key = ('cut', k)
    if key not in batch.cache:
        n = batch.n_reads
        maxspan = int((batch.endpos - batch.pos).max()) if n else 1
        batch.cache[key] = (maxspan, k <= 0 or (n > k and int((batch.pos[k:] - batch.pos[:n - k]).min()) <= maxspan))
    return batch.cache[key]


def _streamed_pieces(bam, c, s, e, truncate, pad, max_depth, state):
    """Position pieces of chrom:[s, e) whose reads are decoded piece by piece through the BAI index while the earlier
    pieces are on the GPU (``BamFile.piece_batches``): yields (c_lo, c_hi, struct pss_reads).  ``state`` receives
    'lo' (first position that can get a row) before the first piece is yielded.

    A cut is exact when every read overlapping it is admitted whatever came before (``shard`` docstring): a piece in
    which the ``max_depth`` rule could refuse a read (``_refusable``) is therefore not followed by another cut -- the rest
    of the region is decoded and computed as one piece."""
    chunk = _table_chunk()
    cuts = list(range(s, e, chunk)) + [e]
    n = len(cuts) - 1
    k = 0
    tail = None
    for batch in bam.piece_batches(c, cuts):
        if batch is None:                               # unusable index: the rest in one ordinary region query
            tail = bam.region_batch(c, cuts[k], e)
        elif _refusable(batch, max_depth)[1] and k + 1 < n:
            tail = bam._region_native(bam.get_tid(c), cuts[k], e, bam.index_path(), True) or bam.region_batch(c, cuts[k], e)
        if tail is not None:
            batch = tail
        last = tail is not None or k == n - 1
        g_lo, g_hi = cuts[k], (e if last else cuts[k + 1])
        nr = batch.n_reads
        # rows: [s, e) with truncate; else every position a fetched read covers (plus [s, e) with pad)
        c_lo, c_hi = g_lo, g_hi
        if not truncate and nr:
            if k == 0:
                c_lo = min(int(batch.pos[0]), s) if pad else int(batch.pos[0])
            if last:
                c_hi = max(int(batch.endpos.max()), e) if pad else int(batch.endpos.max())
        if 'lo' not in state:
            state['lo'] = c_lo
        if (nr or pad) and c_hi > c_lo:
            yield c_lo, c_hi, _reads_struct(batch, False)
        if last:
            return
        k += 1


def _pileup_plan(stat, alignmentfile, fafile, pad, chrom, start, end, one_based, truncate, stepper, max_depth,
                 min_mapq, min_baseq, no_del, no_dup):
    """Per contig of the request: (contig, s, e, lo, hi, pieces, reference bytes, engine keyword arguments,
    need_qual) -- [s, e) the normalised region, [lo, hi) the positions that can get a row (row selection of
    ``iter_pileup_default`` / ``iter_pileup_padded_chrom``, ``opt.pyx:1846-1863``, ``:1936-1984``; ``hi`` may be an
    upper bound), ``pieces`` an iterator of (c_lo, c_hi, struct pss_reads): consecutive position pieces of [lo, hi)
    with the reads that overlap them."""
    if stepper not in _cabi.STEPPERS:
        raise ValueError('unknown stepper: %r' % (stepper,))     # pysam raises ValueError too
    name = stat.name
    bam = bamio.open_bam(alignmentfile)
    fa = bamio.open_fasta(fafile)
    ctx = _cabi.get_context()
    need_qual = bool(ctx.lib.pss_pileup_needs_qual(_cabi.PILEUP_STATS.index(name), int(min_baseq)))
    needs_ref = bool(ctx.lib.pss_pileup_needs_ref(_cabi.PILEUP_STATS.index(name)))
    if chrom is not None:
        contigs = [chrom]
    else:
        contigs = list(bam.references)
        if not pad:
            # iter_pileup_default hands chrom=None to normalise_coords, which drops start / end (opt.pyx:2111-2112);
            # iter_pileup_padded forwards the caller's start / end to every contig (opt.pyx:1927-1932)
            start = end = None
    for c in contigs:
        if pad:
            assert c is not None, 'chromosome is None'
        s, e = normalise_coords(bam, c, start, end, one_based)
        _check_region(s, e)

        def reference():
            refseq = fa.contig(c) if (fa is not None and needs_ref) else np.zeros(0, np.uint8)
            return refseq, dict(stepper=stepper, max_depth=max_depth, min_mapq=min_mapq, min_baseq=min_baseq, no_del=no_del,
                                no_dup=no_dup, window_size=stat.window_size, window_offset=stat.window_offset,
                                ref=refseq if len(refseq) else None, ref_start=0, ref_len=len(refseq), contig_len=len(refseq))

        if not need_qual and e - s > _table_chunk() and bam.can_stream(c):
            # decode -> upload -> compute pipeline: the first pieces are being inflated while the FASTA is read
            state = {}
            it = _streamed_pieces(bam, c, s, e, truncate, pad, max_depth, state)
            first = next(it, None)
            if first is None:
                continue
            refseq, pkw = reference()
            lo = state['lo']
            hi = max(e, bam.lengths[bam.get_tid(c)]) + (1 << 16)          # upper bound; reads do not leave the contig
            import itertools
            yield c, s, e, lo, hi, itertools.chain([first], it), refseq, pkw, need_qual
            continue
        batch = bam.region_batch(c, s, e)
        n = batch.n_reads
        if n == 0 and not pad:
            continue
        if truncate:
            lo, hi = s, e
        elif n:
            if 'span' not in batch.cache:
                batch.cache['span'] = (int(batch.pos[0]), int(batch.endpos.max()))
            lo, hi = batch.cache['span']
            if pad:
                lo, hi = min(lo, s), max(hi, e)
        else:
            lo, hi = s, e
        if hi - lo <= 0:
            continue
        refseq, pkw = reference()
        yield c, s, e, lo, hi, _position_chunks(batch, lo, hi, need_qual, max_depth, stepper), refseq, pkw, need_qual


def pileup_tables(stat, alignmentfile, fafile=None, pad=False, chrom=None, start=None, end=None,
                  one_based=False, truncate=False, stepper='all', max_depth=8000, min_mapq=0, min_baseq=0,
                  no_del=False, no_dup=False):
    """Yield ``(chrom, pos, ref, fields)`` per contig: ``pos`` int64 (already one-based if
    asked), ``ref`` an ``S1`` column or None, ``fields`` a dict name -> int32 column.

    Row selection follows ``iter_pileup_default`` / ``iter_pileup_padded_chrom``
    (``opt.pyx:1846-1863``, ``:1936-1984``): a row for every column htslib builds, plus -- with
    ``pad`` -- every position of ``[start, end)``; with ``truncate`` only ``[start, end)``.
    """
    name = stat.name
    names = config.value_fields(name)
    for c, s, e, lo, hi, pieces, refseq, pkw, need_qual in _pileup_plan(
            stat, alignmentfile, fafile, pad, chrom, start, end, one_based, truncate, stepper, max_depth, min_mapq,
            min_baseq, no_del, no_dup):
        ctx = _cabi.get_context()

        def run_piece(piece, device):
            """One position piece on one GPU -> (pos, ref column, fields) or None when it has no rows."""
            c_lo, c_hi, reads = piece
            pctx = _cabi.get_context(*device)
            cn = c_hi - c_lo
            params = pctx.pileup_params(name, start=c_lo, end=c_hi, **pkw)
            out = bamio.alloc_big((len(names), cn), np.int32)
            rowmask = np.empty(cn, dtype=np.uint8)
            pctx.pileup_host(reads, params, out, rowmask)
            rows = rowmask.view(bool)
            if pad:
                rows = rows.copy()
                rows[max(min(s, c_hi) - c_lo, 0):max(min(e, c_hi) - c_lo, 0)] = True      # [s, e) within the piece
            if rows.all():
                pos = np.arange(c_lo, c_hi, dtype=np.int64)
                fields = {f: out[j] for j, f in enumerate(names)}
            else:
                sel = np.flatnonzero(rows)
                if len(sel) == 0:
                    return None
                pos = sel.astype(np.int64) + c_lo
                taken = bamio.take_cols(out, sel)
                fields = {f: taken[j] for j, f in enumerate(names)}
            ref = _ref_column(refseq, pos) if name in config.ref_column_types else None
            if 'gc' in fields and len(pos) and fields['gc'].min() < 0:
                # the reference stores gc = -1 for an empty window into a u1 column, which numpy
                # refuses
                raise OverflowError('Python integer -1 out of bounds for uint8')
            if one_based:
                pos = pos + 1
            return pos, ref, fields

        for res in _run_pieces(run_piece, pieces, ctx.device):
            if res is not None:
                yield (c,) + res


def pileup_rows(stat, dtype, alignmentfile, fafile=None, pad=False, chrom=None, start=None, end=None,
                one_based=False, truncate=False, stepper='all', max_depth=8000, min_mapq=0, min_baseq=0,
                no_del=False, no_dup=False):
    """What ``load_*`` needs, without the column tables in between: a list of record arrays of ``dtype`` (one per
    contig, in order), packed by the engine's host pipeline straight from its result chunks
    (``pss_pileup_host_rows``) -- same rows and values as ``pileup_tables``.  None when the request has to take
    the column path (a field type the row sink does not write, several GPUs, no address space for the records).
    """
    name = stat.name
    names = config.value_fields(name)
    ctx = _cabi.get_context()
    devs = _devices(ctx.device)
    if ctx.row_layout(name, dtype, names, chrom='') is None:
        return None
    parts = []
    for c, s, e, lo, hi, pieces, refseq, pkw, need_qual in _pileup_plan(
            stat, alignmentfile, fafile, pad, chrom, start, end, one_based, truncate, stepper, max_depth, min_mapq,
            min_baseq, no_del, no_dup):
        layout = ctx.row_layout(name, dtype, names, chrom=c, pos_add=1 if one_based else 0,
                                pad=(s, e) if pad else None, check_field='gc')
        if layout is None:
            return None
        cap = hi - lo
        try:
            buf = bamio.alloc_rows(cap, dtype)
        except (MemoryError, OSError, ValueError):
            return None
        n = 0
        if len(devs) == 1:
            for c_lo, c_hi, reads in pieces:
                params = ctx.pileup_params(name, start=c_lo, end=c_hi, **pkw)
                m, neg = ctx.pileup_host_rows(reads, params, layout, buf, n, cap - n)
                if neg:
                    # the reference stores gc = -1 for an empty window into a u1 column, which numpy
                    # refuses
                    raise OverflowError('Python integer -1 out of bounds for uint8')
                n += m
        else:
            # several GPUs (PSS_GPUS): the position pieces go to the GPUs in turn, piece [c_lo, c_hi) packing its records
            # at row c_lo - lo (it cannot emit more rows than positions); pieces with fewer rows than positions leave
            # gaps, closed afterwards by moving the later pieces down (nothing moves on a fully covered contig)
            def run_rows(piece, device):
                c_lo, c_hi, reads = piece
                pctx = _cabi.get_context(*device)
                params = pctx.pileup_params(name, start=c_lo, end=c_hi, **pkw)
                m, neg = pctx.pileup_host_rows(reads, params, layout, buf, c_lo - lo, c_hi - c_lo)
                return c_lo - lo, m, neg
            for r0, m, neg in _run_pieces(run_rows, pieces, ctx.device):
                if neg:
                    raise OverflowError('Python integer -1 out of bounds for uint8')
                if m and r0 != n:
                    buf[n:n + m] = buf[r0:r0 + m]         # (numpy copies overlapping ranges correctly)
                n += m
        if n == 0:
            continue
        if n * 2 < cap and n * dtype.itemsize < (1 << 30):
            part = np.empty(n, dtype=dtype)          # sparse contig: do not keep the position-sized mapping alive
            part[...] = buf[:n]
        else:
            part = buf[:n]
        parts.append(part)
    return parts


def _devices(default):
    """(device, slot) pairs the pieces of a contig are spread over.  PSS_GPUS = a count, 'all', or an explicit
    comma-separated device list; a device listed k times runs k shards at once, each on its own context (the tests
    use ``0,0`` to run the sharded paths on a one-GPU box).  Default: one, the context's."""
    import os
    v = os.environ.get('PSS_GPUS', '').strip().lower()
    if v in ('', '0', '1'):
        return [(default, 0)]
    if ',' in v:
        ids = [int(x) for x in v.split(',') if x.strip() != '']
        return [(d, ids[:i].count(d)) for i, d in enumerate(ids)] if len(ids) > 1 else [(default, 0)]
    import torch
    n = torch.cuda.device_count()
    k = n if v == 'all' else min(int(v), n)
    return [(d, 0) for d in range(k)] if k > 1 else [(default, 0)]


def _run_pieces(fn, pieces, default_device):
    """Results of fn(piece, (device, slot)) in piece order.  With several GPUs (PSS_GPUS) piece i goes to GPU i mod n
    from a thread per GPU -- the interval sharding of SURVEY.md section 8e behind the public API: contiguous position
    intervals, the halo of reads that start before them, no collective, pieces concatenated in order -- and at most
    n pieces are in flight, so a context never runs two at once."""
    devs = _devices(default_device)
    if len(devs) == 1:
        for piece in pieces:
            yield fn(piece, devs[0])
        return
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(devs)) as ex:
        pending = deque()
        for i, piece in enumerate(pieces):
            if len(pending) == len(devs):
                yield pending.popleft().result()
            pending.append(ex.submit(fn, piece, devs[i % len(devs)]))
        while pending:
            yield pending.popleft().result()


def _pair_links(batch, stepper, max_depth):
    """``pair_first`` of the whole batch (``pss_pairs_host``): for every read the index of the earlier mate its base
    qualities are rewritten against (htslib's overlap hash under pysam's ``ignore_overlaps=True``, call site
    ``opt.pyx:1853``), -1 for the rest -- from the read metadata alone.  Cached per batch."""
    key = ('pairs', stepper, max_depth or 8000)
    if key not in batch.cache:
        ctx = _cabi.get_context()
        batch.cache[key] = ctx.pairs_host(_reads_struct(batch, True), stepper, max_depth)
    return batch.cache[key]


def _position_chunks(batch, lo, hi, need_qual, max_depth, stepper='all', pair_links=None):
    """Cut the positions [lo, hi) of one contig batch into pieces of PSS_TABLE_CHUNK (default 16 Mi) positions,
    each with the sub-batch of reads that overlap it (a zero-copy view: the reads are start sorted), so a whole
    chromosome streams through bounded device (and result) memory.  Yields (c_lo, c_hi, struct pss_reads).

    The ``max_depth`` admission rule is order dependent: a read before a cut can only be refused because of
    earlier reads the next piece would not see.  So the range is only cut when no read of the batch can be
    refused at all (fewer than ``max_depth`` reads start within one maximal read span of any read).

    Requests that read base qualities: the mate-overlap rewrite pairs reads through a QNAME-keyed table, so which
    reads pair depends on every earlier read of the name -- a piece that misses the first read of a three-read chain
    would pair the other two.  The pairs are therefore computed ONCE for the whole batch from its metadata
    (``_pair_links``; ``pair_links`` replaces that call in tests without a GPU) and handed to every piece as its
    ``name_prev`` links: ``name_prev[b] = a`` for a pair (a, b) inside the piece, -1 everywhere else -- chains of
    at most two reads, which the engine pairs exactly as the whole batch did.  A pair with only one read in the
    piece rewrites qualities at positions both reads cover, i.e. outside the piece."""
    from . import shard
    n = batch.n_reads
    chunk = _table_chunk()
    if hi - lo <= chunk or n == 0:
        yield lo, hi, _reads_struct(batch, need_qual)
        return
    # read i has max_depth - 1 earlier reads within maxspan of its start iff pos[i] - pos[i - (max_depth - 1)] <=
    # maxspan (the reads are start sorted): one subtraction over the batch instead of a binary search per read
    maxspan, refusable = _refusable(batch, max_depth)
    if refusable:
        yield lo, hi, _reads_struct(batch, need_qual)
        return
    batch.padded()
    links = None
    if need_qual:
        links = (pair_links or _pair_links)(batch, stepper, max_depth)
    for c_lo in range(lo, hi, chunk):
        c_hi = min(hi, c_lo + chunk)
        i0, i1, _ = shard.halo_range(batch.pos, batch.endpos, c_lo, c_hi, maxspan=maxspan, mask=False)
        d = dict(n_reads=i1 - i0, tid=batch.tid)
        for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off'):
            d[k] = getattr(batch, k)[i0:]
        d['cigar'], d['seq4'] = batch.cigar, batch.seq4
        keep = None
        if need_qual:
            keep = links[i0:i1]
            keep = np.where(keep >= i0, keep - i0, -1).astype(np.int32)      # the piece's own indices
            d['qual'], d['qual_off'], d['name_prev'] = batch.qual, batch.qual_off[i0:], keep
            d['n_qual_bytes'] = int(batch.qual_off[i1] - batch.qual_off[i0])
        else:
            d['qual'] = d['qual_off'] = d['name_prev'] = None
        r = _cabi.make_reads_struct(d, n_reads=i1 - i0)
        r._keep = (batch, keep)     # the struct holds raw addresses
        yield c_lo, c_hi, r


_ITER_BLOCK = 1 << 16


def iter_pileup(stat, alignmentfile, fafile, pad, **kwargs):
    """Generator of one dict per genome position -- the reference's ``opt.iter_pileup``
    (``opt.pyx:1827-1843``), same keyword arguments."""
    for chrom, pos, ref, fields in pileup_tables(stat, alignmentfile, fafile, pad, **kwargs):
        names = list(fields)
        for o in range(0, len(pos), _ITER_BLOCK):     # python objects for a block of rows at a time (bounded memory)
            sl = slice(o, o + _ITER_BLOCK)
            cols = [fields[f][sl].tolist() for f in names]
            refs = [r.decode('ascii') for r in ref[sl].tolist()] if ref is not None else None
            for i, p in enumerate(pos[sl].tolist()):
                rec = {f: col[i] for f, col in zip(names, cols)}
                if refs is not None:
                    rec['ref'] = refs[i]
                rec['chrom'] = chrom
                rec['pos'] = p
                yield rec


# ------------------------------------------------------------------------ binned tables

def binned_tables(stat, alignmentfile, fafile=None, chrom=None, window_size=300, window_offset=None,
                  start=None, end=None, one_based=False, min_mapq=0, no_dup=False):
    """Yield ``(chrom, pos, None, fields)`` per contig for a binned statistic
    (``opt.pyx:1997-2102``)."""
    name = stat.name
    bam = bamio.open_bam(alignmentfile)
    fa = bamio.open_fasta(fafile)
    ctx = _cabi.get_context()
    names = config.value_fields_binned(name)
    if window_offset is None:
        window_offset = window_size / 2          # opt.pyx:2006-2007: a float ...
    window_offset = int(window_offset)           # ... truncated by the typed C int argument
    window_size = int(window_size)
    need_ref = name in ('coverage', 'coverage_ext')
    if need_ref and fa is None:
        raise ValueError('reference sequence is required; please provide fafile argument')
    if chrom is not None:
        contigs = [chrom]
    else:
        contigs = list(bam.references)       # iter_binned forwards start / end to every contig (opt.pyx:2009-2015)
    for c in contigs:
        assert c is not None, 'chromosome is None'
        s, e = normalise_coords(bam, c, start, end, one_based)
        _check_region(s, e)
        batch = bam.region_batch(c, s, e)
        refseq = fa.contig(c) if need_ref else None
        nb = int(ctx.lib.pss_binned_nbins(int(s), int(e), window_size))
        out = np.empty((len(names), nb), dtype=np.int32)
        devs = _devices(ctx.device)
        if len(devs) > 1 and nb >= 2 * len(devs) and batch.n_reads:
            _binned_across_gpus(devs, name, batch, s, e, window_size, nb, min_mapq, no_dup, refseq, out)
        else:
            params = ctx.binned_params(name, s, e, window_size, min_mapq=min_mapq, no_dup=no_dup,
                                       ref=refseq, ref_start=0, ref_len=len(refseq) if need_ref else 0)
            ctx.binned_host(_reads_struct(batch, False), params, out)
        if need_ref and nb and out[names.index('gc')].min() < 0:
            # gc_content() of an empty window (beyond the FASTA): opt.pyx:1538 divides by zero
            raise ZeroDivisionError('float division')
        pos = s + np.arange(nb, dtype=np.int64) * window_size + window_offset
        if one_based:
            pos = pos + 1
        yield c, pos, None, {f: out[j] for j, f in enumerate(names)}


def _binned_across_gpus(devs, name, batch, s, e, window_size, nb, min_mapq, no_dup, refseq, out):
    """The windows of [s, e) spread over several GPUs (PSS_GPUS): contiguous window ranges (``shard.plan_windows``),
    each with the reads that START in it (``shard.window_read_range``: the roll-over rule of ``opt.pyx:2050``; no halo,
    the statistics are keyed by read start), a thread per GPU, columns written side by side into ``out`` -- no
    collective.  A shard is the same request with start / end moved to its window range; its
    first / last window collect exactly what the global first / last window collect because only its own reads
    are passed."""
    from concurrent.futures import ThreadPoolExecutor
    from . import shard
    batch.padded()
    need_ref = refseq is not None

    def run(job):
        dev, (k_lo, k_hi) = job
        if k_hi <= k_lo:
            return
        i0, i1 = shard.window_read_range(batch.pos, s, window_size, k_lo, k_hi, nb)
        s_k = s + k_lo * window_size
        e_k = e if k_hi >= nb else s + k_hi * window_size
        d = dict(n_reads=i1 - i0, tid=batch.tid)
        for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off'):
            d[k] = getattr(batch, k)[i0:]
        d['cigar'], d['seq4'] = batch.cigar, batch.seq4
        d['qual'] = d['qual_off'] = d['name_prev'] = None
        pctx = _cabi.get_context(*dev)
        params = pctx.binned_params(name, s_k, e_k, window_size, min_mapq=min_mapq, no_dup=no_dup, ref=refseq,
                                    ref_start=0, ref_len=len(refseq) if need_ref else 0)
        assert int(params.n_bins) == k_hi - k_lo
        part = np.empty((out.shape[0], k_hi - k_lo), dtype=np.int32)
        pctx.binned_host(_cabi.make_reads_struct(d, n_reads=i1 - i0), params, part)
        out[:, k_lo:k_hi] = part

    with ThreadPoolExecutor(len(devs)) as ex:
        list(ex.map(run, zip(devs, shard.plan_windows(s, e, window_size, len(devs)))))


def iter_binned(stat, alignmentfile, fafile, chrom, window_size=300, window_offset=None, **kwargs):
    """Generator of one dict per window -- the reference's ``opt.iter_binned``
    (``opt.pyx:1997-2023``)."""
    for c, pos, _, fields in binned_tables(stat, alignmentfile, fafile, chrom, window_size, window_offset,
                                           **kwargs):
        names = list(fields)
        cols = [fields[f].tolist() for f in names]
        for i, p in enumerate(pos.tolist()):
            rec = {f: col[i] for f, col in zip(names, cols)}
            rec['chrom'] = c
            rec['pos'] = p
            yield rec
