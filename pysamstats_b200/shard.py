"""Interval sharding of the pileup / binned path across the GPUs of one box.

The path shards naturally: positions are independent given the
reads covering them, windows are independent given the read starts.  So there is no
data-path collective: each rank (one process per GPU) takes a contiguous interval of the
region plus the *halo* of reads that start before it but overlap it, writes only its own
positions, and the host concatenates the pieces in rank order.  ``torch.distributed`` is
used by callers only for the barrier / timing reduction and to gather small results.

What crosses a shard boundary besides the halo reads:

* depth-cap admission (htslib ``max_depth``) is order dependent, but only inside runs where
  the live depth reaches ``max_depth``; shards must not be cut inside such a run.
  ``safe_cut`` moves a boundary to the nearest position where the live depth is below the
  cap (ordinary WGS never saturates, so boundaries stay where they are).
* the mate-overlap quality rewrite pairs reads by QNAME in push order.  Both mates of a pair overlap
  every position the rewrite touches, but WHICH reads of a name pair up depends on every earlier read
  of that name (a three-read chain pairs differently when its first read is missing), so a shard that
  only holds the reads overlapping its interval can pair the rest of a chain differently.  Requests
  that read base qualities (``pss_pileup_needs_qual``: the baseq* statistics, any ``min_baseq > 0``)
  are therefore NOT cut by these helpers: ``plan_intervals(..., need_qual=True)`` returns one shard
  holding the whole range.  (``opt._position_chunks`` does cut such requests: it has the whole contig's
  reads at hand, computes the pairs once from their metadata -- ``pss_pairs_host`` -- and hands every piece
  its pairs as two-read name chains; a rank that only ever sees its own shard cannot do that.)
"""
import numpy as np

TILE = 512   # pss_tile_positions() (PSS_TTILE, the scatter kernel's tile): interior shard boundaries are tile
             # aligned, so no shard starts with a partial tile (tests/test_host_logic.py checks it against the library)


def plan_intervals(start, end, n_shards, align=TILE, need_qual=False):
    """Cut [start, end) into n_shards contiguous intervals of (almost) equal length whose
    inner boundaries are multiples of ``align`` relative to ``start``.
    Returns a list of (lo, hi); empty intervals (lo == hi) are possible for tiny regions.
    With ``need_qual`` (the request reads base qualities) the range is not cut: the first shard
    gets all of it, the others are empty (see the module docstring)."""
    if n_shards < 1:
        raise ValueError('n_shards must be >= 1')
    if need_qual:
        return [(int(start), max(int(start), int(end)))] + [(max(int(start), int(end)),) * 2] * (n_shards - 1)
    span = max(0, int(end) - int(start))
    cuts = [int(start)]
    for k in range(1, n_shards):
        c = (span * k) // n_shards
        c = (c + align // 2) // align * align
        cuts.append(int(start) + min(span, c))
    cuts.append(int(start) + span)
    for k in range(1, len(cuts)):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return [(cuts[k], cuts[k + 1]) for k in range(n_shards)]


def plan_windows(start, end, window_size, n_shards):
    """Binned statistics: cut the windows of [start, end) (bin k = [start + k W, start + (k+1) W))
    into n_shards contiguous window ranges.  Returns [(k_lo, k_hi)]; the bin index always
    uses the GLOBAL start so edges agree between shards."""
    span = int(end) - int(start)
    n_bins = max(1, -(-span // int(window_size))) if span > 0 else 1
    return [((n_bins * k) // n_shards, (n_bins * (k + 1)) // n_shards) for k in range(n_shards)]


def _key(pos, v):
    """Search key in the dtype of ``pos`` (a python int would make numpy convert the whole
    array on every call)."""
    info = np.iinfo(pos.dtype)
    return pos.dtype.type(min(max(int(v), info.min), info.max))


def halo_range(pos, endpos, lo, hi, maxspan=None, mask=True):
    """Index range [i0, i1) of a coordinate-sorted batch that contains every read overlapping
    [lo, hi) (pos < hi and endpos > lo -- htslib's region rule), and (``mask=True``) the
    boolean mask of the reads inside that range that really overlap.  The range is contiguous
    because reads are start-sorted; reads inside it that end before lo are harmless to the
    kernels (they cover no output position) but are reported so callers can count exactly.
    ``maxspan`` (max over the batch of endpos - pos, if known) bounds the backward scan."""
    pos = np.asarray(pos)
    endpos = np.asarray(endpos)
    i1 = int(np.searchsorted(pos, _key(pos, hi), side='left'))
    s0 = 0 if maxspan is None else int(np.searchsorted(pos, _key(pos, lo - int(maxspan) + 1), side='left'))
    s0 = min(s0, i1)
    # every read starting at or after lo overlaps; only the ones before it need a look
    k0 = min(int(np.searchsorted(pos, _key(pos, lo), side='left')), i1)
    nz = np.flatnonzero(endpos[s0:k0] > lo)
    i0 = s0 + int(nz[0]) if len(nz) else k0
    if not mask:
        return i0, i1, None
    ov = np.ones(i1 - i0, dtype=bool)
    ov[:k0 - i0] = endpos[i0:k0] > lo
    return i0, i1, ov


def window_read_range(pos, start, window_size, k_lo, k_hi, n_bins):
    """Index range of the reads that fall in windows [k_lo, k_hi) under the roll-over rule of
    iter_binned_chrom (opt.pyx:2050: strict '>', so window k >= 1 owns start + k W < pos <=
    start + (k+1) W, window 0 everything up to start + W, the last window everything after)."""
    pos = np.asarray(pos)
    i0 = 0 if k_lo == 0 else int(np.searchsorted(pos, _key(pos, start + k_lo * window_size), side='right'))
    i1 = len(pos) if k_hi >= n_bins else int(np.searchsorted(pos, _key(pos, start + k_hi * window_size), side='right'))
    return i0, max(i0, i1)


def live_depth_at(pos, endpos, p):
    """Number of reads covering position p (pos <= p < endpos) of a sorted batch."""
    pos = np.asarray(pos)
    endpos = np.asarray(endpos)
    i1 = int(np.searchsorted(pos, _key(pos, p), side='right'))
    return int(np.count_nonzero(endpos[:i1] > p))


def safe_cut(pos, endpos, cut, max_depth, align=TILE, limit=1 << 20):
    """Move a shard boundary forward (tile by tile) until the live depth there is below
    max_depth, so no saturated depth-cap run spans it."""
    c = int(cut)
    for _ in range(limit):
        if live_depth_at(pos, endpos, c) < max_depth:
            return c
        c += align
    raise ValueError('no unsaturated cut point found near %d' % cut)


def concat_shards(parts):
    """Concatenate per-shard (pos, columns...) pieces in rank order (the only 'merge')."""
    parts = [p for p in parts if p is not None and len(p[0])]
    if not parts:
        return None
    return tuple(np.concatenate([p[i] for p in parts], axis=-1) for i in range(len(parts[0])))
