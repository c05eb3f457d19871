"""Writers: ``write_csv`` / ``write_hdf5`` with the reference's signatures and output
(``pysamstats/io.py:14-244``).

The reference pushes one dict per row through ``csv.writer`` / a list of tuples through
PyTables; with the engine producing whole column tables that would be the bottleneck, so rows
are formatted block-wise from the columns (numpy string ops) and written in large chunks.
The bytes written are those ``csv.writer`` would produce for the same rows.
"""
import csv
import io as _io
import sys
import time

import numpy as np

from . import binned, config, pileup, util

_BLOCK = 1 << 18      # rows formatted per block


def _tables(stats_type, alignmentfile, **kwargs):
    """Column tables of a statistic: generator of (chrom, pos, ref, {field: column})."""
    if stats_type.endswith('_binned'):
        t = stats_type[:-len('_binned')]
        binned._make_stat(t)                                   # ValueError for an unknown type
        return binned._binned_tables(t, alignmentfile=alignmentfile, **kwargs)
    return pileup._pileup_tables(stats_type, alignmentfile=alignmentfile, **kwargs)


def _default_fields(stats_type):
    try:
        return [t[0] for t in getattr(config, 'dtype_' + stats_type)]
    except AttributeError:
        raise AttributeError("module 'pysamstats' has no attribute 'stat_%s'" % stats_type)


def format_rows(fields, chrom, pos, ref, cols, dialect):
    """Text of the rows of one column table, exactly as ``csv.writer(dialect).writerow`` would
    write them."""
    d = csv.get_dialect(dialect)
    special = (d.delimiter, d.quotechar or '"', '\r', '\n')
    n = len(pos)
    if n == 0:
        return ''
    chrom_text = chrom
    if any(c in chrom for c in special):                         # let csv do the quoting
        buf = _io.StringIO()
        csv.writer(buf, dialect=dialect).writerow([chrom])
        chrom_text = buf.getvalue()[:-len(d.lineterminator)]
    if d.quoting != csv.QUOTE_MINIMAL or d.escapechar or not d.doublequote:
        # any other dialect: csv.writer itself, row by row (the tsv / csv formats of the CLI never come here)
        buf = _io.StringIO()
        columns = []
        for f in fields:
            if f == 'chrom':
                columns.append([chrom] * n)
            elif f == 'pos':
                columns.append(np.asarray(pos).tolist())
            elif f == 'ref':
                columns.append([r.decode('ascii') for r in np.asarray(ref).tolist()])
            else:
                columns.append(np.asarray(cols[f]).tolist())
        csv.writer(buf, dialect=dialect).writerows(zip(*columns))
        return buf.getvalue()
    # native formatter (libpss_host, all host threads) whenever no cell of the block needs quoting
    if d.quoting == csv.QUOTE_MINIMAL and ('ref' not in fields or ref is None or
                                           not np.isin(np.asarray(ref), [c.encode() for c in special]).any()):
        native = []
        for f in fields:
            if f == 'chrom':
                native.append(chrom_text)
            elif f == 'pos':
                native.append(np.ascontiguousarray(pos))
            elif f == 'ref':
                native.append(np.asarray(ref))
            else:
                native.append(np.asarray(cols[f]))
        from . import bamio
        text = bamio.format_rows_native(native, d.delimiter, d.lineterminator)
        if text is not None:
            return text
    parts = []
    for f in fields:
        if f == 'chrom':
            text = chrom_text
            parts.append(np.full(n, text, dtype='U%d' % max(1, len(text))))
        elif f == 'pos':
            parts.append(np.asarray(pos).astype('U'))
        elif f == 'ref':
            r = np.asarray(ref).astype('U1')
            bad = np.isin(r, special)
            if bad.any():
                r = r.astype('U4')
                q = d.quotechar or '"'
                r[bad] = np.char.add(np.char.add(q, np.char.replace(r[bad], q, q + q)), q)
            parts.append(r)
        else:
            parts.append(np.asarray(cols[f]).astype('U'))
    line = parts[0]
    for p in parts[1:]:
        line = np.char.add(np.char.add(line, d.delimiter), p)
    return d.lineterminator.join(line.tolist()) + d.lineterminator


def write_csv(stats_type, outfile, alignmentfile, fields=None, dialect='excel-tab', write_header=True,
              progress=None, **kwargs):
    """Write statistics output to a CSV file.

    Parameters
    ----------
    stats_type : string
        Statistics type, one of 'coverage', 'coverage_ext', ..., 'coverage_binned', ...
    outfile : file-like
        Output file to write to.
    alignmentfile : path or alignment file object
        Input BAM file.
    fields : list of strings
        List of field names to output (all by default).
    dialect : string
        CSV dialect.
    write_header : bool
        If True write a header row.
    progress : int
        Log progress to stderr every N rows.
    **kwargs
        Passed through to the statistics function.
    """
    if fields is None:
        fields = _default_fields(stats_type)
    tables = _tables(stats_type, alignmentfile, **kwargs)
    if write_header:
        csv.writer(outfile, dialect=dialect).writerow(fields)
    counter = 0
    before_all = before = time.time()
    next_report = progress
    for chrom, pos, ref, cols in tables:
        missing = [f for f in fields if f not in ('chrom', 'pos', 'ref') and f not in cols]
        if missing or ('ref' in fields and ref is None):
            raise KeyError((missing or ['ref'])[0])               # itemgetter(*fields) in the reference
        for o in range(0, len(pos), _BLOCK):
            sl = slice(o, o + _BLOCK)
            outfile.write(format_rows(fields, chrom, pos[sl], None if ref is None else ref[sl],
                                      {f: c[sl] for f, c in cols.items()}, dialect))
            counter += len(pos[sl])
            while progress and counter >= next_report:
                after = time.time()
                elapsed, batch_elapsed = max(after - before_all, 1e-9), max(after - before, 1e-9)
                print('[pysamstats] %s rows in %.2fs (%d rows/s); batch in %.2fs (%d rows/s)'
                      % (next_report, elapsed, next_report / elapsed, batch_elapsed, progress / batch_elapsed),
                      file=sys.stderr)
                before = after
                next_report += progress
    if progress is not None:
        elapsed_all = max(time.time() - before_all, 1e-9)
        print('[pysamstats] %s rows in %.2fs (%d rows/s)' % (counter, elapsed_all, counter / elapsed_all),
              file=sys.stderr)


def write_hdf5(stats_type, outfile, alignmentfile, fields=None, progress=None, hdf5_group='/',
               hdf5_dataset='data', hdf5_complevel=1, hdf5_complib='zlib', hdf5_shuffle=True,
               hdf5_fletcher32=False, hdf5_chunksize=2**20, dtype=None, **kwargs):
    """Write statistics output to an HDF5 file.  Requires PyTables (same dataset layout, filters
    and chunk shape as the reference: ``io.py:91-244``)."""
    import tables
    default_dtype = dict(getattr(config, 'dtype_' + stats_type))
    if fields is None:
        fields = _default_fields(stats_type)
    default_dtype['chrom'] = 'S{0}'.format(util.determine_max_seqid(alignmentfile))
    if dtype is not None:
        default_dtype.update(dict(dtype))
    dt = np.dtype(default_dtype[fields[0]] if len(fields) == 1 else [(f, default_dtype[f]) for f in fields])
    recs = _tables(stats_type, alignmentfile, **kwargs)
    h5file = None
    try:
        h5file = tables.open_file(outfile, mode='a')
        chunklen = int(hdf5_chunksize / dt.itemsize)
        try:
            h5file.remove_node(hdf5_group, hdf5_dataset)
        except tables.NoSuchNodeError:
            pass
        h5table = h5file.create_table(hdf5_group, hdf5_dataset, dt, title=stats_type,
                                      filters=tables.Filters(complevel=hdf5_complevel, complib=hdf5_complib,
                                                             shuffle=hdf5_shuffle, fletcher32=hdf5_fletcher32),
                                      createparents=True, chunkshape=(chunklen,))
        counter = counter_before = 0
        before_all = before = time.time()
        for chrom, pos, ref, cols in recs:
            for o in range(0, len(pos), chunklen):
                m = len(pos[o:o + chunklen])
                chunk = np.zeros(m, dtype=dt)
                for f in fields:
                    dst = chunk if len(fields) == 1 else chunk[f]
                    if f == 'chrom':
                        dst[...] = chrom.encode('ascii')
                    elif f == 'pos':
                        dst[...] = pos[o:o + m]
                    elif f == 'ref':
                        dst[...] = ref[o:o + m]
                    else:
                        dst[...] = cols[f][o:o + m]
                h5table.append(chunk)
                h5table.flush()
                counter += m
                if progress and (counter % progress) < chunklen:
                    after = time.time()
                    elapsed, batch_elapsed = max(after - before_all, 1e-9), max(after - before, 1e-9)
                    batch_size = counter - counter_before
                    print('[pysamstats] %s rows in %.2fs (%d rows/s); last %s rows in %.2fs (%d rows/s)'
                          % (counter, elapsed, counter / elapsed, batch_size, batch_elapsed,
                             batch_size / batch_elapsed), file=sys.stderr)
                    before, counter_before = after, counter
        if progress:
            elapsed_all = max(time.time() - before_all, 1e-9)
            print('[pysamstats] %s rows in %.2fs (%d rows/s)' % (counter, elapsed_all, counter / elapsed_all),
                  file=sys.stderr)
    finally:
        if h5file is not None:
            h5file.close()
