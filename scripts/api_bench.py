#!/usr/bin/env python
"""End-to-end timing of the PUBLIC python API on a synthetic BAM written to disk: BAM decode -> engine ->
structured array, for load_variation_strand / load_coverage_ext / load_coverage_binned (what a pysamstats
user calls).  Prints seconds per stage and rows / s."""
import argparse, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pysamstats_b200 as pysamstats
from pysamstats_b200 import bamio, synth

ap = argparse.ArgumentParser()
ap.add_argument('--mb', type=float, default=16.0)
ap.add_argument('--no-csv', action='store_true', help='skip the write_csv timings (large runs)')
args = ap.parse_args()
G = int(args.mb * 1e6)
dev = 'cuda' if torch.cuda.is_available() else 'cpu'
batch, (ref, ref_lo) = synth.make_interval_batch(0, G, G, depth=30, device=dev, with_qual=True)
nb = synth.batch_to_numpy(batch)
with tempfile.TemporaryDirectory() as tmp:
    bam, fa = os.path.join(tmp, 's.bam'), os.path.join(tmp, 's.fa')
    bamio.write_bam(bam, ['chrS'], [G], nb, level=1)
    bamio.index_bam(bam)
    seq = ref.cpu().numpy().tobytes()
    with open(fa, 'w') as f:
        f.write('>chrS\n')
        for i in range(0, len(seq), 60):
            f.write(seq[i:i + 60].decode() + '\n')
    print('BAM %.0f MB, %d reads, %.0f Mb' % (os.path.getsize(bam) / 1e6, nb['n_reads'], G / 1e6))
    for rep in range(2):
        bamio._bam_cache.clear()
        t0 = time.time()
        b = bamio.open_bam(bam).load()
        t1 = time.time()
        a = pysamstats.load_variation_strand(bam, fafile=fa)
        t2 = time.time()
        c = pysamstats.load_coverage_ext(bam)
        t3 = time.time()
        d = pysamstats.load_coverage_binned(bam, fafile=fa)
        t4 = time.time()
        r = pysamstats.load_variation(bam, fafile=fa, chrom='chrS', start=G // 2, end=G // 2 + 100000)
        t5 = time.time()
        print('rep %d: decode %.2f s | load_variation_strand %.2f s (%.1f M rows/s, %d rows, itemsize %d) | '
              'load_coverage_ext %.2f s | load_coverage_binned %.3f s | 100 kb region %.3f s'
              % (rep, t1 - t0, t2 - t1, len(a) / (t2 - t1) / 1e6, len(a), a.dtype.itemsize, t3 - t2, t4 - t3, t5 - t4))
    from pysamstats_b200.io import write_csv
    for st in (() if args.no_csv else ('coverage', 'variation_strand')):
        with open(os.path.join(tmp, 'out.tsv'), 'w') as f:
            t0 = time.time()
            write_csv(st, f, bam, fafile=fa if st == 'variation_strand' else None)
            dt = time.time() - t0
            size = f.tell()
        print('write_csv %s (tsv to a file): %.2f s, %.2f M rows/s, %.0f MB/s' % (st, dt, len(a) / dt / 1e6, size / dt / 1e6))
    bamio._bam_cache.clear()
    t0 = time.time()
    r = pysamstats.load_variation(bam, fafile=fa, chrom='chrS', start=G // 2, end=G // 2 + 100000)
    print('cold 100 kb region through the BAI index: %.3f s, %d rows' % (time.time() - t0, len(r)))
    if os.environ.get('API_PROFILE'):
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        a = pysamstats.load_variation_strand(bam, fafile=fa)
        pr.disable()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
