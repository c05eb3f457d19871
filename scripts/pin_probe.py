#!/usr/bin/env python
"""Probe: how much of a load_* call is the upload of the (pageable) read arrays?  Times load_variation_strand /
load_coverage_ext on a synthetic BAM before and after cudaHostRegister of the decoded batch arrays."""
import argparse, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pysamstats_b200 as pysamstats
from pysamstats_b200 import bamio, synth

ap = argparse.ArgumentParser()
ap.add_argument('--mb', type=float, default=16.0)
args = ap.parse_args()
G = int(args.mb * 1e6)
batch, (ref, ref_lo) = synth.make_interval_batch(0, G, G, depth=30, device='cuda', with_qual=True)
nb = synth.batch_to_numpy(batch)
with tempfile.TemporaryDirectory() as tmp:
    bam, fa = os.path.join(tmp, 's.bam'), os.path.join(tmp, 's.fa')
    bamio.write_bam(bam, ['chrS'], [G], nb, level=1)
    seq = ref.cpu().numpy().tobytes()
    with open(fa, 'w') as f:
        f.write('>chrS\n')
        for i in range(0, len(seq), 60):
            f.write(seq[i:i + 60].decode() + '\n')
    b = bamio.open_bam(bam).load()

    def run(tag):
        for rep in range(3):
            t0 = time.time()
            a = pysamstats.load_variation_strand(bam, fafile=fa)
            t1 = time.time()
            c = pysamstats.load_coverage_ext(bam)
            t2 = time.time()
            print('%s rep %d: load_variation_strand %.3f s (%.1f M rows/s) | load_coverage_ext %.3f s' % (
                tag, rep, t1 - t0, len(a) / (t1 - t0) / 1e6, t2 - t1), flush=True)
            del a, c
    run('pageable')
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    c = pysamstats.load_coverage_ext(bam)
    pr.disable()
    del c
    pstats.Stats(pr).sort_stats('tottime').print_stats(12)
    rt = torch.cuda.cudart()
    A = b._load_all_any()
    t0 = time.time()
    nbytes = 0
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'cigar', 'seq_off', 'seq4'):
        x = getattr(A, k)
        r = rt.cudaHostRegister(x.ctypes.data, x.nbytes, 0)
        nbytes += x.nbytes
    print('cudaHostRegister of %.0f MB: %.3f s' % (nbytes / 1e6, time.time() - t0))
    run('pinned  ')
