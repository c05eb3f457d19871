#!/usr/bin/env python
"""Device-resident timing of every pileup / binned statistic on the synthetic 30x workload
(one interval of --mb Mb): ms per call, positions/s, achieved algorithmic GB/s."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pysamstats_b200 import _cabi, config, synth

ap = argparse.ArgumentParser()
ap.add_argument('--mb', type=float, default=50.0)
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--generic', action='store_true')
ap.add_argument('--only', default='')
ap.add_argument('--min-baseq', type=int, default=0)
ap.add_argument('--no-dup', action='store_true')
ap.add_argument('--no-del', action='store_true')
args = ap.parse_args()
dev = torch.device('cuda', 0)
L = int(args.mb * 1e6) // 512 * 512
batch, (ref, ref_lo) = synth.make_interval_batch(0, L, L, depth=30, device=dev, with_qual=True)
n = int(batch['n_reads'])
ctx = _cabi.Context(0)
rs = _cabi.make_reads_struct(batch)
s = torch.cuda.current_stream().cuda_stream
rows = []
for stat in config.pileup_types:
    if args.only and stat not in args.only.split(','):
        continue
    nf = len(config.value_fields(stat))
    out = torch.empty((nf, L), dtype=torch.int32, device=dev)
    mask = torch.empty(L, dtype=torch.uint8, device=dev)
    withref = stat in config.stats_types_withref
    need_qual = stat.startswith('baseq') or args.min_baseq > 0
    p = ctx.pileup_params(stat, start=0, end=L, ref=ref if withref else None, ref_start=ref_lo,
                          ref_len=int(ref.numel()) if withref else 0, contig_len=L,
                          flags=_cabi.PSS_FLAG_GENERIC_KERNEL if args.generic else 0, min_baseq=args.min_baseq, no_dup=args.no_dup,
                          no_del=args.no_del)
    ctx.prepare(rs, with_qual=need_qual, no_dup=args.no_dup, stream=s)
    for _ in range(int(os.environ.get('SWEEP_WARMUP', 2))):
        ctx.pileup_prepared(p, out, mask, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        ctx.pileup_prepared(p, out, mask, stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    # bytes a statistic has to touch: packed bases when it compares them with the reference, qualities when it sums them;
    # a min_baseq threshold alone costs the one mark bit per base k_lowq leaves (the marks are made once per batch)
    reads_q = stat.startswith('baseq')
    seqb = int(batch['n_seq_bytes']) if (withref and stat != 'coverage_gc') else 0
    qb = int(batch['n_qual_bytes']) if reads_q else (int(batch['n_qual_bytes']) // 8 if args.min_baseq > 0 else 0)
    alg = L * (4 * nf + 2) + n * 29 + 4 * int(batch['n_cigar_words']) + seqb + qb
    rows.append((stat, ms, L / ms / 1e6, alg / ms / 1e6))
    del out, mask
print('%-22s %9s %12s %10s %9s' % ('statistic (pileup)', 'ms', 'Gpos/s', 'GB/s', '% of 6557'))
for stat, ms, gp, gb in rows:
    print('%-22s %9.3f %12.3f %10.1f %8.1f%%' % (stat, ms, gp, gb, 100 * gb / 6557.1))
print()
print('%-22s %9s %12s' % ('statistic (binned)', 'ms', 'Greads/s'))
for stat in config.binned_types:
    if args.only and stat + '_binned' not in args.only.split(','):
        continue
    nf = len(config.value_fields_binned(stat))
    p = ctx.binned_params(stat, 0, L, 300, ref=ref, ref_start=ref_lo, ref_len=int(ref.numel()))
    out = torch.empty((nf, int(p.n_bins)), dtype=torch.int32, device=dev)
    for _ in range(2):
        ctx.binned(rs, p, out, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        ctx.binned(rs, p, out, stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print('%-22s %9.3f %12.3f' % (stat + '_binned', ms, n / ms / 1e6))
# several binned statistics from one pass over the records (pss_binned_multi), with the library's per-kernel events
for sub in (('coverage', 'mapq', 'tlen'), tuple(config.binned_types)):
    if args.only and 'multi_binned' not in args.only.split(','):
        continue
    ps = [ctx.binned_params(t, 0, L, 300, ref=ref, ref_start=ref_lo, ref_len=int(ref.numel())) for t in sub]
    outs = [torch.empty((len(config.value_fields_binned(t)), int(ps[0].n_bins)), dtype=torch.int32, device=dev) for t in sub]
    for _ in range(2):
        ctx.binned_multi(rs, ps, outs, stream=s)
    torch.cuda.synchronize()
    ctx.set_kernel_timing(True)
    ctx.pileup_kernel_ms(-1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        ctx.binned_multi(rs, ps, outs, stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    k = {name: ctx.pileup_kernel_ms(tag)[0] for name, tag in (('acc', _cabi.KTAG_BINNED_ACC), ('fin', _cabi.KTAG_BINNED_FIN), ('gc', _cabi.KTAG_BINNED_GC))}
    ctx.set_kernel_timing(False)
    print('%-22s %9.3f %12.3f   k_binned_acc %.3f ms, k_binned_fin %.3f ms, k_binned_gc %.3f ms' % ('+'.join(sub), ms, n / ms / 1e6, k['acc'], k['fin'], k['gc']))
