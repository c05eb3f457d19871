#!/usr/bin/env python
"""Secondary bench (BASELINE.json configs[4]): the whole-genome binned sweep -- coverage_gc (pileup +
sliding GC), mapq_binned, tlen_binned, coverage_binned with window_size=300 -- over the synthetic 30x
workload, one 250 Mb interval per GPU, sharded by contiguous interval / window range with no collective
(torchrun for N > 1, like bench.py).  Prints one JSON line; not the driver's headline metric.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--interval-mb', type=float, default=250.0)
    ap.add_argument('--window-size', type=int, default=300)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pysamstats_b200 import _cabi, config, shard, synth
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG=VERSION on the GPU
        # boxes); stdout carries the one JSON line, so file descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    W = args.window_size
    per = int(args.interval_mb * 1e6) // (shard.TILE * W) * (shard.TILE * W) if False else int(args.interval_mb * 1e6) // 76800 * 76800
    genome = per * world                         # 76800 = lcm(256, 300): shard cuts are tile AND window aligned
    lo, hi = rank * per, (rank + 1) * per
    batch, (ref, ref_lo) = synth.make_interval_batch(lo, hi, genome, device=dev, with_qual=False)
    n_reads = int(batch['n_reads'])
    ctx = _cabi.Context(local_rank)
    s = torch.cuda.current_stream().cuda_stream
    # binned statistics are keyed by read START: the shard's reads are those starting in its windows
    nbins_all = genome // W
    i0, i1 = shard.window_read_range(batch['pos'].cpu().numpy(), 0, W, lo // W, hi // W, nbins_all)
    sub = {k: (batch[k][i0:] if k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off') else batch[k])
           for k in _cabi.READ_FIELDS if batch.get(k) is not None}
    sub['tid'] = 0
    rs_bin = _cabi.make_reads_struct(sub, n_reads=i1 - i0)
    rs = _cabi.make_reads_struct(batch)
    outs = {}
    binned = ('coverage', 'mapq', 'tlen')
    bp = {t: ctx.binned_params(t, lo, hi, W, ref=ref, ref_start=ref_lo, ref_len=int(ref.numel())) for t in binned}
    for t in binned:
        outs[t] = torch.empty((len(config.value_fields_binned(t)), int(bp[t].n_bins)), dtype=torch.int32, device=dev)
    gp = ctx.pileup_params('coverage_gc', start=lo, end=hi, ref=ref, ref_start=ref_lo, ref_len=int(ref.numel()),
                           contig_len=genome, window_size=W)
    gout = torch.empty((3, hi - lo), dtype=torch.int32, device=dev)
    gmask = torch.empty(hi - lo, dtype=torch.uint8, device=dev)

    def step():
        ctx.pileup(rs, gp, gout, gmask, stream=s)
        for t in binned:
            ctx.binned(rs_bin, bp[t], outs[t], stream=s)

    for _ in range(args.warmup):
        step()
    ctx.synchronize(s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # algorithmic bytes of one sweep on this rank: coverage_gc = 3 int32 columns + row mask +
    # reference byte per position and 29 + 4 n_ops bytes per read; every binned statistic 15 bytes per read
    # (pos, flag, mapq, tlen, mtid) + its int32 columns per window; coverage_binned also reads the reference
    nwin = per // W
    alg = (hi - lo) * (4 * 3 + 1 + 1) + n_reads * 29 + 4 * int(batch['n_cigar_words'])
    alg += sum(n_reads * 15 + 4 * len(config.value_fields_binned(t)) * nwin for t in binned) + (hi - lo)
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peak = float(json.load(open(peaks_path))['hbm_gbs']) if os.path.exists(peaks_path) else 6650.0
    if rank == 0:
        achieved = alg / (ms * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                    'algorithmic_bytes_per_step': alg, 'what': 'the whole sweep (tile kernel for coverage + k_gc + 3 x (k_binned_acc, '
                    'k_binned_fin) + k_binned_gc), not one kernel'}
        print(json.dumps({'roofline': roofline, 'metric': 'genome positions/sec (coverage_gc + coverage/mapq/tlen binned sweep, window 300)',
                          'value': genome / (ms * 1e-3), 'unit': 'positions/s', 'reads_per_s': n_reads * world / (ms * 1e-3),
                          'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'scaling': 'weak',
                          'windows_per_gpu': per // W, 'reads_per_gpu': n_reads, 'data': 'synthetic',
                          'checksum': {t: int(outs[t][0].sum().item()) for t in binned}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
