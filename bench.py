#!/usr/bin/env python
"""bench.py -- throughput of the pileup path on the synthetic 30x WGS workload.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]; SURVEY.md section 8d "S1"): per GPU one 250 Mb interval of
a synthetic genome covered 30x by 150 bp paired reads (50 M reads per interval), statistics
``variation_strand`` + ``coverage_ext`` for every position.  A *step* is one pass of both
statistics over the rank's interval.  With N ranks the genome is N x 250 Mb, cut by
``pysamstats_b200.shard.plan_intervals`` into contiguous intervals, each rank holding its
interval's reads plus the halo of reads that start before it (weak scaling, no data-path
collective; torch.distributed only provides the barrier and the max-over-ranks reduction).

Numbers on the JSON line (besides the headline workload the line carries four more objects, each measured in
the same run: ``pileup_stats`` -- at N = 1, EVERY pileup statistic device-resident on one 50 Mb interval, default filters
and min_baseq = 20, each with the fraction of the HBM roofline its algorithmic bytes reach; ``binned`` -- BASELINE.json configs[4], the whole-genome binned sweep coverage_gc + coverage_binned +
mapq_binned + tlen_binned with window_size 300 on the same reads, with its own value / e2e / per-kernel rooflines /
cpu_baseline; ``strong_scaling`` -- the ONE 250 Mb contig of configs[2] cut into N shards of 250 / N Mb with read
halos, device-resident and end to end; ``e2e_api`` -- the public python API from a real BAM + BAI file, decode
included):
  value     genome positions / s, whole job, read batch resident in HBM, results left in HBM
            (C ABI pss_prepare + pss_pileup_prepared per statistic, CUDA events on the launch
            stream, max over ranks)
  e2e       same metric through pss_pileup_host_multi: pinned HOST read arrays in, HOST result
            columns out, every copy inside the timed region (wall clock around the calls)
  roofline  dominant kernel (k_tile<VAR, variation_strand>): algorithmic bytes / its mean device
            time (events recorded by the library around the kernel, see pss_pileup_kernel_ms)
  cpu_baseline   the CPU restatement of the reference algorithm (oracle/, "port") on a bounded
            sample of the same workload, one process per contig chunk on all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'genome positions/sec (stat_variation_strand + stat_coverage_ext, synthetic 30x WGS)'
UNIT = 'positions/s'
STATS = ('variation_strand', 'coverage_ext')


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--interval-mb', type=float, default=250.0, help='positions per GPU, in Mb')
    ap.add_argument('--depth', type=int, default=30)
    ap.add_argument('--read-len', type=int, default=150)
    ap.add_argument('--e2e-chunk-mb', type=float, default=25.0, help='positions per host call (size of the host result buffers)')
    ap.add_argument('--e2e-stream', action='store_true', help='e2e through pss_pileup_host_stream (uploads in pieces, overlapped)')
    ap.add_argument('--e2e-inner-mb', type=float, default=12.5, help='--e2e-stream: positions per upload piece inside one host call')
    ap.add_argument('--cpu-chunk-mb', type=float, default=1.0, help='CPU arm: positions per process, in Mb')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-binned', action='store_true', help='skip the binned sweep (BASELINE.json configs[4])')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling arm (N > 1)')
    ap.add_argument('--no-e2e-api', action='store_true', help='skip the from-file public API measurement')
    ap.add_argument('--no-stats', action='store_true', help='skip the per-statistic device timings (pileup_stats)')
    ap.add_argument('--stats-mb', type=float, default=50.0, help='pileup_stats: interval length, in Mb')
    ap.add_argument('--api-mb', type=float, default=250.0, help='e2e_api: contig length of the BAM written for it, in Mb')
    ap.add_argument('--window-size', type=int, default=300)
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smmax) if smmax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------- reference arm

BINNED_METRIC = 'genome positions/sec (coverage_gc + coverage_binned + mapq_binned + tlen_binned sweep, window_size 300, synthetic 30x WGS)'
BINNED_SWEEP = ('coverage', 'mapq', 'tlen')
ALIGN = 38400      # lcm(512, 300): interval cuts are tile aligned (scatter kernel) AND window aligned (binned sweep)


def cpu_binned(args, steps, warmup):
    """The binned sweep on the CPU restatement of the reference (oracle port), one process per contig chunk."""
    from oracle import cpu_arm
    cores = cpu_arm.usable_cores()
    chunk = int(args.cpu_chunk_mb * 1e6) // 300 * 300
    res = cpu_arm.run(cores=cores, chunk=chunk, steps=steps, warmup=warmup, depth=args.depth, read_len=args.read_len,
                      workload='binned')
    ts = res['step_seconds']
    mean = sum(ts) / len(ts)
    return {'value': res['positions_per_step'] / mean, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'reads_per_s': res['reads_per_step'] / mean, 'ms_per_step': mean * 1e3,
            'sample': '%d contig chunks of %.2f Mb (one per process) of the same synthetic workload per step: stat_coverage_gc + '
                      'coverage_binned + mapq_binned + tlen_binned, window_size 300' % (cores, chunk / 1e6)}


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on the host cores.  The reference's own
    Cython (pysamstats/opt.pyx) needs pysam/htslib, which are not available offline, so this
    is the oracle port (oracle/pileup_oracle.c), one process per contig chunk."""
    if rank != 0:
        return
    from oracle import cpu_arm
    cores = cpu_arm.usable_cores()
    chunk = int(args.cpu_chunk_mb * 1e6)
    res = cpu_arm.run(cores=cores, chunk=chunk, steps=args.steps, warmup=args.warmup, depth=args.depth,
                      read_len=args.read_len)
    ts = res['step_seconds']
    mean = sum(ts) / len(ts)
    value = res['positions_per_step'] / mean
    sample = ('%d contig chunks of %.2f Mb (one per process, %d processes) of the synthetic 30x / %d bp workload per '
              'step' % (cores, chunk / 1e6, cores, args.read_len))
    binned = None
    if not args.no_binned:
        cb = cpu_binned(args, max(1, min(args.steps, 3)), min(args.warmup, 1))
        binned = {'metric': BINNED_METRIC, 'value': cb['value'], 'unit': UNIT, 'ms_per_step': cb['ms_per_step'],
                  'reads_per_s': cb['reads_per_s'], 'cpu_baseline': cb,
                  'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': mean * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
        'config': workload_config(args, args.gpus),
        'reads_per_s': res['reads_per_step'] / mean,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'binned': binned,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {'workload': 'synthetic 30x WGS-shaped reads (150 bp paired, 250 Mb contig interval per GPU), '
                        'stat_variation_strand + stat_coverage_ext',
            'interval_positions_per_gpu': int(args.interval_mb * 1e6) // ALIGN * ALIGN, 'depth': args.depth,
            'read_len': args.read_len, 'stats': list(STATS), 'min_baseq': 0, 'stepper': 'all', 'max_depth': 8000,
            'sharding': 'contiguous intervals + read halo, %d shard(s), no collective' % n,
            'batch': 'records with base qualities (the same arrays the CPU arm gets); these statistics do not read them',
            'l2': 'inputs (> 5 GB of read records) and outputs (68 GB) exceed the 126 MB L2; no flush needed'}


def peak_gbs():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def roof(kernel, alg_bytes, ms, n, peak, what):
    """One per-kernel roofline entry: algorithmic bytes of a launch / its mean CUDA-event time."""
    ach = alg_bytes / (ms * 1e-3) / 1e9 if ms and ms > 0 else None
    return {'kernel': kernel, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': (ach / peak) if ach else None,
            'kernel_ms': ms, 'launches_timed': n, 'algorithmic_bytes_per_launch': int(alg_bytes), 'bytes': what}


# ---------------------------------------------------------------------------- main arm

def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from pysamstats_b200 import _cabi, config, shard, synth

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    def sum_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return float(x)

    peak, peak_src = peak_gbs()
    ctx = _cabi.Context(local_rank)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    nf = {s: int(ctx.lib.pss_pileup_nfields(_cabi.PILEUP_STATS.index(s))) for s in STATS}
    W = args.window_size

    def device_arm(batch, ref, ref_lo, lo, hi, contig_len, steps, warmup, timing=True):
        """prepare + variation_strand + coverage_ext over [lo, hi), batch resident in HBM, results left in HBM.
        Returns (ms per step: CUDA events on the launch stream, launches per run, kernel ms by tag, checksums, rows)."""
        npos = hi - lo
        reads = _cabi.make_reads_struct(batch)
        outs = {s: torch.empty((nf[s], npos), dtype=torch.int32, device=dev) for s in STATS}
        masks = {s: torch.empty(npos, dtype=torch.uint8, device=dev) for s in STATS}
        params = {s: ctx.pileup_params(s, start=lo, end=hi, ref=ref if s == 'variation_strand' else None,
                                       ref_start=ref_lo, ref_len=int(ref.numel()) if s == 'variation_strand' else 0,
                                       contig_len=contig_len) for s in STATS}

        def step():
            # one preparation of the batch (stepper, depth cap, read filters), then both statistics
            ctx.prepare(reads, stepper='all', max_depth=8000, stream=sptr)
            for s in STATS:
                ctx.pileup_prepared(params[s], outs[s], masks[s], stream=sptr)

        for _ in range(warmup):
            step()
        ctx.synchronize(sptr)
        ctx.set_kernel_timing(timing)
        ctx.pileup_kernel_ms(-1)
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        launches = ctx.launch_count() - l0
        kms = {s: ctx.pileup_kernel_ms(s) for s in STATS}
        kms['prep'] = ctx.pileup_kernel_ms(_cabi.KTAG_PREP)
        ctx.set_kernel_timing(False)
        ctx.pileup_kernel_ms(-1)
        ctx.synchronize(sptr)
        # checksum of checksums (a size-independent sanity check)
        chk = {s: int(outs[s][0].sum(dtype=torch.int64).item()) for s in STATS}
        rows = int(masks['variation_strand'].sum(dtype=torch.int64).item())
        del outs, masks
        return ms, launches, kms, chk, rows

    def host_copy(batch, ref):
        """Pinned host copies of the record arrays these statistics read (the decoder's output in a real run)."""
        hb = {}
        for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'cigar', 'seq_off', 'seq4'):
            hb[k] = batch[k].cpu().pin_memory().numpy()
        hb['flag'] = hb['flag'].view(np.uint16)
        hb['cigar'] = hb['cigar'].view(np.uint32)
        hb['endpos'] = batch['endpos'].cpu().numpy()
        hb['maxspan'] = int((hb['endpos'] - hb['pos']).max()) if len(hb['pos']) else 1
        return hb, ref.cpu().pin_memory().numpy()

    def host_view(hb, c_lo, c_hi):
        """struct pss_reads over the sub-range of the pinned batch that overlaps [c_lo, c_hi) (zero copy)."""
        i0, i1, _ = shard.halo_range(hb['pos'], hb['endpos'], c_lo, c_hi, maxspan=hb['maxspan'], mask=False)
        view = dict(n_reads=i1 - i0, tid=0)
        for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off'):
            view[k] = hb[k][i0:]
        view['cigar'], view['seq4'] = hb['cigar'], hb['seq4']
        return _cabi.make_reads_struct(view, n_reads=i1 - i0)

    def e2e_arm(hb, ref_h, ref_lo, lo, hi, contig_len, steps, warmup, expect):
        """The same statistics through pss_pileup_host_multi: pinned HOST arrays in, HOST int32 columns out, every copy
        inside the timed region (wall clock).  Returns (seconds per step, h2d bytes, d2h bytes per step, chunk)."""
        npos = hi - lo
        chunk = max(shard.TILE, int(args.e2e_chunk_mb * 1e6) // shard.TILE * shard.TILE)
        chunk = min(chunk, npos)
        inner = max(shard.TILE, int(args.e2e_inner_mb * 1e6) // shard.TILE * shard.TILE)
        oh = {s: torch.empty((nf[s], chunk), dtype=torch.int32).pin_memory().numpy() for s in STATS}
        mh = {s: torch.empty(chunk, dtype=torch.uint8).pin_memory().numpy() for s in STATS}
        checks = {}

        def e2e_step(check=False):
            for c_lo in range(lo, hi, chunk):
                c_hi = min(hi, c_lo + chunk)
                rs = host_view(hb, c_lo, c_hi)
                n = c_hi - c_lo
                r0 = c_lo - ref_lo
                ps, os_, ms = [], [], []
                for s in STATS:
                    withref = s == 'variation_strand'
                    ps.append(ctx.pileup_params(s, start=c_lo, end=c_hi, ref=ref_h[r0:] if withref else None,
                                                ref_start=c_lo, ref_len=(c_hi - c_lo) if withref else 0,
                                                contig_len=contig_len))
                    os_.append(oh[s] if n == chunk else np.ndarray((nf[s], n), dtype=np.int32, buffer=oh[s]))
                    ms.append(mh[s])
                # one call per position chunk: the reads go up once, both statistics come back
                if args.e2e_stream:
                    ctx.pileup_host_stream(rs, ps, os_, ms, hb['maxspan'], inner)
                else:
                    ctx.pileup_host_multi(rs, ps, os_, ms)
                if check:
                    for s, o in zip(STATS, os_):
                        checks[s] = checks.get(s, 0) + int(o[0].sum(dtype=np.int64))

        for _ in range(max(1, min(warmup, 2))):
            checks.clear()
            e2e_step(check=True)
        assert checks == expect, ('host path disagrees with the device path', checks, expect)
        h0, d0 = ctx.transfer_bytes()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / steps
        barrier()
        h1, d1 = ctx.transfer_bytes()
        return sec, (h1 - h0) / steps, (d1 - d0) / steps, chunk, inner

    # ---- the rank's shard of the genome ------------------------------------------------
    per = int(args.interval_mb * 1e6) // ALIGN * ALIGN
    genome = per * world
    lo, hi = shard.plan_intervals(0, genome, world, align=ALIGN)[rank]
    npos = hi - lo
    t_gen = time.time()
    # (the batch reaches one position beyond the interval: a read that starts exactly at `hi` belongs to the rank's last
    # window under the strict roll-over of opt.pyx:2050; it covers no position of [lo, hi))
    batch, (ref, ref_lo) = synth.make_interval_batch(lo, min(hi + 1, genome), genome, depth=args.depth, read_len=args.read_len,
                                                     device=dev, with_qual=True)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    n_reads = int(batch['n_reads'])

    # ---- device-resident throughput -----------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, launches, kms, chk, rows = device_arm(batch, ref, ref_lo, lo, hi, genome, args.steps, args.warmup)
    dev_ms = max_over_ranks(dev_ms)
    total_pos = sum_over_ranks(npos)
    total_reads = sum_over_ranks(n_reads)
    value = total_pos / (dev_ms * 1e-3)

    # ---- roofline of the dominant kernel (rank 0's launch) --------------------------------
    n_cig = int(batch['n_cigar_words'])
    n_seq = int(batch['n_seq_bytes'])
    alg_bytes = npos * (4 * nf['variation_strand'] + 1 + 1) + n_reads * 29 + 4 * n_cig + n_seq
    # DRAM traffic of the same kernel on the same workload from one `ncu --set full` capture
    traffic = None
    for tname in ('r02_traffic.json', 'r01_traffic.json'):
        tpath = os.path.join(ROOT, 'profiles', tname)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if int(tj.get('interval_positions', -1)) == per:
                traffic = tj['traffic_bytes_per_launch']
                break
    k_ms, k_n = kms['variation_strand']
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    roofline = {'bound': 'hbm', 'kernel': 'k_tile<KIND_VAR, PSS_VARIATION_STRAND>', 'achieved': achieved,
                'peak': peak, 'unit': 'GB/s', 'frac': (achieved / peak) if achieved else None, 'traffic': traffic,
                'peak_source': peak_src, 'kernel_ms': k_ms, 'kernel_launches_timed': k_n,
                'algorithmic_bytes_per_launch': alg_bytes,
                'kernel_share_of_step': (k_ms / dev_ms) if k_ms > 0 else None,
                'coverage_ext_kernel_ms': kms['coverage_ext'][0], 'prepare_ms': kms['prep'][0]}

    # ---- the binned sweep (BASELINE.json configs[4]) on the same reads, device resident ----------------------------
    binned = None
    if not args.no_binned:
        # binned statistics are keyed by read START: the rank's reads are those starting in its windows (bin edges use
        # the genome's start, so they agree between ranks; no halo)
        nbins_all = genome // W if genome % W == 0 else genome // W + 1
        pos_h = batch['pos'].cpu().numpy()
        i0, i1 = shard.window_read_range(pos_h, 0, W, lo // W, -(-hi // W), nbins_all)
        n_bin_reads = i1 - i0
        sub = {k: (batch[k][i0:] if k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off') else batch[k])
               for k in _cabi.READ_FIELDS if batch.get(k) is not None}
        sub['tid'] = 0
        rs_bin = _cabi.make_reads_struct(sub, n_reads=n_bin_reads)
        rs_all = _cabi.make_reads_struct(batch)
        bps = [ctx.binned_params(t, lo, hi, W, ref=ref, ref_start=ref_lo, ref_len=int(ref.numel())) for t in BINNED_SWEEP]
        nwin = int(bps[0].n_bins)
        bouts = [torch.empty((len(config.value_fields_binned(t)), nwin), dtype=torch.int32, device=dev) for t in BINNED_SWEEP]
        gp = ctx.pileup_params('coverage_gc', start=lo, end=hi, ref=ref, ref_start=ref_lo, ref_len=int(ref.numel()),
                               contig_len=genome, window_size=W)
        gout = torch.empty((3, npos), dtype=torch.int32, device=dev)
        gmask = torch.empty(npos, dtype=torch.uint8, device=dev)

        def bstep():
            ctx.pileup(rs_all, gp, gout, gmask, stream=sptr)           # prepare + coverage (tile kernel) + sliding GC%
            ctx.binned_multi(rs_bin, bps, bouts, stream=sptr)          # the three binned statistics, records read once

        for _ in range(args.warmup):
            bstep()
        ctx.synchronize(sptr)
        ctx.set_kernel_timing(True)
        ctx.pileup_kernel_ms(-1)
        bl0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            bstep()
        e1.record(stream)
        barrier()
        b_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        b_launches = ctx.launch_count() - bl0
        tags = {'k_tile<KIND_COV, PSS_COVERAGE_GC>': 'coverage_gc', 'k_gc': _cabi.KTAG_GC, 'k_prep (+ depth cap)': _cabi.KTAG_PREP,
                'k_binned_acc<coverage | mapq | tlen>': _cabi.KTAG_BINNED_ACC, 'k_binned_fin': _cabi.KTAG_BINNED_FIN,
                'k_binned_gc': _cabi.KTAG_BINNED_GC}
        bk = {name: ctx.pileup_kernel_ms(tag) for name, tag in tags.items()}
        ctx.set_kernel_timing(False)
        ctx.pileup_kernel_ms(-1)
        ctx.synchronize(sptr)
        nfb = sum(len(config.value_fields_binned(t)) for t in BINNED_SWEEP)
        # algorithmic bytes per launch: what the kernel must read and write once
        rl = [
            roof('k_tile<KIND_COV, PSS_COVERAGE_GC>', npos * (4 * 2 + 1) + n_reads * 29 + 4 * n_cig, *bk['k_tile<KIND_COV, PSS_COVERAGE_GC>'], peak,
                 '2 int32 columns + row mask per position; 29 B record + CIGAR words per read'),
            roof('k_prep (+ depth cap)', n_reads * (29 + 8 + 32) + 4 * n_cig, *bk['k_prep (+ depth cap)'], peak,
                 '29 B record + 8 B sequence offset + CIGAR words read, 32 B RInfo written per read'),
            roof('k_gc', npos * (1 + 4), *bk['k_gc'], peak, '1 reference byte read + 1 int32 written per position'),
            roof('k_binned_acc<coverage | mapq | tlen>', n_bin_reads * 11 + nwin * (3 * 4 + 5 * 8), *bk['k_binned_acc<coverage | mapq | tlen>'], peak,
                 'pos 4 + flag 2 + mapq 1 + tlen 4 B per read, read ONCE for the three statistics; 3 counts + 5 sums per window'),
            roof('k_binned_gc', npos + 4 * nwin, *bk['k_binned_gc'], peak, '1 reference byte per position, 1 int32 per window'),
            roof('k_binned_fin', nwin * (3 * 4 + 5 * 8 + 4 * (nfb - 1)), *bk['k_binned_fin'], peak, 'accumulators read, int32 columns written per window'),
        ]
        alg_sweep = sum(r['algorithmic_bytes_per_launch'] for r in rl)
        bchk = {t: int(o[1 if t == 'coverage' else 0].sum(dtype=torch.int64).item()) for t, o in zip(BINNED_SWEEP, bouts)}
        bchk['coverage_gc'] = int(gout[1].sum(dtype=torch.int64).item())       # fields: gc, reads_all, reads_pp
        bchk['gc'] = int(gout[0].sum(dtype=torch.int64).item())
        assert bchk['coverage'] == bchk['mapq'] == bchk['tlen'] == sum_mapped_reads(batch, i0, i1), bchk    # every mapped read in exactly one window
        assert bchk['coverage_gc'] == chk['coverage_ext'], (bchk, chk)                                        # same coverage as the pileup arm
        binned = {'metric': BINNED_METRIC, 'value': total_pos / (b_ms * 1e-3), 'unit': UNIT, 'ms_per_step': b_ms,
                  'reads_per_s': total_reads / (b_ms * 1e-3), 'windows_per_s': sum_over_ranks(nwin) / (b_ms * 1e-3),
                  'scaling': 'weak', 'n_gpus': world, 'gpu_launches': int(b_launches),
                  'config': {'workload': 'the same synthetic 30x reads; stat_coverage_gc (per position) + coverage_binned + mapq_binned + '
                                         'tlen_binned (per window), window_size %d, default offset' % W,
                             'sharding': 'contiguous window ranges (global bin edges), reads by start window, %d shard(s), no collective' % world,
                             'windows_per_gpu': nwin, 'binned_reads_per_gpu': n_bin_reads},
                  'roofline': dict(rl[3], peak_source=peak_src, kernel_share_of_step=(rl[3]['kernel_ms'] / b_ms) if rl[3]['kernel_ms'] else None),
                  'kernels': rl,
                  'sweep_algorithmic_bytes': int(alg_sweep), 'sweep_frac_of_peak': alg_sweep / (b_ms * 1e-3) / 1e9 / peak,
                  'checksum_rank0': bchk}
        del gout, gmask, bouts

    # ---- strong scaling: ONE 250 Mb contig cut into `world` shards (BASELINE.json configs[2] as written) ------------------
    strong = None
    if not args.no_strong:
        if world == 1:
            strong = {'contig_positions': per, 'shards': 1, 'value': value, 'unit': UNIT, 'ms_per_step': dev_ms,
                      'note': 'one shard = the headline workload'}
        else:
            s_lo, s_hi = shard.plan_intervals(0, per, world, align=shard.TILE)[rank]
            sb, (sref, sref_lo) = synth.make_interval_batch(s_lo, s_hi, per, depth=args.depth, read_len=args.read_len,
                                                            device=dev, with_qual=True)
            torch.cuda.synchronize()
            s_ms, s_launches, s_kms, s_chk, s_rows = device_arm(sb, sref, sref_lo, s_lo, s_hi, per, max(args.steps, 20), args.warmup)
            s_ms = max_over_ranks(s_ms)
            strong = {'contig_positions': per, 'shards': world, 'shard_positions_rank0': s_hi - s_lo,
                      'shard_reads_rank0': int(sb['n_reads']), 'value': per / (s_ms * 1e-3), 'unit': UNIT, 'ms_per_step': s_ms,
                      'launches_per_step': s_launches / max(args.steps, 20),
                      'kernel_ms_rank0': {'variation_strand': s_kms['variation_strand'][0], 'coverage_ext': s_kms['coverage_ext'][0],
                                          'prepare': s_kms['prep'][0]},
                      'scaling': 'strong', 'sharding': 'plan_intervals(0, contig, N) cuts aligned to %d + read halo; no collective' % shard.TILE,
                      'checksum_sum_over_ranks': {k: int(sum_over_ranks(v)) for k, v in s_chk.items()}}
            if not args.no_e2e:
                shb, sref_h = host_copy(sb, sref)
                sec, h2d, d2h, _, _ = e2e_arm(shb, sref_h, sref_lo, s_lo, s_hi, per, max(2, min(args.steps, 5)), 1, s_chk)
                sec = max_over_ranks(sec)
                strong['e2e'] = {'value': per / sec, 'unit': UNIT, 'ms_per_step': sec * 1e3,
                                 'h2d_bytes_per_step': int(sum_over_ranks(h2d)), 'd2h_bytes_per_step': int(sum_over_ranks(d2h))}
                del shb, sref_h
            del sb, sref

    # ---- end to end through the host-buffer C ABI ------------------------------------------
    e2e = None
    hb = ref_h = None
    if not args.no_e2e:
        torch.cuda.empty_cache()
        hb, ref_h = host_copy(batch, ref)
        e2e_s, h2d, d2h, chunk, inner = e2e_arm(hb, ref_h, ref_lo, lo, hi, genome, args.steps, args.warmup, chk)
        e2e_s = max_over_ranks(e2e_s)
        e2e = {'value': total_pos / e2e_s, 'unit': UNIT, 'ms_per_step': e2e_s * 1e3,
               'h2d_bytes_per_step': int(sum_over_ranks(h2d)), 'd2h_bytes_per_step': int(sum_over_ranks(d2h)),
               'api': ('pss_pileup_host_stream (pinned host read arrays in, host int32 columns of both statistics out), '
                       '%d calls of %d positions per step, reads uploaded in pieces of %d positions'
                       % (-(-npos // chunk), chunk, inner)) if args.e2e_stream else
                      ('pss_pileup_host_multi (pinned host read arrays in, host int32 columns of both statistics out), '
                       '%d position chunks of %d per step' % (-(-npos // chunk), chunk))}
        # ---- the floor under any host path: the API contract hands int32 columns to the caller, so the HOST must write
        # 4 bytes per field and position whatever the engine does; measured write bandwidth of this box (all ranks
        # probing at once, as they share the memory system in the timed run), non-temporal int32 stores
        try:
            from pysamstats_b200 import bamio
            hl = bamio.host_lib()
            threads = max(1, (os.cpu_count() or 1) // world)
            barrier()
            wbw = hl.pss_host_bandwidth(256 << 20, threads, 3, 0)
            wid = hl.pss_host_bandwidth(256 << 20, threads, 3, 1)
            barrier()
            wbw_all, wid_all = sum_over_ranks(wbw), sum_over_ranks(wid)
            int32_bytes = 4 * sum(nf.values()) * total_pos
            e2e['host_floor'] = {'int32_result_bytes_per_step': int(int32_bytes), 'host_write_gbs': wbw_all,
                                 'host_widen_u8_to_i32_gbs_written': wid_all, 'threads': threads * world,
                                 'floor_ms': int32_bytes / (wbw_all * 1e9) * 1e3,
                                 'frac': (int32_bytes / (wbw_all * 1e9)) / e2e_s,
                                 'what': 'seconds the host needs just to WRITE the int32 columns the API contract returns (bytes / measured '
                                         'non-temporal store bandwidth, all ranks probing together) / e2e seconds per step'}
        except Exception as ex:   # noqa: BLE001 -- informative only
            e2e['host_floor'] = {'error': repr(ex)}
        # ---- the binned sweep end to end: pss_pileup_host (coverage_gc) per position chunk + pss_binned_host_multi
        if binned is not None:
            gchunk = chunk
            go = torch.empty((3, gchunk), dtype=torch.int32).pin_memory().numpy()
            gm = torch.empty(gchunk, dtype=torch.uint8).pin_memory().numpy()
            bo = [torch.empty((len(config.value_fields_binned(t)), nwin), dtype=torch.int32).pin_memory().numpy() for t in BINNED_SWEEP]
            hview = dict(n_reads=n_bin_reads, tid=0)
            for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off'):
                hview[k] = hb[k][i0:]
            hview['cigar'], hview['seq4'] = hb['cigar'], hb['seq4']
            rs_bin_h = _cabi.make_reads_struct(hview, n_reads=n_bin_reads)
            bps_h = [ctx.binned_params(t, lo, hi, W, ref=ref_h, ref_start=ref_lo, ref_len=len(ref_h)) for t in BINNED_SWEEP]
            bsum = {}

            def be2e_step(check=False):
                for c_lo in range(lo, hi, gchunk):
                    c_hi = min(hi, c_lo + gchunk)
                    n = c_hi - c_lo
                    # the chunk's reference span plus the GC windows reaching beyond it
                    r_lo, r_hi = max(ref_lo, c_lo - W), min(ref_lo + len(ref_h), c_hi + W)
                    p = ctx.pileup_params('coverage_gc', start=c_lo, end=c_hi, ref=ref_h[r_lo - ref_lo:], ref_start=r_lo,
                                          ref_len=r_hi - r_lo, contig_len=genome, window_size=W)
                    o = go if n == gchunk else np.ndarray((3, n), dtype=np.int32, buffer=go)
                    ctx.pileup_host(host_view(hb, c_lo, c_hi), p, o, gm)
                    if check:
                        bsum['coverage_gc'] = bsum.get('coverage_gc', 0) + int(o[1].sum(dtype=np.int64))
                        bsum['gc'] = bsum.get('gc', 0) + int(o[0].sum(dtype=np.int64))
                ctx.binned_host_multi(rs_bin_h, bps_h, bo)
                if check:
                    for t, o in zip(BINNED_SWEEP, bo):
                        bsum[t] = int(o[1 if t == 'coverage' else 0].sum(dtype=np.int64))

            be2e_step(check=True)
            assert bsum == binned['checksum_rank0'], ('binned host path disagrees with the device path', bsum, binned['checksum_rank0'])
            h0, d0 = ctx.transfer_bytes()
            barrier()
            t0 = time.perf_counter()
            bsteps = max(2, min(args.steps, 5))
            for _ in range(bsteps):
                be2e_step()
            torch.cuda.synchronize()
            bsec = max_over_ranks((time.perf_counter() - t0) / bsteps)
            barrier()
            h1, d1 = ctx.transfer_bytes()
            binned['e2e'] = {'value': total_pos / bsec, 'unit': UNIT, 'ms_per_step': bsec * 1e3, 'steps': bsteps,
                             'h2d_bytes_per_step': int(sum_over_ranks((h1 - h0) / bsteps)),
                             'd2h_bytes_per_step': int(sum_over_ranks((d1 - d0) / bsteps)),
                             'api': 'pss_pileup_host (coverage_gc; %d position chunks of %d) + pss_binned_host_multi (coverage, mapq, tlen in one '
                                    'call): pinned host read arrays in, host int32 columns out' % (-(-npos // gchunk), gchunk)}
            del go, gm, bo
    clocks = sampler.stop() if rank == 0 else None
    del hb, ref_h

    # ---- CPU baselines (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_arm
        cores = cpu_arm.usable_cores()
        cchunk = int(args.cpu_chunk_mb * 1e6)
        res = cpu_arm.run(cores=cores, chunk=cchunk, steps=1, warmup=0, depth=args.depth, read_len=args.read_len)
        cs = res['step_seconds'][0]
        cpu = {'value': res['positions_per_step'] / cs, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '%d contig chunks of %.2f Mb, one process each, same synthetic workload and statistics; '
                         '%.1f s wall' % (cores, cchunk / 1e6, cs),
               'reads_per_s': res['reads_per_step'] / cs}
        if binned is not None:
            binned['cpu_baseline'] = cpu_binned(args, 1, 0)

    # ---- the public python API from a real BAM + BAI (rank 0, N = 1): decode included ----------------------
    e2e_api = None
    if rank == 0 and world == 1 and not args.no_e2e_api:
        try:
            e2e_api = run_e2e_api(args, dev)
        except Exception as e:   # noqa: BLE001 -- reported, the headline numbers stand
            e2e_api = {'error': repr(e)}

    # ---- every pileup statistic device-resident on one interval (N = 1): the kernels beside the headline pair ------
    pileup_stats = None
    if rank == 0 and world == 1 and not args.no_stats:
        try:
            pileup_stats = run_pileup_stats(args, dev, peak)
        except Exception as e:   # noqa: BLE001 -- reported, the headline numbers stand
            pileup_stats = {'error': repr(e)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': dev_ms, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic', 'config': workload_config(args, world),
                'reads_per_s': total_reads / (dev_ms * 1e-3), 'gpu_launches': int(launches),
                'e2e': e2e, 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
                'binned': binned, 'strong_scaling': strong, 'e2e_api': e2e_api, 'pileup_stats': pileup_stats,
                'reads_per_gpu': n_reads, 'rows_emitted_rank0': rows, 'checksum_rank0': chk,
                'setup_seconds': {'synthetic_batch': t_gen}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_pileup_stats(args, dev, peak):
    """Device-resident time of EVERY pileup statistic (the 15 types of pysamstats/config.py) on one synthetic interval
    of --stats-mb Mb with base qualities, default filters and min_baseq = 20: ms per call (CUDA events, 3 calls after
    2 warm-up calls, outputs larger than L2), and the fraction of the measured HBM copy bandwidth its algorithmic bytes
    -- 4 F + 2 per position, 29 + CIGAR (+ packed bases) (+ qualities) per read -- reach.  scripts/stat_sweep.py is the
    stand-alone form of this."""
    import torch
    from pysamstats_b200 import _cabi, config, synth
    L = int(args.stats_mb * 1e6) // 512 * 512
    batch, (ref, ref_lo) = synth.make_interval_batch(0, L, L, depth=args.depth, read_len=args.read_len, device=dev, with_qual=True)
    n = int(batch['n_reads'])
    ctx = _cabi.Context(dev.index)
    rs = _cabi.make_reads_struct(batch)
    s = torch.cuda.current_stream().cuda_stream
    out_stats = {}
    for minq in (0, 20):
        ctx.prepare(rs, with_qual=True, stream=s)
        for stat in config.pileup_types:
            nf = len(config.value_fields(stat))
            out = torch.empty((nf, L), dtype=torch.int32, device=dev)
            mask = torch.empty(L, dtype=torch.uint8, device=dev)
            withref = stat in config.stats_types_withref
            p = ctx.pileup_params(stat, start=0, end=L, ref=ref if withref else None, ref_start=ref_lo,
                                  ref_len=int(ref.numel()) if withref else 0, contig_len=L, min_baseq=minq)
            for _ in range(2):
                ctx.pileup_prepared(p, out, mask, stream=s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ctx.pileup_prepared(p, out, mask, stream=s)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            alg = (L * (4 * nf + 2) + n * 29 + 4 * int(batch['n_cigar_words'])
                   + (int(batch['n_seq_bytes']) if (withref and stat != 'coverage_gc') else 0)
                   + (int(batch['n_qual_bytes']) if stat.startswith('baseq') else 0))
            d = out_stats.setdefault(stat, {})
            if minq == 0:
                d.update(ms=ms, positions_per_s=L / (ms * 1e-3), frac=alg / (ms * 1e-3) / 1e9 / peak)
            else:
                d.update(ms_min_baseq_20=ms)
            del out, mask
    ctx.close()
    return {'what': 'every pileup statistic, device resident, one %g Mb interval (%d reads, base qualities staged); ms per call, '
                    'frac = algorithmic bytes / time / measured HBM copy bandwidth' % (L / 1e6, n),
            'interval_mb': L / 1e6, 'reads': n, 'peak_gbs': peak, 'stats': out_stats}


def sum_mapped_reads(batch, i0, i1):
    """Reads of batch[i0:i1] every BinnedStat.recv counts (mapped: flag & 4 == 0) -- the binned checksum."""
    f = batch['flag'][i0:i1]
    return int(((f & 4) == 0).sum().item())


def run_e2e_api(args, dev):
    """What a pysamstats user experiences: ``load_variation_strand(bam, fasta)`` + ``load_coverage_ext(bam)`` on the
    synthetic workload written as a real BAM + BAI + FASTA, cold (decode included), host record arrays out."""
    import tempfile
    import numpy as np
    import pysamstats_b200 as pysamstats
    from pysamstats_b200 import bamio, synth
    L = int(args.api_mb * 1e6) // ALIGN * ALIGN
    sb, (sref, sref_lo) = synth.make_interval_batch(0, L, L, depth=args.depth, read_len=args.read_len, device=dev, with_qual=True)
    snb = synth.batch_to_numpy(sb)
    ref_ascii = sref.cpu().numpy()
    del sb
    out = {'contig_positions': L, 'reads': int(snb['n_reads'])}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'wgs.bam')
        fa = os.path.join(tmp, 'wgs.fa')
        t0 = time.perf_counter()
        bamio.write_bam(path, ['chrS'], [L], snb, level=1)
        bamio.index_bam(path)
        with open(fa, 'wb') as fh:
            fh.write(b'>chrS\n')
            fh.write(ref_ascii[:L].tobytes())
            fh.write(b'\n')
        out['write_seconds'] = time.perf_counter() - t0
        out['bam_bytes'] = os.path.getsize(path)
        del snb
        rows = 0
        for rep in ('cold', 'warm'):
            if rep == 'cold':
                bamio._bam_cache.clear()
                bamio._fa_cache.clear()
            t0 = time.perf_counter()
            a = pysamstats.load_variation_strand(path, fa, chrom='chrS')
            t1 = time.perf_counter()
            b = pysamstats.load_coverage_ext(path, chrom='chrS')
            t2 = time.perf_counter()
            rows = len(a)
            out[rep] = {'load_variation_strand_s': t1 - t0, 'load_coverage_ext_s': t2 - t1,
                        'positions_per_s': L / (t2 - t0), 'rows': rows, 'row_bytes': a.dtype.itemsize + b.dtype.itemsize}
            assert int(a['reads_all'].astype(np.int64).sum()) == int(b['reads_all'].astype(np.int64).sum())
            del a, b
        bf = bamio.open_bam(path)
        if getattr(bf, 'decode_seconds', None):
            out['decode'] = dict(bf.decode_seconds)
    out['value'] = out['cold']['positions_per_s']
    out['unit'] = UNIT
    out['api'] = ('pysamstats.load_variation_strand(bam, fasta, chrom) + load_coverage_ext(bam, chrom) on a %.0f Mb synthetic 30x BAM '
                  '(BGZF level 1) + BAI + FASTA; cold = first call in the process (BGZF inflate + BAM parse inside), warm = decoded '
                  'batch cached; result = packed host record arrays' % (L / 1e6))
    return out


if __name__ == '__main__':
    sys.exit(main())
