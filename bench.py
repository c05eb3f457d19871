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

Numbers on the JSON line:
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
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': mean * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
        'config': workload_config(args, args.gpus),
        'reads_per_s': res['reads_per_step'] / mean,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {'workload': 'synthetic 30x WGS-shaped reads (150 bp paired, 250 Mb contig interval per GPU), '
                        'stat_variation_strand + stat_coverage_ext',
            'interval_positions_per_gpu': int(args.interval_mb * 1e6), 'depth': args.depth,
            'read_len': args.read_len, 'stats': list(STATS), 'min_baseq': 0, 'stepper': 'all', 'max_depth': 8000,
            'sharding': 'contiguous intervals + read halo, %d shard(s), no collective' % n,
            'l2': 'inputs (> 5 GB of read records) and outputs (68 GB) exceed the 126 MB L2; no flush needed'}


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
    from pysamstats_b200 import _cabi, shard, synth

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

    # ---- the rank's shard of the genome ------------------------------------------------
    per = int(args.interval_mb * 1e6) // shard.TILE * shard.TILE
    genome = per * world
    lo, hi = shard.plan_intervals(0, genome, world)[rank]
    npos = hi - lo
    t_gen = time.time()
    batch, (ref, ref_lo) = synth.make_interval_batch(lo, hi, genome, depth=args.depth, read_len=args.read_len,
                                                     device=dev, with_qual=False)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    n_reads = int(batch['n_reads'])
    ctx = _cabi.Context(local_rank)
    reads = _cabi.make_reads_struct(batch)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    nf = {s: int(ctx.lib.pss_pileup_nfields(_cabi.PILEUP_STATS.index(s))) for s in STATS}
    outs = {s: torch.empty((nf[s], npos), dtype=torch.int32, device=dev) for s in STATS}
    masks = {s: torch.empty(npos, dtype=torch.uint8, device=dev) for s in STATS}
    params = {s: ctx.pileup_params(s, start=lo, end=hi, ref=ref if s == 'variation_strand' else None,
                                   ref_start=ref_lo, ref_len=int(ref.numel()) if s == 'variation_strand' else 0,
                                   contig_len=genome) for s in STATS}

    def step():
        # one preparation of the batch (stepper, depth cap, read filters), then both statistics
        ctx.prepare(reads, stepper='all', max_depth=8000, stream=sptr)
        for s in STATS:
            ctx.pileup_prepared(params[s], outs[s], masks[s], stream=sptr)

    # ---- device-resident throughput -----------------------------------------------------
    for _ in range(args.warmup):
        step()
    ctx.synchronize(sptr)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.set_kernel_timing(True)
    ctx.pileup_kernel_ms(-1)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launch_count() - l0
    kms = {s: ctx.pileup_kernel_ms(s) for s in STATS}
    ctx.set_kernel_timing(False)
    ctx.synchronize(sptr)
    dev_ms = max_over_ranks(dev_ms)
    total_pos = sum_over_ranks(npos)
    total_reads = sum_over_ranks(n_reads)
    value = total_pos / (dev_ms * 1e-3)

    # checksum of checksums (a size-independent sanity check, printed to stderr)
    chk = {s: int(outs[s][0].sum(dtype=torch.int64).item()) for s in STATS}
    rows = int(masks['variation_strand'].sum(dtype=torch.int64).item())

    # ---- roofline of the dominant kernel (rank 0's launch) --------------------------------
    n_cig = int(batch['n_cigar_words'])
    n_seq = int(batch['n_seq_bytes'])
    alg_bytes = npos * (4 * nf['variation_strand'] + 1 + 1) + n_reads * 29 + 4 * n_cig + n_seq
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    # DRAM traffic of the same kernel on the same workload from one `ncu --set full` capture
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get('interval_positions', -1)) == per:
            traffic = tj['traffic_bytes_per_launch']
    k_ms, k_n = kms['variation_strand']
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    roofline = {'bound': 'hbm', 'kernel': 'k_tile<KIND_VAR, PSS_VARIATION_STRAND>', 'achieved': achieved,
                'peak': peak, 'unit': 'GB/s', 'frac': (achieved / peak) if achieved else None, 'traffic': traffic,
                'peak_source': peak_src, 'kernel_ms': k_ms, 'kernel_launches_timed': k_n,
                'algorithmic_bytes_per_launch': alg_bytes,
                'kernel_share_of_step': (k_ms / dev_ms) if k_ms > 0 else None,
                'coverage_ext_kernel_ms': kms['coverage_ext'][0]}

    # ---- end to end through the host-buffer C ABI ------------------------------------------
    e2e = None
    if not args.no_e2e:
        del outs, masks
        torch.cuda.empty_cache()
        hb = {}
        for k in _cabi.READ_FIELDS:
            v = batch.get(k)
            hb[k] = v.cpu().pin_memory().numpy() if torch.is_tensor(v) else None
        hb['flag'] = hb['flag'].view(np.uint16)
        hb['cigar'] = hb['cigar'].view(np.uint32)
        endpos = batch['endpos'].cpu().numpy()
        maxspan = int((endpos - hb['pos']).max()) if n_reads else 1
        ref_h = ref.cpu().pin_memory().numpy()
        chunk = max(shard.TILE, int(args.e2e_chunk_mb * 1e6) // shard.TILE * shard.TILE)
        chunk = min(chunk, npos)
        inner = max(shard.TILE, int(args.e2e_inner_mb * 1e6) // shard.TILE * shard.TILE)
        oh = {s: torch.empty((nf[s], chunk), dtype=torch.int32).pin_memory().numpy() for s in STATS}
        mh = {s: torch.empty(chunk, dtype=torch.uint8).pin_memory().numpy() for s in STATS}
        checks = {}

        def e2e_step(check=False):
            for c_lo in range(lo, hi, chunk):
                c_hi = min(hi, c_lo + chunk)
                i0, i1, _ = shard.halo_range(hb['pos'], endpos, c_lo, c_hi, maxspan=maxspan, mask=False)
                view = dict(n_reads=i1 - i0, tid=0)
                for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off'):
                    view[k] = hb[k][i0:]
                view['cigar'], view['seq4'] = hb['cigar'], hb['seq4']
                rs = _cabi.make_reads_struct(view, n_reads=i1 - i0)
                n = c_hi - c_lo
                r0 = c_lo - ref_lo
                ps, os_, ms = [], [], []
                for s in STATS:
                    withref = s == 'variation_strand'
                    ps.append(ctx.pileup_params(s, start=c_lo, end=c_hi, ref=ref_h[r0:] if withref else None,
                                                ref_start=c_lo, ref_len=(c_hi - c_lo) if withref else 0,
                                                contig_len=genome))
                    os_.append(oh[s] if n == chunk else np.ndarray((nf[s], n), dtype=np.int32, buffer=oh[s]))
                    ms.append(mh[s])
                # one call per position chunk: the reads go up once, both statistics come back
                if args.e2e_stream:
                    ctx.pileup_host_stream(rs, ps, os_, ms, maxspan, inner)
                else:
                    ctx.pileup_host_multi(rs, ps, os_, ms)
                if check:
                    for s, o in zip(STATS, os_):
                        checks[s] = checks.get(s, 0) + int(o[0].sum(dtype=np.int64))

        for _ in range(max(1, min(args.warmup, 2))):
            checks.clear()
            e2e_step(check=True)
        assert checks == chk, ('host path disagrees with the device path', checks, chk)
        h0, d0 = ctx.transfer_bytes()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps
        barrier()
        h1, d1 = ctx.transfer_bytes()
        e2e_s = max_over_ranks(e2e_s)
        e2e = {'value': total_pos / e2e_s, 'unit': UNIT, 'ms_per_step': e2e_s * 1e3,
               'h2d_bytes_per_step': int(sum_over_ranks((h1 - h0) / args.steps)),
               'd2h_bytes_per_step': int(sum_over_ranks((d1 - d0) / args.steps)),
               'api': ('pss_pileup_host_stream (pinned host read arrays in, host int32 columns of both statistics out), '
                       '%d calls of %d positions per step, reads uploaded in pieces of %d positions'
                       % (-(-npos // chunk), chunk, inner)) if args.e2e_stream else
                      ('pss_pileup_host_multi (pinned host read arrays in, host int32 columns of both statistics out), '
                       '%d position chunks of %d per step' % (-(-npos // chunk), chunk))}
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
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

    # ---- host decode (rank 0, N = 1): synthetic BAM slice -> SoA through csrc/pss_host.cpp ----------
    host_decode = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import tempfile
            from pysamstats_b200 import bamio
            sl = 8 * 1000 * 1000
            sb, _ = synth.make_interval_batch(0, sl, sl, depth=args.depth, read_len=args.read_len, device=dev,
                                              with_qual=True)
            snb = synth.batch_to_numpy(sb)
            with tempfile.TemporaryDirectory() as tmp:
                path = os.path.join(tmp, 'slice.bam')
                bamio.write_bam(path, ['chrS'], [sl], snb, level=1)
                t0 = time.perf_counter()
                bf = bamio.BamFile(path).load()
                dt = time.perf_counter() - t0
                host_decode = {'reads_per_s': snb['n_reads'] / dt, 'seconds': dt, 'reads': int(snb['n_reads']),
                               'bam_bytes': os.path.getsize(path), 'raw_bytes': bf.decode_seconds['raw_bytes'],
                               'inflate_seconds': bf.decode_seconds['inflate'], 'threads': os.cpu_count(),
                               'what': 'BGZF inflate + BAM parse of an 8 Mb slice of the workload written as a real BAM '
                                       '(libpss_host.so); not part of value / e2e'}
            del sb, snb
        except Exception as e:   # noqa: BLE001 -- informative only
            host_decode = {'error': repr(e)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': dev_ms, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic', 'config': workload_config(args, world),
                'reads_per_s': total_reads / (dev_ms * 1e-3), 'gpu_launches': int(launches),
                'e2e': e2e, 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
                'host_decode': host_decode, 'reads_per_gpu': n_reads, 'rows_emitted_rank0': rows, 'checksum_rank0': chk,
                'setup_seconds': {'synthetic_batch': t_gen}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
