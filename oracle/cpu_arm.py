"""TEST / BENCH INFRASTRUCTURE ONLY -- the CPU arm of bench.py.

Times the CPU restatement of the reference's pileup path (oracle/pileup_oracle.c, the
"port": the reference's own Cython cannot be built here, see oracle/Makefile) the way
BASELINE.md section 5 prescribes: one process per contig chunk, all processes running at
the same time, wall clock around the slowest.  Used for ``cpu_baseline`` in the default
bench line and for ``bench.py --impl reference``; never on the product path.

Each worker generates its own chunk of the synthetic 30x WGS workload (same generator and
parameters as the GPU arm, CPU random stream), then runs ``steps`` timed passes of
stat_variation_strand + stat_coverage_ext over it between two barriers -- or, with
``workload='binned'``, of the binned sweep of BASELINE.json configs[4]: stat_coverage_gc
(a pileup) + coverage_binned + mapq_binned + tlen_binned with window_size 300.
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STATS = (('variation_strand', True), ('coverage_ext', False))
BINNED_SWEEP = ('coverage', 'mapq', 'tlen')          # + the pileup statistic coverage_gc; window_size 300


def usable_cores(cap=128):
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, cap))


def _pack(nb):
    """oracle.PackedReads over the numpy SoA arrays of a synthetic batch."""
    import numpy as np
    from oracle import oracle as orc
    n = int(nb['n_reads'])
    r = orc.PackedReads.__new__(orc.PackedReads)
    r.n = n
    for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar', 'cigar_off', 'seq4', 'seq_off', 'qual',
              'qual_off'):
        setattr(r, k, np.ascontiguousarray(nb[k]))
    prev = nb['name_prev'].astype('i8')
    me = np.arange(n, dtype='i8')
    r.name_id = np.where(prev >= 0, prev, me)      # synthetic QNAMEs occur at most twice
    r.n_names = n
    r.tid = int(nb.get('tid', 0))
    return r


def _worker(i, n_workers, chunk, depth, read_len, steps, warmup, bar, q, workload='pileup'):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import torch
        torch.set_num_threads(1)
        from pysamstats_b200 import synth
        from oracle import oracle as orc
        orc.lib()
        genome = n_workers * chunk
        lo, hi = i * chunk, (i + 1) * chunk
        batch, (ref, ref_lo) = synth.make_interval_batch(lo, hi, genome, depth=depth, read_len=read_len,
                                                         device='cpu', with_qual=True,
                                                         block=min(1000000, chunk))
        nb = synth.batch_to_numpy(batch)
        reads = _pack(nb)
        # the oracle addresses the reference from position 0 of the contig
        import numpy as np
        refb = np.zeros(ref_lo + ref.numel(), dtype='u1')
        refb[ref_lo:] = ref.numpy()
        refb = refb.tobytes()
        q.put(('ready', i, int(reads.n)))
        rows = 0
        for _ in range(warmup + steps):
            bar.wait()
            rows = 0
            if workload == 'binned':
                pos, F, rc = orc.pileup_packed('coverage_gc', reads, lo, hi, ref=refb, truncate=True, window_size=300)
                rows += len(pos)
                for stat in BINNED_SWEEP:
                    pos, F = orc.binned_packed(stat, reads, lo, hi, 300, 150, ref=refb)
                    rows += len(pos)
            else:
                for stat, withref in STATS:
                    pos, F, rc = orc.pileup_packed(stat, reads, lo, hi, ref=refb if withref else None, truncate=True)
                    rows += len(pos)
            bar.wait()
        q.put(('done', i, rows))
    except BaseException as e:   # noqa: BLE001 -- report and release the barrier
        q.put(('error', i, repr(e)))
        try:
            bar.abort()
        except Exception:
            pass


def run(cores=None, chunk=1000000, steps=1, warmup=0, depth=30, read_len=150, workload='pileup'):
    """Returns dict(step_seconds=[...], positions_per_step, reads_per_step, cores, chunk)."""
    cores = cores or usable_cores()
    ctx = mp.get_context('spawn')
    bar = ctx.Barrier(cores + 1)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(i, cores, chunk, depth, read_len, steps, warmup, bar, q, workload), daemon=True)
             for i in range(cores)]
    for p in procs:
        p.start()
    n_reads = 0
    for _ in range(cores):
        msg = q.get(timeout=1800)
        if msg[0] == 'error':
            for p in procs:
                p.terminate()
            raise RuntimeError('cpu arm worker %d failed: %s' % (msg[1], msg[2]))
        n_reads += msg[2]
    times = []
    for k in range(warmup + steps):
        bar.wait(timeout=3600)
        t0 = time.perf_counter()
        bar.wait(timeout=3600)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
    for _ in range(cores):
        msg = q.get(timeout=600)
        if msg[0] == 'error':
            raise RuntimeError('cpu arm worker %d failed: %s' % (msg[1], msg[2]))
    for p in procs:
        p.join(timeout=60)
    return {'step_seconds': times, 'positions_per_step': cores * chunk, 'reads_per_step': n_reads,
            'cores': cores, 'chunk': chunk}
