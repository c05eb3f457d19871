"""CPU: interval sharding (pysamstats_b200/shard.py) and the block-seeded synthetic workload.

The N > 1 path of bench.py has no data-path collective: every rank computes its interval from
its own reads + halo and rank 0 only concatenates.  The world_size-2 test below runs exactly
that plan over gloo with the CPU oracle standing in for the CUDA engine (no GPU here), and
checks the concatenation against the unsharded result.
"""
import os
import sys

import numpy as np
import pytest
import torch

from pysamstats_b200 import shard, synth
from _util import ROOT


def test_plan_intervals():
    for (s, e, n) in ((0, 250_000_000, 8), (17, 1000, 3), (5, 5, 2), (0, 100, 4), (1000, 20000, 1)):
        iv = shard.plan_intervals(s, e, n)
        assert len(iv) == n and iv[0][0] == s and iv[-1][1] == max(s, e)
        for (a, b), (c, d) in zip(iv, iv[1:]):
            assert b == c and a <= b
        for a, b in iv[:-1]:
            assert (b - s) % shard.TILE == 0 or b == e
    with pytest.raises(ValueError):
        shard.plan_intervals(0, 10, 0)


def test_plan_windows_and_read_ranges():
    rng = np.random.default_rng(3)
    pos = np.sort(rng.integers(0, 5000, 4000)).astype('i4')
    start, end, W = 100, 4700, 300
    n_bins = -(-(end - start) // W)
    # reference roll-over rule (opt.pyx:2050): window k >= 1 owns start + kW < pos <= start + (k+1)W
    k = np.clip((pos.astype('i8') - start - 1) // W, 0, n_bins - 1)
    k[pos <= start + W] = 0
    plan = shard.plan_windows(start, end, W, 3)
    assert plan[0][0] == 0 and plan[-1][1] == n_bins
    seen = 0
    for k_lo, k_hi in plan:
        i0, i1 = shard.window_read_range(pos, start, W, k_lo, k_hi, n_bins)
        assert i0 == seen
        assert np.all((k[i0:i1] >= k_lo) & (k[i0:i1] < k_hi))
        seen = i1
    assert seen == len(pos)


def test_halo_range_matches_region_rule():
    rng = np.random.default_rng(5)
    pos = np.sort(rng.integers(0, 100000, 20000)).astype('i4')
    endpos = pos.astype('i8') + rng.integers(1, 400, len(pos))
    maxspan = int((endpos - pos).max())
    for lo, hi in ((0, 100000), (5000, 5001), (99990, 100500), (40000, 60000), (200000, 300000)):
        m = (pos < hi) & (endpos > lo)
        for ms in (None, maxspan):
            i0, i1, ov = shard.halo_range(pos, endpos, lo, hi, maxspan=ms)
            assert ov.sum() == m.sum()
            if m.any():
                idx = np.flatnonzero(m)
                assert i0 == idx[0] and i1 == idx[-1] + 1 and np.array_equal(ov, m[i0:i1])
            assert shard.halo_range(pos, endpos, lo, hi, maxspan=ms, mask=False)[:2] == (i0, i1)


def test_safe_cut():
    pos = np.array([0, 10, 10, 10, 10, 300, 600], 'i4')
    endpos = np.array([100, 400, 400, 400, 400, 700, 700], 'i8')
    assert shard.live_depth_at(pos, endpos, 50) == 5 and shard.live_depth_at(pos, endpos, 450) == 1
    assert shard.safe_cut(pos, endpos, 256, max_depth=4, align=256) == 512   # depth 4 at 256, 2 at 512
    assert shard.safe_cut(pos, endpos, 256, max_depth=8000) == 256


def test_synthetic_shards_agree():
    G = 600_000
    full, (ref, r0) = synth.make_interval_batch(0, G, G, block=100000, with_qual=True)
    a, (refa, ra) = synth.make_interval_batch(0, 256 * 1000, G, block=100000, with_qual=True)
    b, (refb, rb) = synth.make_interval_batch(256 * 1000, G, G, block=100000, with_qual=True)
    fp, fe = full['pos'].long(), full['endpos']
    for part, lo, hi in ((a, 0, 256000), (b, 256000, G)):
        keep = (fp < hi) & (fe > lo)
        for k in ('pos', 'flag', 'mapq', 'tlen', 'mtid'):
            assert torch.equal(full[k][keep], part[k]), k
        L = 150
        assert torch.equal(full['qual'].view(-1, L)[keep], part['qual'].view(-1, L))
        assert torch.equal(full['seq4'].view(-1, 75)[keep], part['seq4'].view(-1, 75))
    assert torch.equal(ref[ra - r0:ra - r0 + refa.numel()], refa)
    assert torch.equal(ref[rb - r0:rb - r0 + refb.numel()], refb)
    # name_prev links stay inside the shard
    for part in (a, b):
        prev = part['name_prev']
        ok = prev >= 0
        assert torch.all(prev[ok] < torch.nonzero(ok).flatten())
    depth = float((fe - fp).clamp(min=0).sum()) / G
    assert 27 < depth < 33


def _rank_main(rank, world, port, G, stat, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from oracle import oracle as orc
    from _util import oracle_reads_from_arrays
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard.plan_intervals(0, G, world, align=38400)[rank]       # lcm(512, 300): tile AND window aligned, as bench.py cuts
    # (reads that start exactly at hi belong to the rank's last window -- strict roll-over, opt.pyx:2050 -- so the batch reaches one
    # position further than the pileup range; they cover no position below hi)
    batch, (ref, ref_lo) = synth.make_interval_batch(lo, min(hi + 1, G), G, block=100000, with_qual=True)
    nb = synth.batch_to_numpy(batch)
    refb = np.zeros(ref_lo + ref.numel(), 'u1')
    refb[ref_lo:] = ref.numpy()
    pos, F, _ = orc.pileup_packed(stat, oracle_reads_from_arrays(nb), lo, hi, ref=refb.tobytes(), truncate=True)
    # "concatenate on the host": gather the pieces on rank 0 in rank order
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((pos, F), gathered, dst=0)
    n_rows = torch.tensor([len(pos)], dtype=torch.int64)
    dist.all_reduce(n_rows)                      # the only reduction bench.py does: counts / max time
    # the binned sweep of bench.py's `binned` arm: the rank's windows (global bin edges), the reads that START in them
    W = 300
    n_all = -(-G // W)
    i0, i1 = shard.window_read_range(nb['pos'], 0, W, lo // W, -(-hi // W), n_all)
    sub = {k: (v[i0:i1] if k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'endpos', 'name_prev') else v) for k, v in nb.items()}
    for off in ('cigar_off', 'seq_off', 'qual_off'):
        sub[off] = nb[off][i0:i1 + 1]
    sub['n_reads'] = i1 - i0
    binned = {t: orc.binned_packed(t, oracle_reads_from_arrays(sub), lo, hi, W, W // 2, ref=refb.tobytes())[1] for t in ('coverage', 'mapq', 'tlen')}
    bg = [None] * world if rank == 0 else None
    dist.gather_object(binned, bg, dst=0)
    if rank == 0:
        out = shard.concat_shards([(p, f.T) for p, f in gathered])
        np.savez(tmp, pos=out[0], F=out[1].T, total=int(n_rows),
                 **{'binned_' + t: np.concatenate([b[t] for b in bg]) for t in ('coverage', 'mapq', 'tlen')})
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    from _util import oracle_reads_from_arrays
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    G, stat = 400_000, 'variation_strand'
    tmp = str(tmp_path / 'sharded.npz')
    mp.spawn(_rank_main, args=(2, port, G, stat, tmp), nprocs=2, join=True)
    got = np.load(tmp)
    batch, (ref, ref_lo) = synth.make_interval_batch(0, G, G, block=100000, with_qual=True)
    pos, F, _ = orc.pileup_packed(stat, oracle_reads_from_arrays(synth.batch_to_numpy(batch)), 0, G,
                                  ref=ref.numpy().tobytes(), truncate=True)
    assert int(got['total']) == len(pos)
    assert np.array_equal(got['pos'], pos) and np.array_equal(got['F'], F)
    reads = oracle_reads_from_arrays(synth.batch_to_numpy(batch))
    for t in ('coverage', 'mapq', 'tlen'):       # window ranges side by side == the whole contig
        e = orc.binned_packed(t, reads, 0, G, 300, 150, ref=ref.numpy().tobytes())[1]
        assert np.array_equal(got['binned_' + t], e), t
