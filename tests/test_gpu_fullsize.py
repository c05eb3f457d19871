"""GPU: the bench workload at BASELINE.json's full size (one 250 Mb interval, 50 M reads, 30x) checked
through size-independent properties -- the CPU oracle would need hours here:

* per-position linear identities of the counters (strand / proper-pair marginals, bases vs matches),
* checksums against sums computed independently from the read records (covered bases, deleted bases,
  insertions),
* agreement between statistics that share counters, and shard concatenation == whole at this size.
"""
import numpy as np
import pytest

from pysamstats_b200 import _cabi, config, shard, synth

pytestmark = pytest.mark.gpu

L = 250_000_000 // 512 * 512


def test_full_size_properties():
    import torch
    ctx = _cabi.get_context()
    dev = torch.device('cuda', ctx.device)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 110e9:
        pytest.skip('needs ~100 GB of device memory')
    batch, (ref, ref_lo) = synth.make_interval_batch(0, L, L, depth=30, device=dev, with_qual=False)
    n = int(batch['n_reads'])
    assert 49_000_000 < n < 51_000_000
    rs = _cabi.make_reads_struct(batch)
    s = torch.cuda.current_stream().cuda_stream
    ctx.prepare(rs, stepper='all', max_depth=8000, stream=s)
    outs = {}
    for stat in ('variation_strand', 'coverage_ext', 'coverage_strand'):
        nf = len(config.value_fields(stat))
        o = torch.empty((nf, L), dtype=torch.int32, device=dev)
        m = torch.empty(L, dtype=torch.uint8, device=dev)
        withref = stat.startswith('variation')
        p = ctx.pileup_params(stat, start=0, end=L, ref=ref if withref else None, ref_start=ref_lo,
                              ref_len=int(ref.numel()) if withref else 0, contig_len=L)
        ctx.pileup_prepared(p, o, m, stream=s)
        outs[stat] = (o, m)
    ctx.synchronize(s)
    vs, vmask = outs['variation_strand']
    f = {k: j for j, k in enumerate(config.value_fields('variation_strand'))}
    cs = outs['coverage_strand'][0]
    ce = outs['coverage_ext'][0]
    # ---- per-position identities, in slabs (int32 arithmetic, no copies of the 60 GB array) -------------
    step = 25_000_000 // 512 * 512
    tot = dict(reads=0, dels=0, ins=0, rows=0, mism=0)
    for a in range(0, L, step):
        b = min(L, a + step)
        v = vs[:, a:b]
        for g in ('reads', 'matches', 'mismatches', 'deletions', 'insertions', 'A', 'C', 'T', 'G', 'N'):
            base = 'reads_all' if g == 'reads' else g
            k = f[base]
            assert torch.equal(v[k + 1] + v[k + 2], v[k]), (g, 'fwd + rev')
            assert torch.equal(v[k + 4] + v[k + 5], v[k + 3]), (g, 'pp_fwd + pp_rev')
            assert bool((v[k + 3] <= v[k]).all()) and bool((v[k + 4] <= v[k + 1]).all()), (g, 'pp <= all')
        nondel = v[f['matches']] + v[f['mismatches']]
        assert torch.equal(nondel + v[f['deletions']], v[f['reads_all']])             # the workload has no N ops
        assert torch.equal(v[f['A']] + v[f['C']] + v[f['T']] + v[f['G']] + v[f['N']], nondel)
        assert bool((v[f['insertions']] <= nondel).all())
        # statistics that share counters agree
        for j in range(6):
            assert torch.equal(cs[j, a:b], v[j]), j
        assert torch.equal(ce[0, a:b], v[0]) and torch.equal(ce[1, a:b], v[3])
        assert torch.equal((v[0] > 0).to(torch.uint8), vmask[a:b])
        tot['reads'] += int(v[0].sum(dtype=torch.int64))
        tot['dels'] += int(v[f['deletions']].sum(dtype=torch.int64))
        tot['ins'] += int(v[f['insertions']].sum(dtype=torch.int64))
        tot['mism'] += int(v[f['mismatches']].sum(dtype=torch.int64))
        tot['rows'] += int(vmask[a:b].sum(dtype=torch.int64))
    # ---- checksums from the read records themselves --------------------------------------------------------
    flag = batch['flag'].to(torch.int64) & 0xffff
    pos, end = batch['pos'].to(torch.int64), batch['endpos'].to(torch.int64)
    adm = ((flag & 0x704) == 0) & ((flag & 4) == 0)                      # stepper 'all'; depth 30 never meets the cap
    span = (torch.clamp(end, max=L) - torch.clamp(pos, min=0)).clamp(min=0)
    assert tot['reads'] == int(span[adm].sum())
    cig = batch['cigar'].to(torch.int64) & 0xffffffff
    ncig = (batch['cigar_off'][1:] - batch['cigar_off'][:-1]).to(torch.int64)
    owner = torch.repeat_interleave(torch.arange(n, device=dev), ncig)
    op, ln = cig & 15, cig >> 4
    inside = adm[owner] & (end[owner] <= L)                               # reads wholly inside the interval
    edge = adm & (end > L)
    assert int(edge.sum()) < 1000
    d_lo = int(ln[(op == 2) & inside].sum())
    assert d_lo <= tot['dels'] <= d_lo + int(ln[(op == 2) & edge[owner]].sum())
    i_lo = int(((op == 1) & inside).sum())
    assert i_lo <= tot['ins'] <= i_lo + int(((op == 1) & edge[owner]).sum())
    assert 0.004 < tot['mism'] / tot['reads'] < 0.009                     # 0.5 % substitutions + 0.1 % N
    assert tot['rows'] > 0.999 * L
    # ---- shards + halo == whole, at this size (coverage_ext: 8 columns) ------------------------------------
    posn = batch['pos'].cpu().numpy()
    endn = batch['endpos'].cpu().numpy()
    for lo, hi in shard.plan_intervals(0, L, 2):
        i0, i1, _ = shard.halo_range(posn, endn, lo, hi, mask=False)
        view = {k: (batch[k][i0:] if k in ('pos', 'flag', 'mapq', 'tlen', 'mtid', 'l_seq', 'cigar_off', 'seq_off') else batch[k])
                for k in _cabi.READ_FIELDS if batch.get(k) is not None}
        view['tid'] = 0
        sub = _cabi.make_reads_struct(view, n_reads=i1 - i0)
        o = torch.empty((8, hi - lo), dtype=torch.int32, device=dev)
        m = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
        p = ctx.pileup_params('coverage_ext', start=lo, end=hi, contig_len=L)
        ctx.pileup(sub, p, o, m, stream=s)
        ctx.synchronize(s)
        assert torch.equal(o, ce[:, lo:hi]) and torch.equal(m, outs['coverage_ext'][1][lo:hi]), (lo, hi)
