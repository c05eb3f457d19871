"""CPU: pins the parity oracle (oracle/) to the reference's own golden vectors and
known-answer tests.  These are the reference's tests that do not need pysam:

* pysamstats/test/test_regression.py:34-58  (fixture/regression.npz, all 20 stat types)
* pysamstats/test/test_pileup.py:939-1102   (truncate, pad, whole-genome pad, depth limit,
                                             chrom dtype width, steppers)
* pysamstats/test/test_binned.py:583-626    (region / whole-genome bins)
"""
import numpy as np
import pytest

from oracle import bamread, oracle as orc
from _util import fx, assert_recarrays_equal


@pytest.fixture(scope='module')
def test_bam():
    return bamread.Bam(fx('test.bam'))


@pytest.fixture(scope='module')
def ref_fa():
    return bamread.Fasta(fx('ref.fa'))


@pytest.fixture(scope='module')
def golden():
    return np.load(fx('regression.npz'))


@pytest.mark.parametrize('stat', orc.PILEUP_TYPES)
def test_pileup_regression(stat, test_bam, ref_fa, golden):
    a = orc.load_pileup(stat, test_bam, ref_fa if stat in orc.WITHREF else None)
    assert_recarrays_equal(golden[stat], a, stat)


@pytest.mark.parametrize('stat', orc.BINNED_TYPES)
def test_binned_regression(stat, test_bam, ref_fa, golden):
    a = orc.load_binned(stat, test_bam, ref_fa)
    assert_recarrays_equal(golden[stat + '_binned'], a, stat + '_binned')


def test_depth_limit_kat():
    deep = bamread.Bam(fx('deep.bam'))
    for max_depth, expect in ((1000000, 26169), (12000, 12046), (8000, 8052)):
        a = orc.load_pileup('coverage', deep, max_depth=max_depth, fields=['reads_all'])
        assert a[70] == expect


def test_steppers_kat():
    lc = bamread.Bam(fx('longcontignames.bam'))
    for stepper, ra, rp in (('all', 7, 4), ('nofilter', 8, 5), ('samtools', 4, 4)):
        a = orc.load_pileup('coverage', lc, chrom='AS2_scf7180000695891', stepper=stepper, pad=True)
        assert (a[14311]['reads_all'], a[14311]['reads_pp']) == (ra, rp)
        assert a.dtype['chrom'].itemsize == 20
    with pytest.raises(ValueError):
        orc.load_pileup('coverage', lc, chrom='AS2_scf7180000695891', stepper='notastepper')


def test_truncate_pad_kat(test_bam, ref_fa):
    for stat in ('coverage', 'variation', 'tlen', 'baseq_ext', 'coverage_gc'):
        fa = ref_fa if stat in orc.WITHREF else None
        a = orc.load_pileup(stat, test_bam, fa, chrom='Pf3D7_01_v3', start=2000, end=2100)
        assert (a.pos[0], a.pos[-1]) == (1952, 2154)
        a = orc.load_pileup(stat, test_bam, fa, chrom='Pf3D7_01_v3', start=2000, end=2100, truncate=True)
        assert (a.pos[0], a.pos[-1]) == (2000, 2099)
        a = orc.load_pileup(stat, test_bam, fa, chrom='Pf3D7_01_v3', start=0, end=20000)
        assert (a.pos[0], a.pos[-1]) == (924, 9935)
        a = orc.load_pileup(stat, test_bam, fa, chrom='Pf3D7_01_v3', start=0, end=20000, pad=True)
        assert (a.pos[0], a.pos[-1]) == (0, 19999) and np.all(np.diff(a.pos) == 1)


def test_pad_whole_genome_kat(test_bam):
    a = orc.load_pileup('coverage', test_bam)
    assert sorted(set(a.chrom)) == [b'Pf3D7_01_v3', b'Pf3D7_02_v3']
    x = a[a.chrom == b'Pf3D7_02_v3']
    assert (x.pos[0], x.pos[-1]) == (926, 10074)
    a = orc.load_pileup('coverage', test_bam, pad=True)
    for c, last in ((b'Pf3D7_01_v3', 50000), (b'Pf3D7_02_v3', 60000), (b'Pf3D7_03_v3', 70000)):
        x = a[a.chrom == c]
        assert (x.pos[0], x.pos[-1]) == (0, last)


def test_binned_kat(test_bam, ref_fa):
    a = orc.load_binned('coverage', test_bam, ref_fa, chrom='Pf3D7_01_v3', start=1000, end=20000,
                        window_size=200, window_offset=100)
    assert (a.pos[0], a.pos[-1]) == (1100, 19900)
    a = orc.load_binned('mapq', test_bam, window_size=200, window_offset=100)
    for c, last in ((b'Pf3D7_01_v3', 50100), (b'Pf3D7_02_v3', 60100), (b'Pf3D7_03_v3', 70100)):
        x = a[a.chrom == c]
        assert (x.pos[0], x.pos[-1]) == (100, last)
