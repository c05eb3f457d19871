"""TEST INFRASTRUCTURE ONLY -- independent BAM / FASTA reader for the parity oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module.  The product (``pysamstats_b200``) has its own decoder
(``pysamstats_b200/bamio.py``) and never touches ``oracle/``.

The reference obtains reads from pysam / htslib, which are not vendored under
``/root/reference``.  This reader restates what the reference's
call sites consume:

* ``AlignmentFile.pileup(reference, start, end, ...)`` (reference ``pysamstats/opt.pyx:1853``)
  and ``IteratorRowRegion`` (``opt.pyx:2040``) both yield, in file order, the records of
  one contig with ``pos < end and endpos > start`` -- see ``region_reads``.
* ``FastaFile.fetch`` (``opt.pyx:1989``, ``:1505``, ``:1561``) -- see ``Fasta.fetch``.

BGZF is a series of gzip members, so Python's ``gzip`` module is a complete
decompressor for it; the BAM record layout is the published SAM/BAM spec (section 4.2).
Deliberately simple and slow: it shares no code with the product decoder.
"""
import gzip
import struct

import numpy as np

CIGAR_REF = (1, 0, 1, 1, 0, 0, 0, 1, 1)      # M I D N S H P = X consume reference?


class Bam(object):
    """All records of a BAM file as python lists / numpy arrays (file order)."""

    def __init__(self, path):
        with gzip.open(path, 'rb') as fh:
            data = fh.read()
        assert data[:4] == b'BAM\x01', 'not a BAM file: %r' % path
        l_text, = struct.unpack_from('<i', data, 4)
        off = 8 + l_text
        n_ref, = struct.unpack_from('<i', data, off)
        off += 4
        self.references = []
        self.lengths = []
        for _ in range(n_ref):
            l_name, = struct.unpack_from('<i', data, off)
            off += 4
            self.references.append(data[off:off + l_name - 1].decode('ascii'))
            off += l_name
            l_ref, = struct.unpack_from('<i', data, off)
            off += 4
            self.lengths.append(l_ref)
        tid, pos, mapq, flag, l_seq, mtid, mpos, tlen = ([] for _ in range(8))
        qname, cigar, seq4, qual = [], [], [], []
        n = len(data)
        while off < n:
            (bs, r_tid, r_pos, l_rn, r_mapq, _bin, n_cig, r_flag, r_lseq, r_mtid, r_mpos,
             r_tlen) = struct.unpack_from('<iiiBBHHHiiii', data, off)
            p = off + 36
            qname.append(data[p:p + l_rn - 1])
            p += l_rn
            cigar.append(np.frombuffer(data, dtype='<u4', count=n_cig, offset=p))
            p += 4 * n_cig
            nb = (r_lseq + 1) // 2
            seq4.append(np.frombuffer(data, dtype='u1', count=nb, offset=p))
            p += nb
            qual.append(np.frombuffer(data, dtype='u1', count=r_lseq, offset=p))
            tid.append(r_tid); pos.append(r_pos); mapq.append(r_mapq); flag.append(r_flag)
            l_seq.append(r_lseq); mtid.append(r_mtid); mpos.append(r_mpos); tlen.append(r_tlen)
            off += 4 + bs
        self.tid = np.array(tid, dtype='i4')
        self.pos = np.array(pos, dtype='i4')
        self.mapq = np.array(mapq, dtype='u1')
        self.flag = np.array(flag, dtype='u2')
        self.l_seq = np.array(l_seq, dtype='i4')
        self.mtid = np.array(mtid, dtype='i4')
        self.mpos = np.array(mpos, dtype='i4')
        self.tlen = np.array(tlen, dtype='i4')
        self.qname = qname
        self.cigar = cigar
        self.seq4 = seq4
        self.qual = qual
        # htslib bam_endpos(): pos + reference length of the CIGAR, or pos + 1 when the
        # read is unmapped or consumes no reference
        endpos = np.empty(len(tid), dtype='i8')
        for i, c in enumerate(cigar):
            rl = 0
            if not (flag[i] & 4):
                for w in c:
                    if CIGAR_REF[int(w) & 15]:
                        rl += int(w) >> 4
            endpos[i] = pos[i] + (rl if rl > 0 else 1)
        self.endpos = endpos

    def __len__(self):
        return len(self.tid)

    def region_reads(self, chrom, start, end):
        """Indices (file order) of the records htslib's region iterator returns for
        ``chrom:[start, end)``: same tid, ``pos < end`` and ``endpos > start``."""
        t = self.references.index(chrom)
        m = (self.tid == t) & (self.pos < end) & (self.endpos > start)
        return np.flatnonzero(m)


class Fasta(object):
    """Plain FASTA, whole file in memory; ``fetch`` clips like faidx does."""

    def __init__(self, path):
        self.seqs = {}
        name = None
        chunks = []
        with open(path, 'rb') as fh:
            for line in fh:
                if line.startswith(b'>'):
                    if name is not None:
                        self.seqs[name] = b''.join(chunks)
                    name = line[1:].split()[0].decode('ascii')
                    chunks = []
                else:
                    chunks.append(line.strip())
        if name is not None:
            self.seqs[name] = b''.join(chunks)

    def fetch(self, chrom, start, end):
        s = self.seqs[chrom]
        start = max(0, start)
        end = min(len(s), end)
        if end <= start:
            return b''
        return s[start:end]

    def contig(self, chrom):
        return self.seqs.get(chrom, b'')
