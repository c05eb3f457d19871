#!/usr/bin/env python
"""Host decode A/B (no GPU work): whole-file BAM decode with the own DEFLATE decoder vs zlib (PSS_ZLIB_INFLATE=1), and
both decoders single-threaded over the file's BGZF blocks (pss_inflate_raw modes 1 / 2)."""
import ctypes as C, os, struct, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = '/tmp/decode_bench.bam'
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    from pysamstats_b200 import bamio
    best = 1e9
    for rep in range(5):
        bamio._bam_cache.clear()
        t = time.time()
        b = bamio.open_bam(path).load()
        best = min(best, time.time() - t)
    A = b._load_all_any()
    print('%s: best of 5 whole-file decodes %.3f s, %d reads, %.1f M reads/s' % (
        'zlib' if os.environ.get('PSS_ZLIB_INFLATE') else 'own ', best, A.n_reads, A.n_reads / best / 1e6), flush=True)
    sys.exit(0)
from pysamstats_b200 import bamio, synth
G = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 16_000_000
batch, _ = synth.make_interval_batch(0, G, G, depth=30, device='cpu', with_qual=True)
bamio.write_bam(path, ['chrS'], [G], synth.batch_to_numpy(batch), level=1)
for env in ({}, {'PSS_ZLIB_INFLATE': '1'}, {}, {'PSS_ZLIB_INFLATE': '1'}):
    subprocess.check_call([sys.executable, os.path.abspath(__file__), 'child'], env=dict(os.environ, **env))
lib = C.CDLL(bamio.HOST_LIB_PATH)
lib.pss_inflate_raw.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]
data = open(path, 'rb').read()
blocks, p = [], 0
while p < len(data) and len(blocks) < 2000:
    xlen, = struct.unpack_from('<H', data, p + 10)
    bsize = struct.unpack_from('<H', data, p + 16)[0] + 1
    blocks.append((data[p + 12 + xlen:p + bsize - 8], struct.unpack_from('<I', data, p + bsize - 4)[0]))
    p += bsize
tot = sum(i for _, i in blocks)
out = C.create_string_buffer(65536)
for mode, name in ((1, 'own'), (2, 'zlib'), (1, 'own'), (2, 'zlib')):
    t = time.time()
    for pay, isz in blocks:
        lib.pss_inflate_raw(pay, len(pay), out, isz, mode)
    dt = time.time() - t
    print('%s single thread: %.0f MB/s inflated' % (name, tot / dt / 1e6))
