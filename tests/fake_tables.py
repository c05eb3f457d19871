"""TEST INFRASTRUCTURE: a minimal in-memory stand-in for the PyTables calls ``write_hdf5`` makes
(``/root/reference/pysamstats/io.py:91-244``: ``open_file`` / ``remove_node`` / ``create_table`` /
``Filters`` / ``append`` / ``flush`` / ``close``, and ``h5file.root.<dataset>.dtype`` as read by the
reference's ``pysamstats/test/test_io.py:14-40``).  PyTables itself cannot be installed here (no
network).  A "file" is a pickle of {(group, dataset): (dtype descr, array, title, filters, chunkshape)}.
"""
import os
import pickle

import numpy as np


class NoSuchNodeError(Exception):
    pass


class Filters(object):
    def __init__(self, complevel=0, complib='zlib', shuffle=True, fletcher32=False):
        self.complevel, self.complib, self.shuffle, self.fletcher32 = complevel, complib, shuffle, fletcher32


class Table(object):
    def __init__(self, dtype, title, filters, chunkshape):
        self.dtype = np.dtype(dtype)
        self.title, self.filters, self.chunkshape = title, filters, tuple(chunkshape)
        self._chunks = []
        self.n_appends = self.n_flushes = 0

    def append(self, rows):
        a = np.asarray(rows, dtype=self.dtype) if not isinstance(rows, np.ndarray) else rows
        assert a.dtype == self.dtype, (a.dtype, self.dtype)
        assert len(a) <= self.chunkshape[0]
        self._chunks.append(np.array(a))
        self.n_appends += 1

    def flush(self):
        self.n_flushes += 1

    def read(self):
        return np.concatenate(self._chunks) if self._chunks else np.zeros(0, self.dtype)

    def __len__(self):
        return sum(len(c) for c in self._chunks)


class _Group(object):
    def __init__(self, f, path):
        self._f, self._path = f, path

    def __getattr__(self, name):
        key = (self._path, name)
        if key in self._f._nodes:
            return self._f._nodes[key]
        sub = (self._path.rstrip('/') + '/' + name)
        if any(k[0] == sub or k[0].startswith(sub + '/') for k in self._f._nodes):
            return _Group(self._f, sub)
        raise NoSuchNodeError(name)


class File(object):
    def __init__(self, path, mode):
        self.path, self.mode = path, mode
        self._nodes = {}
        if mode in ('a', 'r') and os.path.exists(path) and os.path.getsize(path) > 0:
            with open(path, 'rb') as fh:
                for key, (descr, arr, title, filt, chunkshape) in pickle.load(fh).items():
                    t = Table(np.dtype(descr), title, Filters(*filt), chunkshape)
                    t._chunks = [arr]
                    self._nodes[key] = t
        self.root = _Group(self, '/')
        self.closed = False

    def remove_node(self, group, name):
        if (group, name) not in self._nodes:
            raise NoSuchNodeError(name)
        del self._nodes[(group, name)]

    def create_table(self, group, name, description, title='', filters=None, createparents=False, chunkshape=None):
        assert createparents or group == '/'
        t = Table(description, title, filters or Filters(), chunkshape)
        self._nodes[(group, name)] = t
        return t

    def close(self):
        if self.mode != 'r':
            d = {}
            for key, t in self._nodes.items():
                f = t.filters
                descr = t.dtype.descr if t.dtype.names else t.dtype.str
                d[key] = (descr, t.read(), t.title, (f.complevel, f.complib, f.shuffle, f.fletcher32), t.chunkshape)
            with open(self.path, 'wb') as fh:
                pickle.dump(d, fh)
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def open_file(path, mode='r'):
    return File(path, mode)
