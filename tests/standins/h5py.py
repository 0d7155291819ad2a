"""TEST INFRASTRUCTURE ONLY -- API-subset stand-in for h5py (not installed in this image, no network).

Covers exactly what the reference touches: ``File(path, 'r'|'w'|'a')`` as a context manager,
iteration / ``keys()``, ``create_dataset(name, shape, dtype='f', chunks=True)``, ``f[name][...]`` get / set
(seistorch/io.py:17-25,75-80,97-115,150-162,194-199, dataset.py:33-52).  One file = one ``.npz`` archive
written on close, so the unmodified ``OBSDataset`` / ``SeisIO`` of the reference run against it."""
import os

import numpy as np


class _Dataset:
    def __init__(self, arr):
        self._a = arr

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def __getitem__(self, k):
        return self._a[k]

    def __setitem__(self, k, v):
        self._a[k] = v

    def __len__(self):
        return len(self._a)


class File:
    def __init__(self, path, mode="r", **kwargs):
        self.filename, self.mode = path, mode
        self._d = {}
        if mode in ("r", "a", "r+"):
            if os.path.exists(path):
                with np.load(path, allow_pickle=False) as z:
                    self._d = {k: _Dataset(z[k].copy()) for k in z.files}
            elif mode != "a":
                raise FileNotFoundError(path)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self.mode != "r":
            with open(self.filename, "wb") as f:
                np.savez(f, **{k: v._a for k, v in self._d.items()})

    def keys(self):
        return self._d.keys()

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __contains__(self, k):
        return k in self._d

    def __getitem__(self, k):
        return self._d[k]

    def create_dataset(self, name, shape=None, dtype="f", data=None, chunks=None, **kwargs):
        dt = np.float32 if dtype in ("f", "f4", "float32", None) else np.dtype(dtype)
        arr = np.array(data, dtype=dt) if data is not None else np.zeros(shape, dtype=dt)
        self._d[name] = _Dataset(arr)
        return self._d[name]
