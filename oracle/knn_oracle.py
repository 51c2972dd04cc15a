"""ctypes front-end of the plain-C KNN oracle (oracle/knn_oracle.c) and, when it was built in the
container, of the reference's own compiled CPU KNN (oracle/_ref/knn_ref.so).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref = None


def build(ref=False):
    subprocess.check_call(["make", "-C", _HERE, "all"] + (["ref"] if ref else []), stdout=subprocess.DEVNULL)


def _sig(fn, extra=()):
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                   ctypes.c_void_p, ctypes.c_void_p] + list(extra)
    fn.restype = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libknn_oracle.so")
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
        _sig(_lib.knn_oracle)
        _sig(_lib.knn_oracle_mt, [ctypes.c_int])
    return _lib


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "knn_ref.so"))


def _run(fn, p1, p2, K, *extra):
    p1 = np.ascontiguousarray(p1.detach().cpu().numpy() if torch.is_tensor(p1) else p1, dtype=np.float32)
    p2 = np.ascontiguousarray(p2.detach().cpu().numpy() if torch.is_tensor(p2) else p2, dtype=np.float32)
    n1, D = p1.shape
    idx = np.zeros((n1, K), dtype=np.int64)
    d2 = np.zeros((n1, K), dtype=np.float32)
    fn(p1.ctypes.data, n1, p2.ctypes.data, p2.shape[0], D, K, idx.ctypes.data, d2.ctypes.data, *extra)
    return torch.from_numpy(d2), torch.from_numpy(idx)


def knn_c(p1, p2, K, threads=None):
    """(d2 [N,K] fp32, idx [N,K] int64), ascending by (dist, index)."""
    threads = threads or (os.cpu_count() or 1)
    return _run(lib().knn_oracle_mt, p1, p2, K, int(threads))


def knn_reference(p1, p2, K):
    """The reference's own knn_cpu.cpp (single-threaded), when oracle/_ref/knn_ref.so exists."""
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(os.path.join(_HERE, "_ref", "knn_ref.so"))
        _sig(_ref.knn_ref)
    return _run(_ref.knn_ref, p1, p2, K)
