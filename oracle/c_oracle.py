"""ctypes loader for oracle/libmsda_oracle.so (the plain-C restatement, msda_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of msda_oracle.c.  numpy in, numpy out."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmsda_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "libmsda_oracle.so"])
        _LIB = ctypes.CDLL(path)
        _LIB.msda_oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, attn):
    dt = np.asarray(value).dtype
    assert dt in (np.float32, np.float64)
    value = np.ascontiguousarray(value, dtype=dt)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    n, s, m, d = value.shape
    _, lq, _, nl, p, _ = loc.shape
    dims = [ctypes.c_int(x) for x in (n, s, m, d, nl, lq, p)]
    return dt, value, shapes, lsi, loc, attn, dims, (n, s, m, d, nl, lq, p)


def set_threads(k):
    lib().msda_oracle_set_threads(ctypes.c_int(int(k)))


def num_threads():
    return int(lib().msda_oracle_num_threads())


def forward(value, shapes, lsi, loc, attn):
    dt, value, shapes, lsi, loc, attn, dims, (n, s, m, d, nl, lq, p) = _prep(value, shapes, lsi, loc, attn)
    out = np.empty((n, lq, m * d), dtype=dt)
    fn = getattr(lib(), "msda_oracle_forward_" + ("f32" if dt == np.float32 else "f64"))
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), *dims, _ptr(out))
    return out


def backward(value, shapes, lsi, loc, attn, grad_out):
    dt, value, shapes, lsi, loc, attn, dims, (n, s, m, d, nl, lq, p) = _prep(value, shapes, lsi, loc, attn)
    grad_out = np.ascontiguousarray(grad_out, dtype=dt)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = getattr(lib(), "msda_oracle_backward_" + ("f32" if dt == np.float32 else "f64"))
    fn(_ptr(grad_out), _ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), *dims,
       _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga
