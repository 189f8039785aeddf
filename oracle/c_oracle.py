"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/_build/liboracle.so (oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_mt19937_next.restype = ctypes.c_uint32
    return _LIB


def gae(rewards, terminations, truncations, values, next_values, gamma=0.99, gae_lambda=0.95):
    arrs = [np.ascontiguousarray(x, dtype=np.float32) for x in (rewards, terminations, truncations, values, next_values)]
    T, N = arrs[0].shape
    adv = np.empty((T, N), np.float32)
    ret = np.empty((T, N), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    lib().oracle_gae_f32(*[a.ctypes.data_as(fp) for a in arrs], adv.ctypes.data_as(fp), ret.ctypes.data_as(fp),
                         ctypes.c_int(T), ctypes.c_int(N), ctypes.c_double(gamma), ctypes.c_double(gae_lambda))
    return adv, ret


class MT:
    def __init__(self, seed):
        self.buf = ctypes.create_string_buffer(lib().oracle_mt_state_size())
        lib().oracle_mt19937_seed(self.buf, ctypes.c_uint32(seed))

    def next_u32(self):
        return int(lib().oracle_mt19937_next(self.buf))

    def permutation(self, n):
        out = np.empty(n, np.int64)
        lib().oracle_permutation(self.buf, ctypes.c_int64(n), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        return out
