"""GPU: the hand-written device scan and stable radix sort against numpy (bit-exact, integer work)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    import odeb200
    L = odeb200.lib()
    L.dTestScanB200.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_long, C.POINTER(C.c_int)]
    L.dTestSortB200.argtypes = [C.POINTER(C.c_uint), C.POINTER(C.c_int), C.c_long, C.c_int]
    return L


@pytest.mark.parametrize("n", [1, 31, 512, 4096, 4097, 8191, 8192, 8193, 14344, 16384, 16385, 100003, 1 << 20, 5505060, (1 << 24) + 7])
def test_scan_matches_numpy(n):
    L = _lib()
    rs = np.random.RandomState(n % 1000)
    a = rs.randint(0, 9, size=n).astype(np.int32)
    out = np.zeros(n, np.int32)
    tot = C.c_int(0)
    L.dTestScanB200(a.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(C.POINTER(C.c_int)), n, C.byref(tot))
    ref = np.concatenate([[0], np.cumsum(a.astype(np.int64))[:-1]])
    assert tot.value == int(a.sum())
    assert np.array_equal(out.astype(np.int64), ref)


@pytest.mark.parametrize("n,bits", [(1, 8), (33, 8), (512, 10), (513, 16), (2000, 11), (4097, 24), (10006, 24), (32769, 24), (70000, 24), (262149, 24), (786437, 24),
                                    (1048581, 24), (3000001, 10)])
def test_sort_is_stable_and_sorted(n, bits):
    L = _lib()
    rs = np.random.RandomState(n % 977)
    keys = rs.randint(0, 1 << bits, size=n, dtype=np.int64).astype(np.uint32)
    if n > 10:
        keys[-5:] = (1 << bits) - 1  # sentinels, like the big-geom keys
    vals = np.arange(n, dtype=np.int32)
    k2, v2 = keys.copy(), vals.copy()
    L.dTestSortB200(k2.ctypes.data_as(C.POINTER(C.c_uint)), v2.ctypes.data_as(C.POINTER(C.c_int)), n, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order])
