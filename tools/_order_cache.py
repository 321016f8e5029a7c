"""Developer tools only: the METIS ordering of a generated test matrix, cached in a file
(default /tmp) so that the child processes of one GPU call do not each spend ~15 s in
METIS_NodeND on the 27-point 100^3 problem.  The cached vector is exactly what
spral_b200.analyse(order=None) computes (spral_ssids_b200_metis_order), so results are
unchanged."""
import hashlib
import os

import numpy as np


def cached_metis_order(n, ptr, row, cache_dir=None):
    from spral_b200 import _lib
    cache_dir = cache_dir or os.environ.get("SPRAL_B200_ORDER_CACHE", "/tmp")
    h = hashlib.sha1()
    h.update(np.int64(n).tobytes()); h.update(np.ascontiguousarray(ptr).tobytes()); h.update(np.ascontiguousarray(row).tobytes())
    path = os.path.join(cache_dir, f"spral_b200_order_{h.hexdigest()[:16]}.npy")
    if os.path.exists(path):
        try:
            order = np.load(path)
            if order.shape == (n,) and order.dtype == np.int32:
                return order
        except Exception:
            pass
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    order = np.zeros(n, dtype=np.int32)
    rc = _lib.load().spral_ssids_b200_metis_order(n, ptr.ctypes.data, row.ctypes.data, order.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"metis_order failed: {rc}")
    try:
        tmp = path + f".{os.getpid()}.tmp.npy"
        np.save(tmp, order)
        os.replace(tmp, path)
    except OSError:
        pass
    return order
