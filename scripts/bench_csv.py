#!/usr/bin/env python
"""The transition CSV either side of the step: the library's host-side export / import
(g2048_csv_export / g2048_csv_import, no GPU involved) timed beside the reference's own numpy code
(`training_data.export_csv` / `import_csv`, /root/reference) on the same rows, and the two files compared
byte for byte.  Runs in the build container (needs /root/reference):

    python scripts/bench_csv.py [rows]
"""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("G2048_REFERENCE", "/root/reference"))

from gym_2048_b200._lib import check, lib  # noqa: E402
from oracle import oracle  # noqa: E402  (only to produce realistic boards)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    import training_data as td                                   # the reference
    envs, steps = 2048, -(-n // 2048)
    batch = oracle.OracleBatch(envs, seed=5, threads=8)
    batch.reset()
    rng = np.random.default_rng(1)
    B, A, R, NB, D = [], [], [], [], []
    for _ in range(steps):
        act = rng.integers(0, 4, envs).astype(np.uint8)
        B.append(batch.boards.copy())
        o = batch.step(act)
        A.append(act); R.append(o["rewards"].astype(np.float64)); NB.append(batch.boards.copy()); D.append(o["dones"])
    b = np.concatenate(B)[:n]; a = np.concatenate(A)[:n]; r = np.concatenate(R)[:n]
    nb = np.concatenate(NB)[:n]; d = np.concatenate(D)[:n]
    L = lib()
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    tmp = tempfile.mkdtemp(prefix="g2048_csv_")
    ours, ref = os.path.join(tmp, "ours.csv"), os.path.join(tmp, "ref.csv")

    t0 = time.perf_counter()
    check(L.g2048_csv_export(ours.encode(), p(b), p(a), p(r), p(nb), p(d), None, n, 0))
    t_exp = time.perf_counter() - t0

    t = td.training_data()
    t._x = oracle.exp_to_values(b).reshape(-1, 4, 4).astype(int)
    t._y_digit = a.reshape(-1, 1).astype(int)
    t._reward = r.reshape(-1, 1)
    t._next_x = oracle.exp_to_values(nb).reshape(-1, 4, 4).astype(int)
    t._done = d.reshape(-1, 1).astype(bool)
    t0 = time.perf_counter()
    t.export_csv(ref)
    t_ref_exp = time.perf_counter() - t0
    same = open(ours, "rb").read() == open(ref, "rb").read()
    size_mb = os.path.getsize(ours) / 1e6

    b2 = np.empty_like(b); a2 = np.empty_like(a); r2 = np.empty_like(r); nb2 = np.empty_like(nb); d2 = np.empty_like(d)
    rows, has_ret = C.c_uint64(), C.c_int()
    t0 = time.perf_counter()
    check(L.g2048_csv_rows(ref.encode(), C.byref(rows), C.byref(has_ret)))
    check(L.g2048_csv_import(ref.encode(), p(b2), p(a2), p(r2), p(nb2), p(d2), None, n))
    t_imp = time.perf_counter() - t0
    ok = rows.value == n and np.array_equal(b2, b) and np.array_equal(a2, a) and np.array_equal(r2, r) and \
        np.array_equal(nb2, nb) and np.array_equal(d2, d)
    u = td.training_data()
    t0 = time.perf_counter()
    u.import_csv(ours)
    t_ref_imp = time.perf_counter() - t0
    print("| %d rows, %.1f MB | library | reference (numpy) | ratio |" % (n, size_mb))
    print("|---|---|---|---|")
    print("| export | %.3f s (%.0f MB/s) | %.2f s | %.0fx |" % (t_exp, size_mb / t_exp, t_ref_exp, t_ref_exp / t_exp))
    print("| import | %.3f s (%.0f MB/s) | %.2f s | %.0fx |" % (t_imp, size_mb / t_imp, t_ref_imp, t_ref_imp / t_imp))
    print("files byte-identical: %s; import round trip exact: %s" % (same, ok))
    for f in (ours, ref):
        os.remove(f)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
