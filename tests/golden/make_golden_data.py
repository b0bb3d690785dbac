#!/usr/bin/env python
"""Generate tests/golden/transitions.npz and tests/golden/ref_export*.csv by running the
UNMODIFIED reference `training_data.py` (numpy only) in the build container:

    python tests/golden/make_golden_data.py

Everything stored here is computed by the reference's own code (`hflip`, `rotate`, `augment`,
`get_discounted_return`, `export_csv`, `import_csv`); neither the C oracle nor the CUDA
kernels are involved.  Boards are stored as uint8 exponents (value = 2**e).

  transitions.npz
    csv/*      the 848 rows of the reference's data/test_data.csv as loaded by import_csv
    syn/*      300 synthetic rows (exponents 0..17, zero / negative / large rewards, dones)
    for each set S in {csv, syn}: S/hflip_*, S/rot{1,2,3}_*, S/aug_* (x, y, next_x [, reward, done]),
    S/ret_090, S/ret_099 (get_discounted_return)
  ref_export.csv, ref_export_returns.csv   export_csv of the synthetic set (add_returns False / True)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("G2048_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import training_data as td  # noqa: E402  (the reference)


def to_exp(values):
    v = np.asarray(values, dtype=np.int64)
    e = np.zeros(v.shape, np.uint8)
    nz = v > 0
    e[nz] = np.round(np.log2(v[nz])).astype(np.uint8)
    assert np.array_equal(np.where(nz, np.int64(1) << e.astype(np.int64), 0), v)
    return e


def make(x, y, r, nx, d):
    t = td.training_data()
    t._x = np.asarray(x, dtype=int).reshape(-1, 4, 4)
    t._y_digit = np.asarray(y, dtype=int).reshape(-1, 1)
    t._reward = np.asarray(r, dtype=float).reshape(-1, 1)
    t._next_x = np.asarray(nx, dtype=int).reshape(-1, 4, 4)
    t._done = np.asarray(d, dtype=bool).reshape(-1, 1)
    t._check_lengths()
    return t


def record(out, name, t):
    def put(tag, u, full=False):
        out["%s/%s_x" % (name, tag)] = to_exp(u.get_x()).reshape(-1, 16)
        out["%s/%s_y" % (name, tag)] = u.get_y_digit().reshape(-1).astype(np.uint8)
        out["%s/%s_next_x" % (name, tag)] = to_exp(u.get_next_x()).reshape(-1, 16)
        if full:
            out["%s/%s_reward" % (name, tag)] = u.get_reward().reshape(-1).astype(np.float64)
            out["%s/%s_done" % (name, tag)] = u.get_done().reshape(-1).astype(np.uint8)
    put("in", t, full=True)
    u = t.copy(); u.hflip(); put("hflip", u)
    for k in (1, 2, 3):
        u = t.copy(); u.rotate(k); put("rot%d" % k, u)
    u = t.copy(); u.hflip(); u.rotate(3); put("hflip_rot3", u)
    u = t.copy(); u.augment(); put("aug", u, full=True)
    out["%s/ret_090" % name] = t.get_discounted_return().reshape(-1)
    out["%s/ret_099" % name] = t.get_discounted_return(gamma=0.99).reshape(-1)


def main():
    out = {}
    t = td.training_data()
    t.import_csv(os.path.join(REF, "data", "test_data.csv"))
    assert t.size() == 848
    record(out, "csv", t)

    rng = np.random.default_rng(20261017)
    n = 300
    ex = rng.integers(0, 18, (n, 16))
    ex[rng.random((n, 16)) < 0.35] = 0
    nex = rng.integers(0, 18, (n, 16))
    nex[rng.random((n, 16)) < 0.35] = 0
    vals = lambda e: np.where(e > 0, np.int64(1) << e.astype(np.int64), 0)      # noqa: E731
    rew = rng.choice([0.0, 0.0, 4.0, 8.0, 12.0, 2064.0, 131072.0, -1.0, 36.0], n)
    done = rng.random(n) < 0.12
    s = make(vals(ex), rng.integers(0, 4, n), rew, vals(nex), done)
    record(out, "syn", s)
    s.export_csv(os.path.join(HERE, "ref_export.csv"))
    s.export_csv(os.path.join(HERE, "ref_export_returns.csv"), add_returns=True)
    # the reference reads its own file back to the same table
    back = td.training_data()
    back.import_csv(os.path.join(HERE, "ref_export.csv"))
    assert np.array_equal(back.get_x(), s.get_x()) and np.array_equal(back.get_next_x(), s.get_next_x())
    assert np.array_equal(back.get_y_digit(), s.get_y_digit()) and np.array_equal(back.get_done(), s.get_done())
    assert np.allclose(back.get_reward(), s.get_reward())
    np.savez_compressed(os.path.join(HERE, "transitions.npz"), **out)
    print("wrote transitions.npz (%d arrays), ref_export.csv, ref_export_returns.csv" % len(out))


if __name__ == "__main__":
    main()
