"""CPU checks of the transition-data side: the oracle restatements of
training_data.hflip / rotate / augment / get_discounted_return against fixtures produced by
the UNMODIFIED reference (tests/golden/make_golden_data.py), and the product's HOST-side CSV
reader/writer (g2048_csv_*, no GPU involved) byte for byte against files the reference's
export_csv wrote."""
import ctypes as C
import os

import numpy as np
import pytest

import gym_2048_b200 as g
from conftest import GOLDEN, load_golden
from oracle import oracle


@pytest.fixture(scope="module")
def gold():
    return load_golden("transitions.npz")


@pytest.mark.parametrize("name", ["csv", "syn"])
def test_oracle_symmetries_match_reference(gold, name):
    x, y = gold[name + "/in_x"], gold[name + "/in_y"]
    for tag, h, k in (("hflip", 1, 0), ("rot1", 0, 1), ("rot2", 0, 2), ("rot3", 0, 3), ("hflip_rot3", 1, 3)):
        ob, oa = oracle.symmetry(x, y, h, k)
        assert np.array_equal(ob, gold["%s/%s_x" % (name, tag)]), tag
        assert np.array_equal(oa, gold["%s/%s_y" % (name, tag)]), tag
        nb, _ = oracle.symmetry(gold[name + "/in_next_x"], None, h, k)
        assert np.array_equal(nb, gold["%s/%s_next_x" % (name, tag)]), tag


@pytest.mark.parametrize("name", ["csv", "syn"])
def test_oracle_augment_matches_reference(gold, name):
    o = oracle.augment(gold[name + "/in_x"], gold[name + "/in_next_x"], gold[name + "/in_y"],
                       gold[name + "/in_reward"].astype(np.float32), gold[name + "/in_done"])
    assert np.array_equal(o["boards"], gold[name + "/aug_x"])
    assert np.array_equal(o["next_boards"], gold[name + "/aug_next_x"])
    assert np.array_equal(o["actions"], gold[name + "/aug_y"])
    assert np.array_equal(o["rewards"].astype(np.float64), gold[name + "/aug_reward"])
    assert np.array_equal(o["dones"], gold[name + "/aug_done"])


@pytest.mark.parametrize("name", ["csv", "syn"])
def test_oracle_discounted_return_matches_reference_bit_for_bit(gold, name):
    r, d = gold[name + "/in_reward"].astype(np.float32), gold[name + "/in_done"]
    assert np.array_equal(r.astype(np.float64), gold[name + "/in_reward"])      # rewards are exact in f32
    assert np.array_equal(oracle.discounted_return(r, d, 0.9), gold[name + "/ret_090"])
    assert np.array_equal(oracle.discounted_return(r, d, 0.99), gold[name + "/ret_099"])


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _export(path, gold, returns, append=0, lo=0, hi=None):
    L = g._lib.lib()
    sl = slice(lo, hi)
    b = np.ascontiguousarray(gold["syn/in_x"][sl])
    nb = np.ascontiguousarray(gold["syn/in_next_x"][sl])
    a = np.ascontiguousarray(gold["syn/in_y"][sl])
    r = np.ascontiguousarray(gold["syn/in_reward"][sl])
    d = np.ascontiguousarray(gold["syn/in_done"][sl])
    ret = None if returns is None else np.ascontiguousarray(returns[sl])
    rc = L.g2048_csv_export(path.encode(), _p(b), _p(a), _p(r), _p(nb), _p(d), _p(ret), len(a), append)
    assert rc == 0, L.g2048_last_error()


def test_csv_export_is_byte_identical_to_the_reference_file(gold, tmp_path):
    p = str(tmp_path / "out.csv")
    _export(p, gold, None)
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "ref_export.csv"), "rb").read()
    _export(p, gold, gold["syn/ret_090"])
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "ref_export_returns.csv"), "rb").read()
    # append mode: two halves give the same file
    _export(p, gold, None, hi=120)
    _export(p, gold, None, append=1, lo=120)
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "ref_export.csv"), "rb").read()


@pytest.mark.parametrize("fname,with_ret", [("ref_export.csv", False), ("ref_export_returns.csv", True)])
def test_csv_import_reads_the_reference_file(gold, fname, with_ret):
    L = g._lib.lib()
    path = os.path.join(GOLDEN, fname).encode()
    n, has = C.c_uint64(0), C.c_int(-1)
    assert L.g2048_csv_rows(path, C.byref(n), C.byref(has)) == 0
    assert n.value == 300 and has.value == int(with_ret)
    n = n.value
    b, nb = np.zeros((n, 16), np.uint8), np.zeros((n, 16), np.uint8)
    a, d, r, ret = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n), np.zeros(n)
    rc = L.g2048_csv_import(path, _p(b), _p(a), _p(r), _p(nb), _p(d), _p(ret) if with_ret else None, n)
    assert rc == 0, L.g2048_last_error()
    assert np.array_equal(b, gold["syn/in_x"]) and np.array_equal(nb, gold["syn/in_next_x"])
    assert np.array_equal(a, gold["syn/in_y"]) and np.array_equal(d, gold["syn/in_done"])
    assert np.array_equal(r, gold["syn/in_reward"])
    if with_ret:
        assert np.allclose(ret, gold["syn/ret_090"], rtol=0, atol=5e-7)        # '%f' keeps 6 decimals


def test_csv_errors_are_reported_not_thrown(tmp_path):
    L = g._lib.lib()
    n, has = C.c_uint64(0), C.c_int(0)
    assert L.g2048_csv_rows(str(tmp_path / "missing.csv").encode(), C.byref(n), C.byref(has)) == -1
    assert b"cannot open" in L.g2048_last_error()
    bad = tmp_path / "bad.csv"
    head = open(os.path.join(GOLDEN, "ref_export.csv")).readline()
    bad.write_text(head + ",".join(["3"] * 16 + ["0", "1.0"] + ["0"] * 16 + ["0"]) + "\n")
    assert L.g2048_csv_rows(str(bad).encode(), C.byref(n), C.byref(has)) == 0 and n.value == 1
    z16, z1, zr = np.zeros((1, 16), np.uint8), np.zeros(1, np.uint8), np.zeros(1)
    rc = L.g2048_csv_import(str(bad).encode(), _p(z16), _p(z1), _p(zr), _p(z16.copy()), _p(z1.copy()), None, 1)
    assert rc == -1 and b"not 0 or a power of two" in L.g2048_last_error()
    short = tmp_path / "short.csv"
    short.write_text("a,b,c\n1,2,3\n")
    assert L.g2048_csv_rows(str(short).encode(), C.byref(n), C.byref(has)) == -1


def test_csv_number_forms_round_trip(tmp_path):
    """The importer's own number parsing (integers by hand, '%f' of integral values by hand, everything else
    through strtod): fractional, negative, huge and exponent-form rewards, signs and '\r\n' line ends."""
    L = g._lib.lib()
    rewards = np.array([0.0, 4.0, -1.0, -0.5, 0.25, 2064.0, 1e15, 123456789012345.0, 3.999999, 1e-6, -123456.789012])
    n = len(rewards)
    rng = np.random.default_rng(2)
    b = rng.integers(0, 18, (n, 16)).astype(np.uint8)
    nb = rng.integers(0, 18, (n, 16)).astype(np.uint8)
    a = rng.integers(0, 4, n).astype(np.uint8)
    d = rng.integers(0, 2, n).astype(np.uint8)
    path = str(tmp_path / "forms.csv")
    assert L.g2048_csv_export(path.encode(), _p(b), _p(a), _p(rewards), _p(nb), _p(d), None, n, 0) == 0
    txt = open(path).read()
    assert txt.splitlines()[1:] == [("%d," * 17 + "%f," + "%d," * 16 + "%i") %
                                    (*[(1 << int(e)) if e else 0 for e in b[i]], a[i], rewards[i],
                                     *[(1 << int(e)) if e else 0 for e in nb[i]], d[i]) for i in range(n)]
    # hand-edited forms a spreadsheet or another writer would produce
    lines = txt.splitlines()
    lines[2] = lines[2].replace(",4.000000,", ",+4.0e0,")
    lines[3] = lines[3].replace(",-1.000000,", ",-1,")
    open(path, "w").write("\r\n".join(lines) + "\r\n")
    b2, nb2 = np.zeros_like(b), np.zeros_like(nb)
    a2, d2, r2 = np.zeros_like(a), np.zeros_like(d), np.zeros(n)
    cnt, has = C.c_uint64(0), C.c_int(0)
    assert L.g2048_csv_rows(path.encode(), C.byref(cnt), C.byref(has)) == 0 and cnt.value == n
    rc = L.g2048_csv_import(path.encode(), _p(b2), _p(a2), _p(r2), _p(nb2), _p(d2), None, n)
    assert rc == 0, L.g2048_last_error()
    assert np.array_equal(b2, b) and np.array_equal(nb2, nb) and np.array_equal(a2, a) and np.array_equal(d2, d)
    assert np.array_equal(r2, np.array([float("%f" % v) for v in rewards]))


def test_oracle_sample_actions_is_uniform_over_the_allowed_set():
    n = 200000
    masks = np.random.default_rng(3).integers(0, 16, n).astype(np.uint8)
    a = oracle.sample_actions(masks, n, 7, 42, 5)
    allowed = np.where(masks == 0, 15, masks)
    assert np.all((allowed >> a) & 1 == 1)
    sel = masks == 0b1010                                    # Right and Left only
    frac = np.mean(a[sel] == 1)
    assert abs(frac - 0.5) < 0.02
    u = oracle.sample_actions(None, n, 7, 42, 5)
    assert np.allclose(np.bincount(u, minlength=4) / n, 0.25, atol=0.01)
    # sharding invariance: a slice of the batch draws the same actions
    assert np.array_equal(oracle.sample_actions(masks[1000:2000], 1000, 7 + 1000, 42, 5), a[1000:2000])
