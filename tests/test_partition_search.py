"""ecckd_partition_search_* (ecckd_b200/csrc/partition_search.cu) against the reference's own Equipartition
(equipartition.cpp:160-805, compiled in place into oracle/_ref/libecckd_ref_tools.so) - host logic, no GPU.

Both are driven by the SAME interval-error function (a Python callable), so bounds, errors, number of
intervals and status have to agree bit for bit, whatever the number of line-search candidates evaluated
speculatively per batch.  The error functions are of the kind find_g_points meets: sums over a DISCRETE
sorted spectrum (bounds snapped to indices, find_g_points.cpp:282-317), non-differentiable in the bounds."""
import numpy as np
import pytest

import ecckd_b200 as E
from ecckd_b200.api import EP_STATUS, PartitionSearch
from tests import ref_tools as RT

pytestmark = pytest.mark.skipif(not RT.available(), reason="oracle/_ref/libecckd_ref_tools.so not built")


@pytest.fixture(scope="module")
def R():
    return RT.ref_tools()


def discrete_error(n, seed, power=2.0, noise=0.0):
    """error of [b1, b2] = (spread of a sorted random spectrum inside the snapped index range) ** power"""
    rng = np.random.default_rng(seed)
    x = np.sort(rng.lognormal(0.0, 2.0, n))
    lx = np.log(x)
    c1 = np.concatenate([[0.0], np.cumsum(lx)])
    c2 = np.concatenate([[0.0], np.cumsum(lx * lx)])
    bump = rng.random(n) * noise

    def f(b1, b2):
        i1 = min(max(int(np.ceil(b1 * (n - 1))), 0), n - 1)
        i2 = min(max(int(np.floor(b2 * (n - 1))), i1), n - 1)
        m = i2 - i1 + 1
        s1, s2 = c1[i2 + 1] - c1[i1], c2[i2 + 1] - c2[i1]
        var = max(s2 - s1 * s1 / m, 0.0)
        return float(var ** (0.5 * power) * (1.0 + bump[i1]) + 1e-9 * m)
    return f


def smooth_error(power):
    def f(b1, b2):
        return float((b2 ** power - b1 ** power) * (1.0 + 0.3 * np.sin(9.0 * b1) * np.sin(5.0 * b2)) + 1e-6)
    return f


def batched(fn):
    return lambda b1, b2: np.array([fn(float(a), float(b)) for a, b in zip(b1, b2)])


CASES = [
    ("smooth2", smooth_error(2.0), {}),
    ("smooth3", smooth_error(3.0), dict(partition_tolerance=0.02, partition_max_iterations=60)),
    ("discrete2k", discrete_error(2000, 1), dict(partition_tolerance=0.02, partition_max_iterations=60, resolution=1 / 2000)),
    ("discrete300", discrete_error(300, 2, 1.0), dict(partition_max_iterations=40, resolution=1 / 300)),
    ("noisy", discrete_error(5000, 3, 2.0, 0.5), dict(partition_tolerance=0.02, partition_max_iterations=60,
                                                        resolution=1 / 5000)),
    ("stddev", discrete_error(1000, 4), dict(minimize_frac_range=0, partition_tolerance=0.01)),
    ("cubic", smooth_error(2.5), dict(cubic_interpolation=1)),
    ("maxit", discrete_error(3000, 6, 2.0, 0.8), dict(partition_tolerance=1e-4, partition_max_iterations=4)),
    ("coarse", discrete_error(50, 8), dict(resolution=0.05, partition_tolerance=1e-3)),
    ("short_ls", discrete_error(4000, 5, 2.0, 0.3), dict(line_search_max_iterations=3, partition_max_iterations=30)),
]


@pytest.mark.parametrize("name,fn,opt", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("ni", [2, 3, 5, 12])
def test_equipartition_n_equals_the_reference(R, name, fn, opt, ni):
    b0 = np.sqrt(np.arange(ni + 1) / ni)  # the start of find_g_points.cpp:1239-1242
    st_ref, b_ref, e_ref, calls = R.equipartition_n_callback(fn, b0, **opt)
    seen = []
    for spec in (1, 4, 10):
        ps = PartitionSearch(batched(fn), speculative_line_search=spec, **opt)
        st, b, e = ps.equipartition_n(b0)
        stats = ps.stats()
        ps.close()
        assert st == EP_STATUS[st_ref], (spec, st, EP_STATUS[st_ref])
        assert np.array_equal(b, b_ref), (spec, b, b_ref)
        assert np.array_equal(e, e_ref), (spec, e, e_ref)
        seen.append(stats)
    # one candidate per batch evaluates exactly the reference's intervals; batching never takes more calls
    assert seen[0]["intervals"] == calls and seen[0]["speculative_unused"] == 0
    assert seen[1]["batches"] <= seen[0]["batches"] and seen[2]["batches"] <= seen[1]["batches"]


@pytest.mark.parametrize("name,fn,opt", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("frac", [0.5, 0.1, 0.02])
def test_equipartition_e_equals_the_reference(R, name, fn, opt, frac):
    target = fn(0.0, 1.0) * frac
    st_ref, b_ref, e_ref, calls = R.equipartition_e_callback(fn, target, 0.0, 1.0, **opt)
    for spec in (1, 4):
        ps = PartitionSearch(batched(fn), speculative_line_search=spec, **opt)
        st, b, e = ps.equipartition_e(target, 0.0, 1.0, capacity=512)
        stats = ps.stats()
        ps.close()
        assert st == EP_STATUS[st_ref]
        assert b.size == b_ref.size and np.array_equal(b, b_ref), (b, b_ref)
        assert np.array_equal(e, e_ref)
        if spec == 1:
            assert stats["intervals"] == calls


def test_given_errors_are_used(R):
    fn = discrete_error(1000, 7)
    b0 = np.linspace(0.0, 1.0, 7)
    e0 = np.array([fn(a, b) for a, b in zip(b0[:-1], b0[1:])])
    st_ref, b_ref, e_ref, _ = R.equipartition_n_callback(fn, b0)
    ps = PartitionSearch(batched(fn), speculative_line_search=1)
    st, b, e = ps.equipartition_n(b0, error=e0)
    assert st == EP_STATUS[st_ref] and np.array_equal(b, b_ref) and np.array_equal(e, e_ref)
    # the first calc_error_all of the reference (6 intervals) was not repeated
    st2, _, _, calls = R.equipartition_n_callback(fn, b0)
    assert ps.stats()["intervals"] == calls - 6
    ps.close()


def test_input_errors_and_evaluator_failure():
    fn = smooth_error(2.0)
    ps = PartitionSearch(batched(fn))
    st, _, _ = ps.equipartition_n([0.0, 0.6, 0.5, 1.0])  # not increasing: EP_INPUT_ERROR (equipartition.cpp:362-367)
    assert st == "input-error"
    st, b, e = ps.equipartition_e(1.0, 1.0, 0.0)
    assert st == "input-error" and e.size == 0
    ps.close()
    with pytest.raises(E.EcckdError):
        PartitionSearch(batched(fn), no_such_option=1)

    def failing(b1, b2):
        raise E.EcckdError(135, "evaluator failed")
    ps = PartitionSearch(failing)
    with pytest.raises(E.EcckdError) as ei:
        ps.equipartition_n(np.linspace(0, 1, 5))
    assert ei.value.code == 135
    ps.close()
