"""Accuracy of the kernels' own transcendental building blocks (ecckd_b200/csrc/trans_core.cuh):
1 - exp(-kd*tau) of the transmission averaging methods (average_optical_depth.cpp:49-61) and the fused
Planck weight (planck_function.cpp:47-53), swept on the GPU through the ecckd_debug_eval hook."""
import numpy as np
import pytest

import ecckd_b200 as E

pytestmark = pytest.mark.gpu

KD = [1.66, 2 * 1.66, 3 * 1.66, 10 * 1.66]  # transmission, -2, -3, -10


# 0: 32-entry lane table, degree 6; 2: 256-entry shared table, degree 4; 3: replicated 64-entry table, degree 5
@pytest.mark.parametrize("which", [0, 2, 3])
@pytest.mark.parametrize("kd", KD)
def test_one_minus_exp_absolute_error(ctx, kd, which):
    rng = np.random.default_rng(int(kd * 100))
    tau = np.concatenate([10.0 ** rng.uniform(-14, 3.5, 200000), [0.0, 5e-324, 1e-300, 1e-20, 426.0 / kd * 1.66, 1e6, 1e30]])
    got = E.debug_eval(ctx, which, kd, tau)
    ref = -np.expm1(-kd * tau.astype(np.longdouble)).astype(np.float64)
    # the reference's 1.0 - exp(x) carries up to ~1.7e-16 (exp rounded at ulp(1)/2, the difference once more);
    # the table scheme with its two final FMAs stays within 2 ulp of 1
    assert np.max(np.abs(got - ref)) <= 2.3e-16
    assert np.all(got >= 0.0) and np.all(got <= 1.0)
    assert got[-1] == 1.0 and got[-2] == 1.0  # huge optical depths saturate exactly (clamped exponent)
    assert got[len(tau) - 7] == 0.0            # tau = 0


@pytest.mark.parametrize("which", [0, 2, 3])
def test_one_minus_exp_is_monotone_on_a_fine_grid(ctx, which):
    tau = np.linspace(0.0, 30.0, 400001)
    got = E.debug_eval(ctx, which, 1.66, tau)
    assert np.all(np.diff(got) >= -2.3e-16)


@pytest.mark.parametrize("temperature", [120.0, 216.5, 288.15, 350.0])
def test_fused_planck_weight_relative_error(ctx, O, temperature):
    nu = np.concatenate([np.linspace(0.0002, 3260.0, 100001), np.linspace(250.0, 50000.0, 50001)])
    got = E.debug_eval(ctx, 1, temperature, nu)
    ref = O.planck_function(np.array([temperature]), nu, np.ones_like(nu))[0]
    ok = ref > 1e-300
    rel = np.abs(got[ok] - ref[ok]) / ref[ok]
    # Two correct evaluations differ by the conditioning of the formula, not by a few ulp: the exponent
    # x = (h/k) freq / T is rounded (relative change eps*x of exp(x)), and exp(x) - 1 cancels for x -> 0
    # (relative change eps/x).  Inside that envelope the weights agree.
    x = (6.62606896e-34 / 1.3806504e-23) * (nu[ok] * 100.0 * 2.99792458e8) / temperature
    eps = 2.3e-16
    assert np.all(rel < 4 * eps * (1.0 + x) + 4 * eps / np.minimum(x, 1.0))
