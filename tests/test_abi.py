"""The C-ABI library loads and exports every symbol include/ecckd_b200.h declares
(no compute calls: this file runs without a GPU)."""
import ctypes
import os

import pytest

from ecckd_b200 import _lib


def test_library_exists_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    L = _lib.lib()
    assert L.ecckd_version() == 100


def test_exports_every_header_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _lib.header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_method_strings():
    L = _lib.lib()
    names = ["linear", "transmission", "transmission-2", "transmission-3", "transmission-10", "square-root",
             "logarithmic", "hybrid-logarithmic-transmission-3"]
    for i, n in enumerate(names):
        assert L.ecckd_averaging_method_from_string(n.encode()) == i
    # known to the interval fits of find_g_points only (find_g_points.cpp:146); the averaging entry points reject
    # id 8 with PARAMETER_ERROR exactly like any unknown name (tests/test_reduce_gpu.py)
    assert L.ecckd_averaging_method_from_string(b"total-transmission") == 8
    assert L.ecckd_averaging_method_from_string(b"nonsense") == -1


def test_overlap_g_points_is_host_only(O):
    """ecckd_overlap_g_points is integer host logic: it must run without a device
    and match the oracle (single_gas_data.cpp:23-124)."""
    import numpy as np

    import ecckd_b200 as E

    rng = np.random.default_rng(4)
    for ngas, nband in [(1, 1), (2, 1), (3, 4), (5, 13)]:
        n_g = rng.integers(1, 6, size=(ngas, nband)).astype(np.int32)
        sv = [np.sort(rng.uniform(0, 10, n_g[a].sum())) for a in range(ngas)]
        ng, gmin, gmax, band = E.overlap_g_points(n_g, sv)
        ng_o, gmin_o, gmax_o, band_o = O.overlap_g_points(n_g, sv)
        assert ng == ng_o == int((1 - ngas) * nband + n_g.sum())  # Eq. 7 of Hogan (2010)
        assert np.array_equal(gmin, gmin_o) and np.array_equal(gmax, gmax_o) and np.array_equal(band, band_o)


@pytest.mark.skipif(_lib.lib().ecckd_device_count() > 0, reason="a B200 is present")
def test_no_cpu_fallback_without_gpu():
    """Without a device the product path must fail loudly, never compute on the CPU."""
    import ecckd_b200 as E

    with pytest.raises(E.EcckdError) as ei:
        E.Context(0)
    assert ei.value.code == 131
    assert "no CPU fallback" in str(ei.value)


def test_cxx_consumer_library_resolves_every_symbol():
    """oracle/_ref/libecckd_ref_tools_gpu.so - the reference's tool mains linked against
    integration/ecckd_gpu_bindings.cpp and libecckd_b200.so - loads with every symbol bound (RTLD_NOW), no GPU needed."""
    import ctypes
    import os

    from tests import ref_tools as RT

    if not RT.gpu_available():
        pytest.skip("oracle/_ref/libecckd_ref_tools_gpu.so not built")
    lib = ctypes.CDLL(RT.GPU_LIB, mode=os.RTLD_NOW)
    assert lib.ref_tool_run is not None


def test_scale_lut_scaling_matches_the_tool_formulas():
    """scale_lut.cpp:117-133, 180-181 (host-sized arithmetic: no device needed by the call itself)."""
    import numpy as np

    import ecckd_b200 as E

    rng = np.random.default_rng(11)
    nz, ng, mu0 = 12, 7, 0.6
    flux = np.cumprod(rng.uniform(0.3, 1.0, size=(nz + 1, ng)), axis=0)
    flux[9:, 2] = 0.0  # the beam is extinguished in g point 2
    od_total = rng.uniform(0.1, 3.0, size=(nz, ng))
    best, scaling = E.scale_lut_scaling(mu0, flux, od_total)
    ref = np.full((nz, ng), -1.0)
    for ig in range(ng):
        top = flux[0, ig]
        for iz in range(nz):
            base = flux[iz + 1, ig]
            with np.errstate(all="ignore"):
                ref[iz, ig] = -1.0 if base <= 0.0 else -mu0 * np.log(base / top)
            top = base
    # (0/0 after extinction is NaN in the tool as well: log(0/0); compare where defined)
    ok = np.isfinite(ref)
    np.testing.assert_allclose(best[ok], ref[ok], rtol=1e-15)
    ref_s = np.where(ref <= 0.0, 1.0, ref / od_total)
    np.testing.assert_allclose(scaling[ok], ref_s[ok], rtol=1e-15)
