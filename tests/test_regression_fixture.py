"""tests/golden/basic_run_small.npz (tools/make_regression_fixture.py): inputs of one small basic_run star as the
product built them + the product's outputs on a B200.  Without a GPU the CPU oracle must reproduce the stored CUDA
results from the stored inputs (tables to 1e-10, T_eff(t) to 1e-6, same solver counters); with a GPU the CUDA path
must reproduce the stored tables bit for bit."""
import os

import numpy as np
import pytest

from parity_helpers import oracle_evolve_star, oracle_micro_tables

HERE = os.path.dirname(os.path.abspath(__file__))
TABS = ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK")


class FrozenStar:
    """Quacks like dstar_b200.NScool for the oracle helpers: arrays come from the fixture."""

    def __init__(self, d):
        self.d = d

    def get_int(self, name):
        return int(self.d[name])

    def array(self, name):
        return np.array(self.d[name], dtype=np.float64)


@pytest.fixture(scope="module")
def fx():
    return dict(np.load(os.path.join(HERE, "golden", "basic_run_small.npz")))


def test_oracle_reproduces_stored_cuda_results(oracle, fx):
    s = FrozenStar(fx)
    ref = oracle_micro_tables(oracle, s, gaps=("ns", "sfb03", "t72"))
    assert ref["ierr"] == 0
    for k in TABS:
        got = fx["gpu_" + k].reshape(-1, 4); want = ref[k].reshape(-1, 4)
        assert np.max(np.abs(got[:, 0] - want[:, 0])) < 1e-10, k          # d ln x = relative deviation of x
    assert abs(fx["rho_bar"][0] / ref["rho_bar1"] - 1.0) < 1e-12
    tabs = {k: fx["gpu_" + k] for k in TABS}
    tabs["tab_lnT"] = fx["tab_lnT"]
    ev = oracle_evolve_star(oracle, s, fx["lnT0"], tabs=tabs, heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.5, 28.5, 0.1],
                            T_atm=2.4e8, ctrl=dict(turn_on_extra_heating=1))
    assert np.max(np.abs(fx["gpu_Teff_monitor"] / ev["Teff_monitor"] - 1.0)) < 1e-6
    assert np.max(np.abs(fx["gpu_t_monitor"] - ev["t_monitor"])) < 1e-9
    assert np.max(np.abs(fx["gpu_lnT"] - ev["lnT"])) < 1e-6
    np.testing.assert_array_equal(fx["gpu_counters"], ev["counters"])


@pytest.mark.gpu
def test_cuda_reproduces_stored_tables(fx):
    import dstar_b200 as ds
    nz = int(fx["nz"]); nch = int(fx["ncharged"])
    ctx = ds.Context(0, gaps=("ns", "sfb03", "t72"))
    try:
        b = ds.Batch(ctx, np.array([0, nz]), nch)
        rho_bar = fx["rho_bar"].copy()
        b.upload_structure(fx["rho"], rho_bar, fx["P"], fx["P_bar"], fx["dm"], fx["ionic"], fx["ionic_bar"], fx["Yion"],
                           fx["Yion_bar"], fx["Zion"], fx["Aion"])
        b.micro_tables(ds.MicroCtrl.defaults(), 3)
        out = b.download_tables()
        for k in TABS:
            assert np.array_equal(out[k].view(np.uint64), fx["gpu_" + k].view(np.uint64)), k
        assert out["rho_bar1"][0] == fx["rho_bar"][0]
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_species", [40, 120])
def test_large_networks_take_the_global_memory_paths(fx, n_species):
    """Networks larger than the shared-memory budgets of the coefficient pass (20 species for the per-species
    constants, 96 for the abundances and the active-species list): padding the fixture's 18-species network with
    zero-abundance nuclides must not change a single bit of the tables."""
    import dstar_b200 as ds
    nz = int(fx["nz"]); nch = int(fx["ncharged"])
    pad = n_species - nch
    Zion = np.concatenate([fx["Zion"], 28.0 + (np.arange(pad) % 10)])
    Aion = np.concatenate([fx["Aion"], 60.0 + np.arange(pad)])

    def widen(Y):
        out = np.zeros((nz, n_species))
        out[:, :nch] = Y.reshape(nz, nch)
        return out
    ctx = ds.Context(0, gaps=("ns", "sfb03", "t72"))
    try:
        b = ds.Batch(ctx, np.array([0, nz]), n_species)
        b.upload_structure(fx["rho"], fx["rho_bar"].copy(), fx["P"], fx["P_bar"], fx["dm"], fx["ionic"], fx["ionic_bar"],
                           widen(fx["Yion"]), widen(fx["Yion_bar"]), Zion, Aion)
        b.micro_tables(ds.MicroCtrl.defaults(), 3)
        out = b.download_tables()
        for k in TABS:
            assert np.array_equal(out[k].view(np.uint64), fx["gpu_" + k].view(np.uint64)), k
    finally:
        ctx.close()
