"""GPU tests of the drop-in: the UNMODIFIED reference package (oracle/_ref, vendored by oracle/vendor_ref.py) driven through
its own public API -- `Ipeps(config)`, `ipeps.load`, `ipeps.renormalize`, `ipeps.measure`, `ipeps.evolve` -- with
`device="cuda"` and `evolution.backend="b200"` after `acetn_b200.integration.install()`.

  * the reference's own pin, tests/integration/test_ground_states.py:24-35, verbatim (known-answer energy rel 1e-10 after
    load, rel 1e-4 after `evolve(0.01, 10)`) -- on libacetn_b200.so instead of torch;
  * `evolve` issues ZERO CTMRG moves on the reference torch mover (fast_full_update.py:72-129 routed, VERDICT r01 Missing #1)
    and the library's launch counter moves;
  * the README quickstart (BASELINE config 1: Heisenberg 2x2, D=2, chi=20, evolve(0.01, 100) + measure) against the reference
    torch path on the same GPU with the same seed (same init noise, same Omega stream).
"""
import os

import pytest
import torch

from oracle import vendor_ref
from tests.dropin_util import CASES, MoveCounter, config, energies, ref_dir, setup

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(vendor_ref.import_path() is None, reason="reference package not available (oracle/_ref)")]


def _cuda_cfg(cfg, b200=True):
    cfg = dict(cfg)
    cfg["device"] = "cuda"
    cfg["evolution"] = dict(cfg.get("evolution", {}), backend="b200" if b200 else "torch", disable_progressbar=True)
    cfg["ctmrg"] = dict(cfg.get("ctmrg", {}), disable_progressbar=True)
    return cfg


@pytest.mark.parametrize("case", CASES)
def test_reference_ground_state_pins_on_b200(case):
    """tests/integration/test_ground_states.py:24-35, verbatim, with backend='b200' on the GPU."""
    from acetn_b200 import ops
    Ipeps = setup()
    torch.manual_seed(0)
    ipeps = Ipeps(_cuda_cfg(config(case)))
    assert ipeps.config.evolution.backend == "b200"
    ipeps.load(os.path.join(ref_dir(), "ipeps_gs", case + ".pt"))
    assert ipeps.site_states_initialized, "site_states_initialized not true after loading tensors"
    converged_energy = energies()[case]
    ops.reset_launch_count()
    with MoveCounter() as ref_moves:
        measured_energy = ipeps.measure()['Energy']
        assert measured_energy.item() == pytest.approx(converged_energy, rel=1e-10)
        n_measure = ops.launch_count()
        assert n_measure > 0, "measure did not launch any kernel of libacetn_b200.so"
        ipeps.evolve(dtau=0.01, steps=10)
        measured_energy = ipeps.measure()['Energy']
        assert measured_energy.item() == pytest.approx(converged_energy, rel=1e-4)
        assert ref_moves.calls == 0, "evolve / renormalize issued CTMRG moves on the reference torch path"
    assert ops.launch_count() > 10 * n_measure
    assert ipeps[(0, 0)]['C'][0].is_cuda


def test_renormalize_matches_reference_torch_on_same_gpu():
    """`ipeps.renormalize()` (40 sweeps) from the reference's converged Ising state on both backends, same seed: energy and
    order parameters agree to the north-star tolerance (1e-9)."""
    Ipeps = setup()
    case = CASES[0]
    out = {}
    for b200 in (False, True):
        torch.manual_seed(5)
        ip = Ipeps(_cuda_cfg(config(case), b200))
        ip.load(os.path.join(ref_dir(), "ipeps_gs", case + ".pt"))
        ip.renormalize()
        out[b200] = {k: float(v) for k, v in ip.measure().items()}
    for k, v in out[False].items():
        assert out[True][k] == pytest.approx(v, abs=1e-9), f"{k}: torch {v} vs b200 {out[True][k]}"


def test_readme_quickstart_on_b200():
    """BASELINE config 1 (README.md:16-39 of the reference): Heisenberg 2x2, D=2, chi=20, `evolve(0.01, 100)` + `measure()`.
    Checked against the reference torch path on the same GPU (same seed).  The two runs differ only by the orthonormal bases
    of the rSVD / QR steps, which the 800 bond updates amplify to ~1e-8 (SURVEY.md App. E3 saw 7.8e-9 seed-to-seed); the
    assertion is 1e-6 absolute, and 1e-4 against the survey's CPU known answer (-0.65828131, different Omega device)."""
    from acetn_b200 import ops
    Ipeps = setup()
    base = {"dtype": "float64", "device": "cuda", "TN": {"nx": 2, "ny": 2, "dims": {"phys": 2, "bond": 2, "chi": 20}},
            "model": {"name": "heisenberg", "params": {"J": 1.0}}}
    import time
    res, calls, secs = {}, {}, {}
    for b200 in (False, True):
        torch.manual_seed(0)
        ip = Ipeps(_cuda_cfg(base, b200))
        ops.reset_launch_count()
        with MoveCounter() as ref_moves:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ip.evolve(dtau=0.01, steps=100)
            torch.cuda.synchronize()
            secs[b200] = time.perf_counter() - t0
            res[b200] = {k: float(v) for k, v in ip.measure().items()}
        calls[b200] = (ref_moves.calls, ops.launch_count())
    print("quickstart torch:", res[False], "b200:", res[True], "calls:", calls, "evolve seconds (incl. first-use warm-up):", secs)
    assert calls[False][0] > 0 and calls[False][1] == 0
    assert calls[True][0] == 0 and calls[True][1] > 0
    assert res[True]["Energy"] == pytest.approx(res[False]["Energy"], abs=1e-6)
    assert res[True]["Energy"] == pytest.approx(-0.6582813104072897, abs=1e-4)
    assert abs(res[True]["sz"]) == pytest.approx(abs(res[False]["sz"]), abs=1e-5)


def test_evolve_with_als_pinv_matches_reference_torch():
    """evolution.als_method = "pinv" (als_solver.py:226-228): routed to acetn_b200.evolution.ALSSolver.solve_pinv (host-driven loop on
    the library's kernels); a short evolution agrees with the reference torch path on the same GPU, same seed."""
    Ipeps = setup()
    base = {"dtype": "float64", "device": "cuda", "TN": {"nx": 2, "ny": 2, "dims": {"phys": 2, "bond": 2, "chi": 12}},
            "model": {"name": "heisenberg", "params": {"J": 1.0}}, "ctmrg": {"steps": 4}, "evolution": {"als_method": "pinv"}}
    res = {}
    for b200 in (False, True):
        torch.manual_seed(3)
        ip = Ipeps(_cuda_cfg(base, b200))
        assert ip.config.evolution.als_method == "pinv"
        ip.evolve(dtau=0.05, steps=5)
        res[b200] = float(ip.measure()["Energy"])
    assert res[True] == pytest.approx(res[False], abs=1e-8)


def test_checkpoint_round_trip_on_b200(tmp_path):
    """`.pt` compatibility (tensor_network.py:94-122): the reference's own save / load with backend='b200' -- the boundary tensors the
    B200 mover produced (incl. arena tensors of the graph path) pickle, reload onto the GPU and measure to the same energy."""
    Ipeps = setup()
    torch.manual_seed(2)
    ip = Ipeps(_cuda_cfg(config(CASES[0]), True))
    ip.load(os.path.join(ref_dir(), "ipeps_gs", CASES[0] + ".pt"))
    ip.renormalize()
    e0 = float(ip.measure()["Energy"])
    path = str(tmp_path / "state.pt")
    ip.save(path)
    ip2 = Ipeps(_cuda_cfg(config(CASES[0]), True))
    ip2.load(path)
    assert ip2[(0, 0)]['C'][0].is_cuda
    assert float(ip2.measure()["Energy"]) == pytest.approx(e0, rel=1e-13)
