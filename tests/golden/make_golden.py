"""
Generates the committed golden fixtures under tests/golden/ by IMPORTING THE REFERENCE
(ace-tn at /root/reference) in the build container.  The reference cannot travel to the
GPU box, so its outputs are frozen here as plain tensors (torch.save, weights_only-safe).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Produces
  gs_ising_D2_chi20.pt, gs_heisenberg_D3_chi16.pt
      the reference's converged ground states (tests/integration/ipeps_gs/*.pt, which are
      pickles of reference classes) re-saved as plain tensors, plus the known-answer energy
      of tests/integration/ipeps_gs/energies.csv and measure() of the reference itself.
  ref_vectors.pt
      seeded reference runs on small shapes: quarter tensors, rSVD spectra with the recorded
      Omega tape, projectors, absorbed C/E, post-sweep corner spectra, RDMs and energies.
"""
import csv
import os
import sys

sys.dont_write_bytecode = True
REF = os.environ.get("ACETN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import torch  # noqa: E402

import acetn  # noqa: E402,F401
from acetn.ipeps import Ipeps  # noqa: E402
from acetn.ipeps.bond import Bond  # noqa: E402
from acetn.measurement.rdm import RDM  # noqa: E402
from acetn.renormalization.ctmrg import ctmrg  # noqa: E402
from acetn.renormalization.directional_mover import DirectionalMover  # noqa: E402
from acetn.renormalization.projectors import ProjectorCalculator  # noqa: E402

from oracle import ctmrg_oracle as orc  # noqa: E402  (only for the shared synthetic-input generator)


def plain_state(ipeps):
    out = {"nx": ipeps.nx, "ny": ipeps.ny, "dims": dict(ipeps.dims), "sites": {}}
    for site in ipeps.site_list:
        st = ipeps[site]
        out["sites"][f"{site[0]},{site[1]}"] = {
            "A": st["A"].detach().clone().cpu(),
            "C": [c.detach().clone().cpu() for c in st["C"]],
            "E": [e.detach().clone().cpu() for e in st["E"]],
        }
    return out


def make_ipeps(nx, ny, D, chi, d, model="heisenberg", params=None, ctm=None):
    cfg = {
        "dtype": "float64", "device": "cpu",
        "TN": {"nx": nx, "ny": ny, "dims": {"phys": d, "bond": D, "chi": chi}},
        "model": {"name": model, "params": params or {"J": 1.0}},
        "ctmrg": dict({"steps": 1, "projectors": "half-system", "disable_progressbar": True}, **(ctm or {})),
    }
    return Ipeps(cfg)


def load_cell_into(ipeps, cell):
    """Overwrite the reference's site tensors with the oracle-side synthetic cell (same numbers)."""
    for site in ipeps.site_list:
        s = cell[site]
        ipeps[site]["A"] = s.A
        ipeps[site]["C"] = s.C
        ipeps[site]["E"] = s.E


class RandnRecorder:
    """Wraps torch.randn to record the Omega draws of the reference's randomized SVD."""

    def __init__(self):
        self.tape = []
        self._orig = torch.randn

    def __enter__(self):
        def rec(*a, **k):
            t = self._orig(*a, **k)
            self.tape.append(t.detach().clone())
            return t
        torch.randn = rec
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig


def golden_ground_states():
    energies = {}
    with open(os.path.join(REF, "tests/integration/ipeps_gs/energies.csv"), newline="") as f:
        for row in csv.reader(f):
            energies[row[0]] = float(row[1])
    import toml
    for case, out in [("ising_dims_2_20_dtau_001_hx_295", "gs_ising_D2_chi20.pt"),
                      ("heisenberg_dims_3_16_dtau_001", "gs_heisenberg_D3_chi16.pt")]:
        cfg = toml.load(os.path.join(REF, "tests/integration/input", case + ".toml"))
        ip = Ipeps(cfg)
        ip.load(os.path.join(REF, "tests/integration/ipeps_gs", case + ".pt"))
        meas = ip.measure()
        st = plain_state(ip)
        st["energy_csv"] = energies[case]
        st["reference_measure"] = {k: float(v) for k, v in meas.items()}
        st["model"] = cfg["model"]
        rdm = RDM(ip)
        st["site_rdm_00"] = rdm[(0, 0)].clone()
        st["bond_rdm_0"] = rdm[ip.bond_list[0]].clone()
        st["bond_rdm_last"] = rdm[ip.bond_list[-1]].clone()
        # 2 further reference sweeps from the converged state, with the Omega tape
        torch.manual_seed(7)
        ip.config.ctmrg.steps = 2
        ip.config.ctmrg.disable_progressbar = True
        with RandnRecorder() as rec:
            ip.renormalize()
        st["omega_tape_2sweeps"] = rec.tape
        st["after2_energy"] = float(ip.measure()["Energy"])
        st["after2_corner_svals"] = {f"{s[0]},{s[1]},{k}": torch.linalg.svdvals(ip[s]["C"][k])
                                     for s in ip.site_list for k in range(4)}
        torch.save(st, os.path.join(HERE, out))
        print(out, "E_csv", energies[case], "E_ref", st["reference_measure"]["Energy"], "E_after2", st["after2_energy"])


def golden_vectors():
    vec = {}
    # ---- single projector + absorption on synthetic random cells -------------------------
    for (D, chi, d, seed) in [(2, 8, 2, 0), (3, 12, 2, 1), (4, 16, 2, 2), (3, 10, 3, 3)]:
        cell = orc.random_cell(2, 2, D, chi, d, seed=seed)
        ip = make_ipeps(2, 2, D, chi, d)
        load_cell_into(ip, cell)
        case = {"D": D, "chi": chi, "d": d, "seed": seed}
        for k in range(4):
            q, shp = ProjectorCalculator.make_quarter_tensor(ip[(0, 0)], k)
            case[f"quarter_k{k}"] = q.clone()
            case[f"quarter_k{k}_shape"] = tuple(shp)
        mover = DirectionalMover(ip.config.ctmrg)
        torch.manual_seed(100 + seed)
        with RandnRecorder() as rec:
            p1, p2 = mover.calculate_left_projectors(ip, 0, 0)
        case["left_proj_omega"] = rec.tape
        case["left_proj1"], case["left_proj2"] = p1.clone(), p2.clone()
        # spectra through the reference's own rSVD on the same Omega
        from acetn.linalg import fused_matmul_svd_lowrank
        Q1, _ = ProjectorCalculator.make_quarter_tensor(ip[(0, 0)], 0)
        Q4, _ = ProjectorCalculator.make_quarter_tensor(ip[(0, 1)], 3)
        orig = torch.randn
        torch.randn = lambda *a, **k: rec.tape[0].clone()
        try:
            U, S, V = fused_matmul_svd_lowrank(Q1, Q4, q=chi + 2, niter=2)
        finally:
            torch.randn = orig
        case["left_rsvd_S"] = S.clone()
        case["left_exact_S"] = torch.linalg.svdvals(Q1 @ Q4)
        for k in range(4):
            st = ip[(0, 0)]
            pj1 = torch.rand(chi, D, D, chi - 1, dtype=torch.float64, generator=torch.Generator().manual_seed(k))
            pj2 = torch.rand(chi, D, D, chi - 1, dtype=torch.float64, generator=torch.Generator().manual_seed(10 + k))
            case[f"absorb_k{k}_proj1"], case[f"absorb_k{k}_proj2"] = pj1, pj2
            case[f"absorb_k{k}_c1"] = DirectionalMover.renormalize_cj1(st["C"][(3 + k) % 4], st["E"][(2 + k) % 4], pj1).clone()
            case[f"absorb_k{k}_c2"] = DirectionalMover.renormalize_cj2(st["C"][k], st["E"][k], pj2).clone()
            case[f"absorb_k{k}_e"] = DirectionalMover.renormalize_ej(st["E"][(3 + k) % 4], st.bond_permute(k), pj2, pj1).clone()
        rdm = RDM(ip)
        case["site_rdm_00"] = rdm[(0, 0)].clone()
        case["bond_rdm_h"] = rdm[ip.bond_list[0]].clone()
        case["bond_rdm_v"] = rdm[ip.bond_list[-1]].clone()
        vec[f"single_D{D}_chi{chi}_d{d}"] = case

    # ---- full sweeps: synthetic random and default product-state starts --------------------
    for (name, D, chi, d, nsweep, seed, kind, nx, ny, proj) in [
        ("sweep_random_D2_chi8", 2, 8, 2, 2, 0, "random", 2, 2, "half-system"),
        ("sweep_random_D3_chi12", 3, 12, 2, 2, 1, "random", 2, 2, "half-system"),
        ("sweep_product_D2_chi10", 2, 10, 2, 3, 2, "product", 2, 2, "half-system"),
        ("sweep_random_D2_chi6_3x2", 2, 6, 2, 2, 3, "random", 3, 2, "half-system"),
        ("sweep_random_D2_chi8_full", 2, 8, 2, 2, 4, "random", 2, 2, "full-system"),
    ]:
        if kind == "random":
            cell = orc.random_cell(nx, ny, D, chi, d, seed=seed)
        else:
            cell = orc.product_cell(nx, ny, D, chi, d, seed=seed,
                                    state_map=lambda s: [1.0, 0.0] if (s[0] + s[1]) % 2 == 0 else [0.0, 1.0])
        ip = make_ipeps(nx, ny, D, chi, d, ctm={"steps": nsweep, "projectors": proj})
        load_cell_into(ip, cell)
        torch.manual_seed(1000 + seed)
        with RandnRecorder() as rec:
            ctmrg(ip, ip.config.ctmrg)
        case = {"D": D, "chi": chi, "d": d, "seed": seed, "kind": kind, "nx": nx, "ny": ny,
                "nsweep": nsweep, "projectors": proj, "omega_tape": rec.tape}
        case["state_after"] = plain_state(ip)
        case["energy_heisenberg"] = float(ip.measure()["Energy"])
        case["corner_svals"] = {f"{s[0]},{s[1]},{k}": torch.linalg.svdvals(ip[s]["C"][k])
                                for s in ip.site_list for k in range(4)}
        rdm = RDM(ip)
        case["site_rdm_00"] = rdm[(0, 0)].clone()
        case["bond_rdm_0"] = rdm[ip.bond_list[0]].clone()
        vec[name] = case
        print(name, "E", case["energy_heisenberg"], "C shape", tuple(ip[(0, 0)]["C"][0].shape))

    # ---- rSVD unit vectors (reference tests/unit/test_linalg.py style) ------------------------
    from acetn.linalg import svd_lowrank, fused_matmul_svd_lowrank, fused_3matmul_svd_lowrank
    g = torch.Generator().manual_seed(5)
    A = torch.randn(60, 40, dtype=torch.float64, generator=g)
    B = torch.randn(40, 50, dtype=torch.float64, generator=g)
    C = torch.randn(50, 30, dtype=torch.float64, generator=g)
    Dm = torch.randn(30, 45, dtype=torch.float64, generator=g)
    lowA = torch.randn(60, 5, dtype=torch.float64, generator=g) @ torch.randn(5, 40, dtype=torch.float64, generator=g)
    case = {"A": A, "B": B, "C": C, "D": Dm, "lowA": lowA}
    for nm, fn, args in [("svd_lowrank", svd_lowrank, (A,)), ("fused2", fused_matmul_svd_lowrank, (A, B)),
                         ("fused3", fused_3matmul_svd_lowrank, (A, B, C, Dm)),
                         ("svd_lowrank_lowA", svd_lowrank, (lowA,)), ("fused2_lowA", fused_matmul_svd_lowrank, (lowA, B))]:
        torch.manual_seed(11)
        with RandnRecorder() as rec:
            U, S, V = fn(*args, q=12, niter=2)
        case[nm] = {"omega": rec.tape[0], "U": U.clone(), "S": S.clone(), "V": V.clone()}
    vec["rsvd_unit"] = case

    torch.save(vec, os.path.join(HERE, "ref_vectors.pt"))
    print("ref_vectors.pt", os.path.getsize(os.path.join(HERE, "ref_vectors.pt")) / 1e6, "MB")


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    golden_ground_states()
    golden_vectors()
