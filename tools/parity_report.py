"""Dev tool: oracle-vs-B200 parity numbers (per-projector truncated-spectrum error, energy, corner spectra) per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import linalg
from acetn_b200.ipeps import CTMRGConfig, Ipeps
from acetn_b200.renormalization import DirectionalMover, ctmrg
from oracle import ctmrg_oracle as orc
from tests.util import cell_from_plain, load_golden, model_terms


def to_oracle_cell(ip):
    sites = {s: orc.Site(ip[s]['A'].cpu(), [c.cpu() for c in ip[s]['C']], [e.cpu() for e in ip[s]['E']]) for s in ip.site_list}
    return orc.Cell(ip.nx, ip.ny, ip.dims, sites)


def report(name, cell, nsweep, hb, hs=None, projectors="half-system", tape=None):
    chi = cell.dims["chi"]
    ref = cell.clone()
    t = orc.OmegaTape(tape) if tape is not None else orc.OmegaTape()
    rec = {}
    orc.ctmrg(ref, orc.CtmrgConfig(steps=nsweep, projectors=projectors), omega_fn=t, record=rec)
    ip = Ipeps.from_plain(cell, CTMRGConfig(steps=nsweep, projectors=projectors))
    linalg.set_omega_source(orc.OmegaTape(t.tape))
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.spectra = []
    try:
        ctmrg(ip, ip.ctmrg_config, mover)
    finally:
        linalg.set_omega_source(None)
    got = to_oracle_cell(ip)
    ds = [float((a - b)[:chi].abs().max()) for a, b in zip(rec["spectra"], mover.projector_calculator.spectra)]
    e0 = float(orc.measure(ref, hb, hs)["Energy"])
    e1 = float(orc.measure(got, hb, hs)["Energy"])
    csv = 0.0
    shapes_ok = True
    for s in ref.site_list:
        for k in range(4):
            shapes_ok &= got[s].C[k].shape == ref[s].C[k].shape and got[s].E[k].shape == ref[s].E[k].shape
            if got[s].C[k].shape == ref[s].C[k].shape:
                a = torch.linalg.svdvals(got[s].C[k]); b = torch.linalg.svdvals(ref[s].C[k])
                csv = max(csv, float((a / a[0] - b / b[0]).abs().max()))
    out = {"case": name, "nsweep": nsweep, "dS_per_projector": ds, "dS_first_move": max(ds[:cell.ny]), "dS_max": max(ds),
           "E_ref": e0, "E_b200": e1, "dE": abs(e0 - e1), "dCsv": csv, "shapes_ok": bool(shapes_ok)}
    print(json.dumps({k: v for k, v in out.items() if k != "dS_per_projector"}))
    print("   dS:", " ".join(f"{x:.0e}" for x in ds))
    return out


H = orc.heisenberg_bond_hamiltonian(1.0)
res = []
torch.manual_seed(0)
res.append(report("random D2 chi8", orc.random_cell(2, 2, 2, 8, 2, seed=0), 2, H))
res.append(report("random D3 chi12", orc.random_cell(2, 2, 3, 12, 2, seed=1), 2, H))
res.append(report("random D4 chi16", orc.random_cell(2, 2, 4, 16, 2, seed=5), 1, H))
res.append(report("random D2 chi8 full-system", orc.random_cell(2, 2, 2, 8, 2, seed=4), 1, H, projectors="full-system"))
res.append(report("product D2 chi10", orc.product_cell(2, 2, 2, 10, 2, seed=2, state_map=lambda s: [1., 0.] if (s[0] + s[1]) % 2 == 0 else [0., 1.]), 3, H))
res.append(report("product D3 chi18", orc.product_cell(2, 2, 3, 18, 2, seed=3), 3, H))
for nm in ["gs_ising_D2_chi20.pt", "gs_heisenberg_D3_chi16.pt"]:
    st = load_golden(nm)
    hb, hs, _ = model_terms(st["model"])
    res.append(report(nm, cell_from_plain(st), 2, hb, hs, tape=st["omega_tape_2sweeps"]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/parity_report.json", "w"), indent=1)
