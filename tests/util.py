"""Shared helpers for the parity tests (oracle side only; no product code here)."""
import os

import torch

from oracle import ctmrg_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=True)


def cell_from_plain(st):
    sites = {}
    for key, v in st["sites"].items():
        x, y = (int(t) for t in key.split(","))
        sites[(x, y)] = orc.Site(v["A"].clone(), [c.clone() for c in v["C"]], [e.clone() for e in v["E"]])
    return orc.Cell(st["nx"], st["ny"], st["dims"], sites)


def model_terms(model):
    """(bond_ham, site_ham, site_ops) of the two reference models used by the goldens."""
    X, iY, Z, _ = orc.pauli()
    if model["name"] == "heisenberg":
        def ops(site):
            sg = 1.0 if (site[0] + site[1]) % 2 == 0 else -1.0
            return {"sx": sg * X, "sz": sg * Z}
        return orc.heisenberg_bond_hamiltonian(model["params"]["J"]), None, ops
    if model["name"] == "ising":
        hs, hb = orc.ising_hamiltonians(model["params"]["jz"], model["params"]["hx"])
        return hb, hs, (lambda site: {"sx": X, "sz": Z})
    raise ValueError(model["name"])


def rel_err(a, b):
    return float((a - b).norm() / b.norm())


class rotated_qr:
    """Context manager: torch.linalg.qr returns Q @ G (G a fixed random orthogonal matrix) instead of Q.
    Any orthonormal basis of range(Y) is mathematically equivalent inside the randomized SVD, so the oracle run
    under this patch measures how far the *reference algorithm itself* moves under an equivalent re-implementation
    of its QR step -- the reproducibility envelope that bounds what parity can be asked of a different QR kernel."""

    def __init__(self, seed=1):
        self.seed = seed

    def __enter__(self):
        import collections
        self._real = torch.linalg.qr
        QR = collections.namedtuple("QR", ["Q", "R"])
        real, seed = self._real, self.seed

        def qr(Y, mode="reduced"):
            Q, R = real(Y, mode=mode)
            g = torch.Generator().manual_seed(seed * 1000003 + Q.shape[0] * 7 + Q.shape[1])
            G = real(torch.randn(Q.shape[1], Q.shape[1], dtype=Q.dtype, generator=g)).Q.to(Q.device)
            return QR(Q @ G, G.mH @ R)
        torch.linalg.qr = qr
        return self

    def __exit__(self, *exc):
        torch.linalg.qr = self._real
