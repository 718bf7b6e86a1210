"""Minimal stand-alone substrate with the reference's accessors (acetn/ipeps/{site_tensor,tensor_network,ipeps,
ipeps_config}.py) so the B200 path can run where the reference package is not installed (the GPU box).  When the
reference *is* installed, use acetn_b200.integration.install() and the reference's own Ipeps instead."""
from dataclasses import dataclass
from typing import Dict

import torch


@dataclass
class CTMRGConfig:          # ipeps_config.py:16-24
    steps: int = 40
    projectors: str = "half-system"
    svd_type: str = "rsvd"
    svd_cutoff: float = 1e-12
    rsvd_niter: int = 2
    rsvd_oversampling: int = 2
    disable_progressbar: bool = True
    thin_engine: str = "auto"   # not a reference field: "auto" | "dmma" | "i8" (renormalization.ProjectorCalculator)


class SiteTensor:
    """site_tensor.py:41-150 accessors: st['A'], st['C'][k], st['E'][k], st.bond_permute(k)."""

    def __init__(self, A, C, E):
        self._A = A
        self._C = list(C)
        self._E = list(E)

    def __getitem__(self, key):
        if key == 'A':
            return self._A
        if key == 'C':
            return self._C
        if key == 'E':
            return self._E
        raise ValueError(f"Invalid key: '{key}' provided.")

    def __setitem__(self, key, val):
        if key == 'A':
            self._A = val.detach().clone()
        elif key == 'C':
            self._C = [v.detach().clone() for v in val]
        elif key == 'E':
            self._E = [v.detach().clone() for v in val]
        else:
            raise ValueError(f"Invalid key: '{key}' provided.")

    def bond_permute(self, k):
        return self._A.permute([(i + k) % 4 for i in range(4)] + [4])

    def to(self, device):
        return SiteTensor(self._A.to(device), [c.to(device) for c in self._C], [e.to(device) for e in self._E])


class Ipeps:
    """nx x ny unit cell (tensor_network.py:13-56,153-193) + ctmrg config + renormalize/measure entry points
    (ipeps.py:93-128) routed to the b200 backend."""

    def __init__(self, nx, ny, dims, sites: Dict, ctmrg: CTMRGConfig = None, device="cuda"):
        self.nx, self.ny = nx, ny
        self.dims = dict(dims)
        self.device = torch.device(device)
        self._sites = {tuple(k): v.to(self.device) for k, v in sites.items()}
        self.site_list = [(x, y) for x in range(nx) for y in range(ny)]
        self.bond_list = [((x, y), ((x + 1) % nx, y), 2) for x in range(nx) for y in range(ny)] + \
                         [((x, y), (x, (y + 1) % ny), 1) for y in range(ny) for x in range(nx)]
        self.ctmrg_config = ctmrg or CTMRGConfig()
        self.rank, self.world_size, self.is_distributed = 0, 1, False

    def __getitem__(self, site):
        st = self._sites.get(tuple(site))
        if st is None:
            raise ValueError(f"Site tensor not defined at site {site}.")
        return st

    def __setitem__(self, site, st):
        self._sites[tuple(site)] = st

    def set_chi(self, chi):
        self.dims['chi'] = chi

    def renormalize(self, mover=None):
        from .renormalization import ctmrg
        return ctmrg(self, self.ctmrg_config, mover)

    def measure(self, bond_ham, site_ham=None, site_ops=None):
        from .measurement import measure
        return measure(self, bond_ham, site_ham, site_ops)

    @staticmethod
    def from_plain(cell, ctmrg: CTMRGConfig = None, device="cuda"):
        """Build from any object exposing nx, ny, dims, site_list and cell[site]['A'/'C'/'E'] (e.g. the oracle Cell)."""
        sites = {s: SiteTensor(cell[s]['A'].clone(), [c.clone() for c in cell[s]['C']], [e.clone() for e in cell[s]['E']])
                 for s in cell.site_list}
        return Ipeps(cell.nx, cell.ny, cell.dims, sites, ctmrg, device)
