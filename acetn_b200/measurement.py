"""B200 counterparts of acetn/measurement: RDM (site / bond reduced density matrices) and measure().

The contraction sequences of acetn/measurement/rdm.py:35-154 run behind two C-ABI entry points, acetn_b200_site_rdm and
acetn_b200_bond_rdm (acetn_b200/csrc/environment.cu: the bond RDM is blocked over the bra physical index like
build_bond_rdm_core_blocked, but each half is computed d times instead of d^2 times; every leg permutation is folded into the
index descriptors of the K1 GEMMs).  The d x d (d^2 x d^2) traces against the operators are host-side scalars, as in
acetn/measurement/measure.py:114-198."""
import torch

from . import ops


class RDM:
    """rdm.py:6-31 : rdm[site] -> (d,d) [bra,ket]; rdm[bond] -> (d,d,d,d) [P,Q,p,q]; bond = (s1, s2, k)."""

    def __init__(self, ipeps):
        self.ipeps = ipeps

    def __getitem__(self, key):
        if hasattr(key, "k") or (isinstance(key, (tuple, list)) and len(key) == 3 and not isinstance(key[0], int)):
            return self.build_bond_rdm(key)      # reference Bond dataclass (bond.py:5-22) or a plain (s1, s2, k)
        return self.build_site_rdm(key)

    def build_site_rdm(self, site):
        """rdm.py:35-67."""
        st = self.ipeps[site]
        return ops.site_rdm(st['C'], st['E'], st['A'])

    def build_bond_rdm(self, bond):
        """rdm.py:69-154."""
        s1, s2, k = bond
        return ops.bond_rdm(self.ipeps[s1], self.ipeps[s2], k)


def measure(ipeps, bond_ham, site_ham=None, site_ops=None):
    """measure.py:5-28 : {'Energy': per-site energy, <name>: site-averaged one-site observables}.
    bond_ham: (d^2,d^2) two-site Hamiltonian (or callable bond -> matrix); site_ham: (d,d), callable or None;
    site_ops: callable site -> {name: (d,d)} or None."""
    d = ipeps.dims['phys']
    rdm = RDM(ipeps)
    dev = ipeps[ipeps.site_list[0]]['A'].device
    out = {'Energy': torch.zeros((), dtype=torch.float64, device=dev)}
    names = list(site_ops(ipeps.site_list[0]).keys()) if site_ops else []
    for nme in names:
        out[nme] = torch.zeros((), dtype=torch.float64, device=dev)
    for site in ipeps.site_list:
        rho = rdm[site]
        nrm = torch.einsum("pp->", rho)
        hs = site_ham(site) if callable(site_ham) else site_ham
        if hs is not None:
            out['Energy'] = out['Energy'] + torch.einsum("Pp,pP->", rho, hs.to(dev)) / nrm
        if site_ops:
            for nme, op in site_ops(site).items():
                out[nme] = out[nme] + torch.einsum("Pp,pP->", rho, op.to(dev)) / nrm
    for nme in names:
        out[nme] = out[nme] / len(ipeps.site_list)
    for bond in ipeps.bond_list:
        rho = rdm[bond]
        nrm = torch.einsum("pqpq->", rho)
        hb = bond_ham(bond) if callable(bond_ham) else bond_ham
        out['Energy'] = out['Energy'] + torch.einsum("PQpq,pqPQ->", rho, hb.to(dev).reshape(d, d, d, d)) / nrm
    out['Energy'] = out['Energy'] / len(ipeps.site_list)
    return out
