"""B200 counterparts of acetn/measurement: RDM (site / bond reduced density matrices) and measure().

The contraction sequences are those of acetn/measurement/rdm.py:35-154 (the bond RDM is blocked over the bra physical
index like build_bond_rdm_core_blocked, but each half is computed d times instead of d^2 times); every pairwise
contraction runs as gather + batched K1 DGEMM in libacetn_b200.so (ops.contract).  The d x d (d^2 x d^2) traces against
the operators are host-side scalars, as in acetn/measurement/measure.py:114-198."""
import torch

from .ops import contract


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
        c1, c2, c3, c4 = st['C']
        e1, e2, e3, e4 = st['E']
        a1 = st['A']
        t1 = contract("ab,bclL->aclL", c4, e4)
        t1 = contract("aclL,eadD->clLedD", t1, e3)
        t1 = contract("clLedD,LURDP->cledURP", t1, a1.conj())
        t2 = contract("ab,bcuU->acuU", c1, e1)
        t3 = contract("ab,carR->bcrR", c3, e2)
        t3 = contract("ec,bcrR->ebrR", c2, t3)
        t3 = contract("ebrR,aeuU->brRauU", t3, t2)
        t3 = contract("erRcuU,cledURP->ruldP", t3, t1)
        return contract("ruldP,lurdp->Pp", t3, a1)

    def build_bond_rdm(self, bond):
        """rdm.py:69-154."""
        s1, s2, k = bond
        a, b = self.ipeps[s1], self.ipeps[s2]
        c12, e12, e11 = a['C'][(k + 1) % 4], a['E'][(k + 1) % 4], a['E'][k % 4]
        c13, e13 = a['C'][(k + 2) % 4], a['E'][(k + 2) % 4]
        a1 = a.bond_permute(k)
        c21, e21, e24 = b['C'][k % 4], b['E'][k % 4], b['E'][(k + 3) % 4]
        c24, e23 = b['C'][(k + 3) % 4], b['E'][(k + 2) % 4]
        a2 = b.bond_permute(k)
        d = a1.shape[-1]

        # both halves through the gather-free environment contraction of the norm tensor (evolution.env_front / env_back): the
        # bra site factor with its physical index fixed, the ket one with it open
        #   right[P][f,c,L,(l,p)] = tmp_r2[a,f,d,D] (tr1 conj(a1)[L,U,R,D,P] a1[l,u,r,d,p])[a,c,L,D,l,d,p]      (rdm.py:95-99, 133-139)
        #   left[Q][f,c,R,(r,q)]  = tmp_l2[e,f,d,D] (tl1 conj(a2)[L,U,R,D,Q] a2[l,u,r,d,q])[c,e,R,D,r,d,q]      (rdm.py:100-104, 141-147)
        from .evolution import closing_left, closing_right, env_back, env_front
        D = a1.shape[0]
        D2 = D * D
        ket1 = a1.permute(2, 1, 3, 0, 4).contiguous().reshape(D2, D * D * d)             # [(r,u)][(d,(l,p))]
        ket2 = a2.permute(1, 0, 3, 2, 4).contiguous().reshape(D2, D * D * d)             # [(u,l)][(d,(r,q))]
        tr, dims_r = env_front(c12, e12, e11)
        A5r = closing_right(c13, e13)
        right = []
        for P in range(d):
            bra = a1[..., P].permute(2, 1, 3, 0).contiguous().reshape(D2, D2)           # [(R,U)][(D,L)]
            right.append(env_back(tr, dims_r, A5r, bra, ket1, D, D * d, left=False).reshape(A5r.shape[0], -1, D, D, d))
        del tr
        tl, dims_l = env_front(c21, e21, e24)
        A5l = closing_left(c24, e23)
        left = []
        for Q in range(d):
            bra = a2[..., Q].permute(1, 0, 2, 3).contiguous().reshape(D2, D2)           # [(U,L)][(R,D)]
            left.append(env_back(tl, dims_l, A5l, bra, ket2, D, D * d, left=True).reshape(A5l.shape[0], -1, D, D, d))
        del tl
        rho = torch.empty(d, d, d, d, dtype=a1.dtype, device=a1.device)
        for P in range(d):
            for Q in range(d):
                rho[P, Q] = contract("fcRrp,fcRrq->pq", right[P], left[Q])
        return rho


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
