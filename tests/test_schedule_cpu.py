"""CPU tests of the host-side scheduling logic: the paired-phase order and the site-sharded schedule
(acetn_b200/distributed.py) driven with the oracle's compute functions on gloo, world_size 2 -- they must reproduce
the reference's sequential sweep (ctmrg.py:25-31) exactly, independent of the number of ranks."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ctmrg_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleCompute:
    """acetn_b200.distributed compute backend made of oracle functions (test-only)."""

    def __init__(self, cfg, tape):
        self.cfg, self.tape = cfg, tape

    def tasks(self, cell, k, line):
        return [dict(k=k, line=line, key=key, plaq=plaq, s1=s1, s2=s2, i=i, j=j) for key, plaq, s1, s2, i, j in orc.move_tasks(cell, k, line)]

    def draw_omega(self, cell, task):
        D, chi, k = cell.dims["bond"], cell.dims["chi"], task["k"]
        n = cell[task["plaq"][3]].E[(3 + k + 3) % 4].shape[0] * D * D
        m = cell[task["plaq"][0]].E[k % 4].shape[1] * D * D
        return self.tape(n, min(chi + self.cfg.rsvd_oversampling, m, n))

    def projectors(self, cell, tasks, omegas, group=None, group_rank=0, group_size=1):
        out = []
        if group_size > 1:
            from acetn_b200.sharded_projector import ShardedHalfSystemProjector
            sp = ShardedHalfSystemProjector(self.cfg, group, group_rank, group_size, la=TorchLinAlg())
            pend = [sp.begin(cell, t["plaq"], t["k"], om) for t, om in zip(tasks, omegas)]
            out = [sp.finish(pd) for pd in pend]
        else:
            for t, om in zip(tasks, omegas):
                out.append(orc.half_system_projectors(cell, t["plaq"], t["k"], self.cfg, omega_fn=lambda n, q, dt=None, dv=None, om=om: om))
        return [(a.contiguous(), b.contiguous()) for a, b in out]

    def absorb_coop(self, cell, task, p1i, p2i, p1j, p2j, group, g, G):
        """Row-sharded edge absorption (un-normalised partials over the shared leg a, summed, then normalised)."""
        from acetn_b200.distributed import _a_block
        k, a = task["k"], cell[task["s1"]]
        c1 = orc.absorb_corner1(a.C[(3 + k) % 4], a.E[(2 + k) % 4], p1i)
        c2 = orc.absorb_corner2(a.C[k], a.E[k], p2j)
        ei, ai = a.E[(3 + k) % 4], a.bond_permute(k)
        a0, a1 = _a_block(ei.shape[0], G, g)
        t = torch.einsum("ablL,buUx->alLuUx", ei[a0:a1], p1j)
        t = torch.einsum("LURDP,alLuUx->RDPalux", ai.conj(), t)
        t = torch.einsum("lurdp,RDpalux->rdRDax", ai, t)
        e = torch.einsum("rdRDax,adDy->yxrR", t, p2i[a0:a1]).contiguous()
        dist.all_reduce(e, op=dist.ReduceOp.SUM, group=group)
        return c1, c2, e / e.norm()

    def absorb(self, cell, task, p1i, p2i, p1j, p2j):
        k, a = task["k"], cell[task["s1"]]
        return (orc.absorb_corner1(a.C[(3 + k) % 4], a.E[(2 + k) % 4], p1i), orc.absorb_corner2(a.C[k], a.E[k], p2j),
                orc.absorb_edge(a.E[(3 + k) % 4], a.bond_permute(k), p2i, p1j))


def _max_diff(a, b):
    err = 0.0
    for s in a.site_list:
        for k in range(4):
            assert a[s].C[k].shape == b[s].C[k].shape and a[s].E[k].shape == b[s].E[k].shape
            err = max(err, float((a[s].C[k] - b[s].C[k]).abs().max()), float((a[s].E[k] - b[s].E[k]).abs().max()))
    return err


def _worker(rank, world, port, nx, ny, kind, ret, group_size=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from acetn_b200.distributed import ShardedCtmrg
        cfg = orc.CtmrgConfig(steps=2)
        cell = orc.random_cell(nx, ny, 2, 6, 2, seed=3) if kind == "random" else orc.product_cell(nx, ny, 2, 8, 2, seed=3)
        ref = cell.clone()
        tape = orc.OmegaTape()
        orc.ctmrg(ref, cfg, omega_fn=tape)
        sh = ShardedCtmrg(cell, cfg, rank, world, compute=OracleCompute(cfg, orc.OmegaTape(tape.tape)), group_size=group_size)
        sh.run()
        H = orc.heisenberg_bond_hamiltonian(1.0)
        de = abs(float(orc.measure(ref, H)["Energy"]) - float(orc.measure(cell, H)["Energy"]))
        ret[rank] = (_max_diff(ref, cell) if group_size == 1 else de, sh.bytes_exchanged)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny,kind", [(2, 2, "random"), (3, 2, "random"), (2, 2, "product")])
def test_sharded_schedule_gloo_world2(nx, ny, kind):
    world = 2
    port = 29500 + (os.getpid() % 2000) + 7 * nx + ny
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, nx, ny, kind, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, nbytes = ret[rank]
        assert err == 0.0, f"rank {rank}: sharded sweep differs from the sequential oracle sweep by {err}"
        assert nbytes > 0


def test_group_cooperative_schedule_gloo_world2():
    """world 2, group_size 2: both ranks compute every projector cooperatively (row-sharded), absorptions alternate.
    Product-state start (well conditioned): the energy after two sweeps equals the sequential oracle's to 1e-9."""
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 2, 2, "product", ret, 2), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        de, nbytes = ret[rank]
        assert de < 1e-9, de
        assert nbytes > 0
    assert ret[0][0] == ret[1][0]


def test_sharded_schedule_single_rank_equals_sequential():
    from acetn_b200.distributed import ShardedCtmrg
    cfg = orc.CtmrgConfig(steps=2)
    cell = orc.random_cell(2, 2, 2, 6, 2, seed=5)
    ref = cell.clone()
    tape = orc.OmegaTape()
    orc.ctmrg(ref, cfg, omega_fn=tape)
    sh = ShardedCtmrg(cell, cfg, 0, 1, compute=OracleCompute(cfg, orc.OmegaTape(tape.tape)))
    sh.run()
    assert _max_diff(ref, cell) == 0.0


def test_full_system_sharding_rejected():
    from acetn_b200.distributed import ShardedCtmrg
    with pytest.raises(ValueError):
        ShardedCtmrg(orc.random_cell(2, 2, 2, 4, 2), orc.CtmrgConfig(projectors="full-system"), 0, 1, compute=object())


def test_move_tasks_match_oracle_pickers():
    """Host mirror of the plaquette pickers / renormalize_boundary indices (directional_mover.py:23-181)."""
    from acetn_b200.renormalization import DirectionalMover
    for nx, ny in [(2, 2), (3, 2), (4, 4), (1, 3)]:
        cell = orc.random_cell(nx, ny, 2, 2, 2, seed=0)
        for k in range(4):
            for line in range(nx if k in (0, 2) else ny):
                mine = DirectionalMover.move_tasks(cell, k, line)
                ref = orc.move_tasks(cell, k, line)
                assert [(t["key"], t["plaq"], t["s1"], t["s2"], t["i"], t["j"]) for t in mine] == [tuple(r) for r in ref]


# ---------------------------------------------------------------------------------------------------------------------
# row-sharded projector (acetn_b200/sharded_projector.py) on gloo with a torch linear-algebra backend
# ---------------------------------------------------------------------------------------------------------------------
class TorchLinAlg:
    """Test-only backend of ShardedHalfSystemProjector: the exchange logic is what is under test here."""

    def quarter_rows(self, site, k, c0, c1, absmax):
        sl = orc.Site(site.A, list(site.C), list(site.E))
        sl.E[k % 4] = site.E[k % 4][:, c0:c1]
        ak = sl.bond_permute(k)
        t = torch.einsum("ab,bcuU->acuU", sl.C[k % 4], sl.E[k % 4])
        t = torch.einsum("acuU,ealL->cuUelL", t, sl.E[(3 + k) % 4])
        t = torch.einsum("cuUelL,LURDP->cuelRDP", t, ak.conj())
        t = torch.einsum("lurdp,cuelRDp->crRedD", ak, t)
        shp = t.shape
        Q = t.reshape(shp[0] * shp[1] * shp[2], shp[3] * shp[4] * shp[5])
        if Q.numel():
            absmax[0] = max(float(absmax[0]), float(Q.abs().max()))
        return Q

    def matmul(self, A, B, transpose_a=False, out=None):
        r = (A.T if transpose_a else A) @ B
        if out is not None:
            out.copy_(r)
            return out
        return r

    def orthonormalize(self, Y):
        Y.copy_(torch.linalg.qr(Y).Q)
        return Y

    def core_svd(self, R, chi, cutoff):
        U, S, Vh = torch.linalg.svd(R)          # R = U S Vh  =>  Jt = U^T, Wt = Vh
        keep = min(chi, int((S / S[0] > cutoff).sum()))
        return S, Vh.contiguous(), U.T.contiguous(), torch.tensor([keep, 0])

    def keep_of(self, info):
        return int(info[0])


def _proj_worker(rank, world, port, kind, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from acetn_b200.sharded_projector import ShardedHalfSystemProjector
        cfg = orc.CtmrgConfig()
        if kind == "random":
            cell = orc.random_cell(2, 2, 3, 7, 2, seed=4)       # chi = 7: uneven row blocks
        else:
            cell = orc.product_cell(2, 2, 2, 10, 2, seed=2)
            orc.ctmrg(cell, orc.CtmrgConfig(steps=1))             # ragged chi legs
        plaq = orc.plaquette(cell, 0, 0, 0)
        tape, rec = orc.OmegaTape(), {}
        p1r, p2r = orc.half_system_projectors(cell, plaq, 0, cfg, omega_fn=tape, record=rec)
        sp = ShardedHalfSystemProjector(cfg, None, rank, world, la=TorchLinAlg())
        sp.spectra = []
        p1, p2 = sp.finish(sp.begin(cell, plaq, 0, tape.tape[0]))
        chi = cell.dims["chi"]
        ds = float((rec["spectra"][0] - sp.spectra[0])[:chi].abs().max())
        m = p1r.shape[0] * p1r.shape[1] * p1r.shape[2]
        pi_ref = p2r.reshape(m, -1) @ p1r.reshape(m, -1).T
        pi = p2.reshape(m, -1) @ p1.reshape(m, -1).T
        ret[rank] = (tuple(p1.shape) == tuple(p1r.shape), ds, float((pi - pi_ref).norm() / pi_ref.norm()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["random", "product"])
def test_row_sharded_projector_gloo_world2(kind):
    world = 2
    port = 31500 + (os.getpid() % 2000) + (3 if kind == "random" else 5)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_proj_worker, args=(world, port, kind, ret), nprocs=world, join=True)
    for rank in range(world):
        same_shape, ds, dpi = ret[rank]
        assert same_shape
        assert ds < 1e-10, ds
        assert dpi < 1e-7, dpi
    assert ret[0][1:] == ret[1][1:]          # both ranks hold identical results


@pytest.mark.parametrize("inflight", [1, 2])
def test_inflight_window_does_not_change_results(inflight, monkeypatch):
    """The bounded in-flight window of the phase scheduler (DirectionalMover._pipeline: at most W tasks hold their quarter
    tensors; task n + W re-uses the slot of task n) must give the tensors of the unbounded schedule bit for bit -- host
    logic, checked with the CPU stand-ins of the C-ABI wrappers on a 3x2 cell (6 tasks per up/down phase)."""
    from acetn_b200.ipeps import CTMRGConfig, Ipeps
    from acetn_b200.renormalization import DirectionalMover, ctmrg
    from tests.cpu_emulation import emulated

    def run(w):
        monkeypatch.setenv("ACETN_B200_INFLIGHT", str(w))
        cell = orc.random_cell(3, 2, 2, 6, 2, seed=5)
        torch.manual_seed(21)
        with emulated():
            ip = Ipeps.from_plain(cell, CTMRGConfig(steps=2), device="cpu")
            mover = DirectionalMover(ip.ctmrg_config)
            assert mover.inflight == w
            ctmrg(ip, ip.ctmrg_config, mover)
            assert len(mover._slots) == min(w, 6)
        return ip

    a, b = run(16), run(inflight)
    for s in a.site_list:
        for k in range(4):
            assert torch.equal(a[s]['C'][k], b[s]['C'][k]) and torch.equal(a[s]['E'][k], b[s]['E'][k])
