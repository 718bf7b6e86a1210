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

    def projectors(self, cell, tasks, omegas):
        out = []
        for t, om in zip(tasks, omegas):
            out.append(orc.half_system_projectors(cell, t["plaq"], t["k"], self.cfg, omega_fn=lambda n, q, dt=None, dv=None, om=om: om))
        return [(a.contiguous(), b.contiguous()) for a, b in out]

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


def _worker(rank, world, port, nx, ny, kind, ret):
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
        sh = ShardedCtmrg(cell, cfg, rank, world, compute=OracleCompute(cfg, orc.OmegaTape(tape.tape)))
        sh.run()
        ret[rank] = (_max_diff(ref, cell), sh.bytes_exchanged)
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
