"""Drop-in installation behind the reference's own entry points.

    import acetn, acetn_b200.integration as b200
    b200.install()                                   # registers evolution.backend = "b200"
    ipeps = acetn.ipeps.Ipeps({..., "device": "cuda", "evolution": {"backend": "b200"}})
    ipeps.renormalize()                              # CTMRG now runs in libacetn_b200.so

`install()` only touches the seams the reference itself exposes by name:
  * `acetn.ipeps.ipeps_config.EvolutionConfig.backend` accepts "b200"; `IpepsConfig.validate_backend` RAISES (no fallback)
    when "b200" is requested without the library or a CUDA device (the reference falls back to "torch" for "cutensor",
    ipeps_config.py:103-109 -- the north star forbids that here);
  * `acetn.ipeps.ipeps.ctmrg` (the name `Ipeps.renormalize` calls, ipeps.py:93-97) is wrapped: backend "b200" routes to
    acetn_b200.renormalization.ctmrg, anything else to the untouched reference function (which stays the oracle path);
  * `acetn.evolution.full_update.FullUpdater.tensor_update` (one bond update: QR split, norm tensor, positive_approx,
    gauge_fix, ALS, finalisation) routes to `acetn_b200.evolution.full_update_bond` for backend "b200";
  * `acetn.evolution.fast_full_update.FastFullUpdater` (what `ipeps.evolve` builds for update_type="full", evolve.py:11-15): its
    `mover` (fast_full_update.py:27) becomes `acetn_b200.renormalization.DirectionalMover`, and `absorb_bond` /
    `absorb_bond_dist` (:72-129) -- the two CTMRG moves after EVERY bond update, which dominate `evolve` --
    route to `DirectionalMover.absorb_bond` / the site-sharded phases of acetn_b200.distributed;
  * `acetn.renormalization.projectors.{svd_lowrank,fused_matmul_svd_lowrank,fused_3matmul_svd_lowrank}` are NOT replaced
    globally: backend "torch" keeps the reference numerics bit for bit.
The reference's SiteTensor / TensorNetwork objects are used as they are: the B200 mover only needs `ipeps[site]['A'|'C'|'E']`,
`bond_permute`, `ipeps.nx/ny/dims` and writes list items in place exactly like directional_mover.py:295-303.
"""
import functools
from types import SimpleNamespace

from . import _lib


def install(acetn_module=None):
    """Register backend='b200' in an importable reference package.  Returns the patched module."""
    if acetn_module is None:
        import acetn as acetn_module  # noqa: F401  (the reference must be importable)
    import acetn.ipeps.ipeps as ipeps_mod
    import acetn.ipeps.ipeps_config as cfg_mod
    from . import renormalization as b200_renorm

    if getattr(ipeps_mod, "_acetn_b200_installed", False):
        return acetn_module

    # ---- config: accept and validate the new literal -------------------------------------------------------------
    orig_validate = cfg_mod.IpepsConfig.validate_backend

    def validate_backend(self):
        if self.evolution.backend == "b200":
            import torch
            if not _lib.available():
                raise RuntimeError("backend='b200' requested but libacetn_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
            if not torch.cuda.is_available() or torch.device(self.device).type != "cuda":
                raise RuntimeError("backend='b200' requires device='cuda' on a B200; there is no CPU fallback")
            if self.dtype != torch.float64:
                raise RuntimeError("backend='b200' supports dtype float64 only")
            return
        return orig_validate(self)

    cfg_mod.IpepsConfig.validate_backend = validate_backend

    # ---- CTMRG: route Ipeps.renormalize() ----------------------------------------------------------------------------
    ref_ctmrg = ipeps_mod.ctmrg

    @functools.wraps(ref_ctmrg)
    def ctmrg(ipeps, config):
        if getattr(ipeps.config.evolution, "backend", "torch") == "b200":
            if getattr(ipeps, "is_distributed", False):
                from .distributed import ShardedCtmrg
                return ShardedCtmrg(ipeps, config, ipeps.rank, ipeps.world_size).run()
            return b200_renorm.ctmrg(ipeps, config)
        return ref_ctmrg(ipeps, config)

    ipeps_mod.ctmrg = ctmrg

    # ---- measure: RDM contractions (measure.py:5-28 builds RDM(ipeps) and indexes it) ---------------------------------
    import acetn.measurement.measure as meas_mod
    from .measurement import RDM as B200RDM
    ref_rdm = meas_mod.RDM

    def rdm_factory(ipeps):
        if getattr(ipeps.config.evolution, "backend", "torch") == "b200":
            return B200RDM(ipeps)
        return ref_rdm(ipeps)

    meas_mod.RDM = rdm_factory

    # ---- full update: norm tensor + ALS inner loop (full_update.py:54, als_solver.py:48-51) ------------------------------
    import acetn.evolution.als_solver as als_mod
    import acetn.evolution.full_update as fu_mod
    from . import evolution as b200_evo
    ref_norm = fu_mod.build_norm_tensor

    def build_norm_tensor(ipeps, bond, a1q, a2q):
        if getattr(ipeps.config.evolution, "backend", "torch") == "b200":
            return b200_evo.build_norm_tensor(ipeps, bond, a1q, a2q)
        return ref_norm(ipeps, bond, a1q, a2q)

    fu_mod.build_norm_tensor = build_norm_tensor
    ref_solve = als_mod.ALSSolver.solve

    def solve(self):
        if self.backend == "b200":
            cfg = SimpleNamespace(als_niter=self.niter, als_tol=self.tol, als_method=self.method, als_epsilon=self.epsilon)
            return b200_evo.ALSSolver(self.n12, self.a12g, tuple(self.ar_shape), cfg).solve()
        return ref_solve(self)

    als_mod.ALSSolver.solve = solve

    # ---- the callers around them: one whole bond update on the library's kernels ---------------------------
    ref_tensor_update = fu_mod.FullUpdater.tensor_update

    def tensor_update(self, a1, a2, bond):
        if self.backend == "b200":
            return b200_evo.full_update_bond(self.ipeps, bond, a1.contiguous(), a2.contiguous(), self.gate[bond], self.config)
        return ref_tensor_update(self, a1, a2, bond)

    fu_mod.FullUpdater.tensor_update = tensor_update

    # ---- evolve: the CTMRG moves that absorb every updated bond (fast_full_update.py:27, 72-129) ----------------------------
    import acetn.evolution.fast_full_update as ffu_mod
    FFU = ffu_mod.FastFullUpdater
    ref_ffu_init, ref_absorb, ref_absorb_dist = FFU.__init__, FFU.absorb_bond, FFU.absorb_bond_dist

    def ffu_init(self, ipeps, gate, config):
        ref_ffu_init(self, ipeps, gate, config)
        if getattr(config, "backend", "torch") == "b200":
            self.mover = b200_renorm.DirectionalMover(ipeps.config.ctmrg)

    @ffu_mod.record_runtime
    def b200_absorb(self, bond):
        self.mover.absorb_bond(self.ipeps, bond)

    def absorb_bond(self, bond):
        if self.backend == "b200":
            return b200_absorb(self, bond)
        return ref_absorb(self, bond)

    def absorb_bond_dist(self, bond):
        # the reference's body (mover.left_right_move_dist / up_down_move_dist, :117-129) runs unchanged: for backend "b200" the
        # mover is the B200 one, whose *_dist moves are the site-sharded phases of acetn_b200.distributed.ShardedCtmrg
        return ref_absorb_dist(self, bond)

    FFU.__init__ = ffu_init
    FFU.absorb_bond = absorb_bond
    FFU.absorb_bond_dist = absorb_bond_dist
    ipeps_mod._acetn_b200_installed = True
    return acetn_module
