"""Shared helpers of the drop-in tests (tests/test_dropin_cpu.py, tests/test_gpu_dropin.py): the UNMODIFIED reference package
from oracle/_ref (or /root/reference inside the build container) + acetn_b200.integration.install()."""
import csv
import os

from oracle import vendor_ref

CASES = ["ising_dims_2_20_dtau_001_hx_295", "heisenberg_dims_3_16_dtau_001"]


def ref_dir():
    return os.path.join(vendor_ref.import_path(), "tests", "integration")


def setup():
    vendor_ref.enable()
    import toml  # noqa: F401
    from acetn.ipeps import Ipeps
    import acetn_b200.integration as b200
    b200.install()
    return Ipeps


def config(case, **over):
    import toml
    cfg = toml.load(os.path.join(ref_dir(), "input", case + ".toml"))
    cfg.setdefault("ctmrg", {})["disable_progressbar"] = True
    cfg.setdefault("evolution", {})["disable_progressbar"] = True
    for k, v in over.items():
        cfg[k].update(v) if isinstance(v, dict) else cfg.__setitem__(k, v)
    return cfg


def energies():
    with open(os.path.join(ref_dir(), "ipeps_gs", "energies.csv"), newline="") as f:
        return {r[0]: float(r[1]) for r in csv.reader(f)}


class MoveCounter:
    """Counts calls into the REFERENCE's DirectionalMover moves (acetn/renormalization/directional_mover.py:23-97,183-271)."""
    NAMES = ["left_move", "right_move", "up_move", "down_move", "left_right_move_dist", "up_down_move_dist"]

    def __enter__(self):
        from acetn.renormalization.directional_mover import DirectionalMover as RefMover
        self.cls, self.saved, self.calls = RefMover, {}, 0
        for n in self.NAMES:
            fn = getattr(RefMover, n)
            self.saved[n] = fn

            def wrap(this, *a, _fn=fn, **kw):
                self.calls += 1
                return _fn(this, *a, **kw)
            setattr(RefMover, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, fn in self.saved.items():
            setattr(self.cls, n, fn)
