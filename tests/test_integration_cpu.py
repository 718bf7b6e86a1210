"""CPU test of the drop-in wiring: acetn_b200.integration.install() on the actual reference package when it is present
in this container (/root/reference; skipped elsewhere -- the reference does not travel to the GPU box)."""
import os
import sys

import pytest
import torch

REF = os.environ.get("ACETN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "acetn")), reason="reference package not available")


def test_install_registers_backend_and_keeps_torch_path():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import acetn  # noqa: F401
    from acetn.ipeps import Ipeps
    import acetn_b200.integration as b200
    b200.install()
    b200.install()      # idempotent
    base = {"dtype": "float64", "device": "cpu", "TN": {"nx": 2, "ny": 2, "dims": {"phys": 2, "bond": 2, "chi": 4}},
            "model": {"name": "heisenberg", "params": {"J": 1.0}}, "ctmrg": {"steps": 1, "disable_progressbar": True}}
    # the untouched reference path still runs
    ip = Ipeps(dict(base))
    ip.renormalize()
    assert ip[(0, 0)]['C'][0].shape[0] <= 4
    # requesting the b200 backend without a CUDA device must raise, not fall back (north star)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            Ipeps(dict(base, evolution={"backend": "b200"}))
