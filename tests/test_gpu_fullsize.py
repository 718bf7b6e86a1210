"""GPU tests at BASELINE.json's full sizes (D=8 chi=256; D=6 chi=144) through size-independent properties of the
domain (the CPU oracle needs minutes per site-move there):
  * bi-orthogonality: proj1^T proj2 = c * I  (U^T (Q1 Q4) V = diag(S) exactly for the rSVD factors; projectors.py:166-173)
  * unit Frobenius norm and shapes of the absorbed C, C, E (directional_mover.py:323,343,366)
  * bit-reproducibility of a whole site-move (deterministic reductions, fixed-order split-K)
  * truncated spectrum is descending, s[0] = 1, and invariant under rescaling of the inputs."""
import pytest
import torch

from oracle import ctmrg_oracle as orc

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from acetn_b200 import linalg, ops
    from acetn_b200.ipeps import CTMRGConfig, Ipeps
    from acetn_b200.renormalization import DirectionalMover


def site_move(ip, omega_seed, record=None):
    torch.manual_seed(omega_seed)
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.spectra = [] if record is None else record
    p1, p2 = mover._projectors_of_tasks(ip, mover.move_tasks(ip, 0, 0))
    mover2 = DirectionalMover(ip.ctmrg_config)
    torch.manual_seed(omega_seed)
    mover2.left_move(ip, 0)
    return p1, p2


@pytest.mark.parametrize("D,chi", [(8, 256), (6, 144)])
def test_full_size_site_move_properties(D, chi):
    d = 2
    cell = orc.random_cell(2, 2, D, chi, d, seed=0)
    ip = Ipeps.from_plain(cell, CTMRGConfig())
    spectra = []
    p1, p2 = site_move(ip, 3, spectra)
    m = chi * D * D
    for key in p1:
        a, b = p1[key].reshape(m, -1), p2[key].reshape(m, -1)
        g = ops.matmul(a, b, transpose_a=True)                      # proj1^T proj2, (chi', chi')
        dg = torch.diagonal(g)
        off = g - torch.diag(dg)
        assert float(off.abs().max() / dg.abs().mean()) < 1e-9
        assert float((dg - dg.mean()).abs().max() / dg.abs().mean()) < 1e-9
    for s in spectra[:2]:
        assert float(s[0]) == 1.0
        assert torch.all(s[:-1] >= s[1:])
    # the left move wrote C[3], C[0], E[3] of column 1: unit norm, right shapes
    for y in range(2):
        st = ip[(1, y)]
        for t in (st['C'][3], st['C'][0], st['E'][3]):
            assert abs(float(t.norm()) - 1.0) < 1e-12
        assert tuple(st['C'][3].shape) == (chi, chi) and tuple(st['E'][3].shape) == (chi, chi, D, D)
    # bit-reproducibility: the same move from the same state and Omega gives identical bits
    ip2 = Ipeps.from_plain(cell, CTMRGConfig())
    site_move(ip2, 3)
    for y in range(2):
        for a, b in ((ip[(1, y)]['C'][3], ip2[(1, y)]['C'][3]), (ip[(1, y)]['C'][0], ip2[(1, y)]['C'][0]),
                     (ip[(1, y)]['E'][3], ip2[(1, y)]['E'][3])):
            assert torch.equal(a, b)


def test_spectrum_invariant_under_input_scaling():
    """s/s[0] and the normalised absorbed tensors do not depend on the overall scale of C and E (every output is
    normalised; the quarter tensors are not divided by their max but the factor is carried, projectors.py:59)."""
    D, chi = 8, 64
    cell = orc.random_cell(2, 2, D, chi, 2, seed=1)
    scaled = cell.clone()
    for s in scaled.site_list:
        scaled[s].C = [c * 37.5 for c in scaled[s].C]
        scaled[s].E = [e * 0.013 for e in scaled[s].E]
    out = []
    for c in (cell, scaled):
        ip = Ipeps.from_plain(c, CTMRGConfig())
        rec = []
        site_move(ip, 5, rec)
        out.append((rec, ip))
    for a, b in zip(out[0][0], out[1][0]):
        assert float((a - b)[:chi].abs().max()) < 1e-10
    for y in range(2):
        a, b = out[0][1][(1, y)], out[1][1][(1, y)]
        sa, sb = torch.linalg.svdvals(a['C'][0].cpu()), torch.linalg.svdvals(b['C'][0].cpu())
        assert float((sa - sb).abs().max() / sa[0]) < 1e-9
