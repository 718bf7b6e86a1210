"""B200 counterparts of acetn/renormalization: ProjectorCalculator, DirectionalMover, ctmrg.

Class/method names, argument meaning, tensor layouts, mutation semantics and error behaviour follow the reference
(acetn/renormalization/{projectors,directional_mover,ctmrg}.py) so this module drops in behind
`Ipeps.renormalize()` (see acetn_b200.integration).  All contractions, the randomized SVD and the normalisations
run in libacetn_b200.so; torch only allocates and draws Omega."""
import os

import torch

from . import linalg, ops


class ProjectorCalculator:
    """acetn/renormalization/projectors.py:6-236.

    `calculate(ipeps, sites, k)` has the reference's signature.  Internally a projector is computed in two phases so
    that the independent site tasks of one directional move can be put on different CUDA streams:
    `begin(...)` enqueues quarter tensors + randomized SVD (no host sync), `finish(pending)` reads the truncated
    rank chi' (the reference's one host sync per projector, projectors.py:164) and enqueues the projector GEMMs."""

    def __init__(self, config):
        self.projectors = config.projectors
        self.svd_type = config.svd_type
        self.svd_cutoff = config.svd_cutoff
        self.rsvd_niter = config.rsvd_niter
        self.rsvd_oversampling = config.rsvd_oversampling
        self.spectra = None          # optional recorder: list receiving the normalised spectrum of every projector
        # engine of the rSVD / projector "big x thin" products: "dmma" = FP64 DMMA GEMM (K1); "i8" = exact integer products on
        # the INT8 tensor cores (K7, tcgen05); "auto" = K7 when the quarter tensors are large enough to pay for the encoding
        self.thin_engine = os.environ.get("ACETN_B200_THIN_ENGINE", getattr(config, "thin_engine", "auto"))
        if self.thin_engine not in ("auto", "dmma", "i8"):
            raise ValueError(f"Invalid thin_engine: {self.thin_engine} provided.")
        self.set_calculate()

    def set_calculate(self):
        if self.projectors is None or self.projectors == "full-system":
            self.calculate = self.calculate_full_system
            self.begin = self.begin_full_system
        elif self.projectors == "half-system":
            self.calculate = self.calculate_half_system
            self.begin = self.begin_half_system
        else:
            raise ValueError(f"Invalid ctmrg projector type: {self.projectors} provided.")

    @staticmethod
    def make_quarter_tensor(site_tensor, k, normalize=True, stream=None, absmax=None, out=None, enc_storage=None):
        """projectors.py:36-60 -> (Q matrix (chi D^2, chi D^2), 6-tuple shape).
        normalize=False skips the max-abs division (projectors.py:59): s/s[0], U and V are invariant under a rescaling
        of Q1/Q4, and the internal callers re-apply the factor 1/max|Q| to the small projector instead (absmax), which
        saves two passes over the 2 GiB tensor and reproduces the reference's projectors exactly."""
        ak = site_tensor.bond_permute(k)
        ck = site_tensor['C'][(0 + k) % 4]
        ek1 = site_tensor['E'][(3 + k) % 4]
        ek2 = site_tensor['E'][(0 + k) % 4]
        return ops.quarter_tensor(ck, ek2, ek1, ak, normalize=normalize, stream=stream, absmax=absmax, out=out, enc_storage=enc_storage)

    @staticmethod
    def quarter_shape(site_tensor, k):
        """(rows, cols) of make_quarter_tensor(site_tensor, k)."""
        D = site_tensor['A'].shape[0]
        return site_tensor['E'][(0 + k) % 4].shape[1] * D * D, site_tensor['E'][(3 + k) % 4].shape[0] * D * D

    @staticmethod
    def quarter_numel(site_tensor, k):
        """Number of elements of make_quarter_tensor(site_tensor, k): (chi_c D^2) x (chi_e D^2)."""
        D = site_tensor['A'].shape[0]
        return site_tensor['E'][(0 + k) % 4].shape[1] * site_tensor['E'][(3 + k) % 4].shape[0] * D ** 4

    I8_MIN_DIM = 4096        # "auto": quarter tensors at least this large go through K7

    def _use_i8(self, mats_shapes, q):
        """True when every factor's thin products should run on the INT8 tensor cores (K7)."""
        if self.thin_engine == "dmma":
            return False
        ok = all(ops.i8_supported(r, c, q) for r, c in mats_shapes)
        if self.thin_engine == "i8":
            if not ok:
                raise RuntimeError(f"thin_engine='i8': shapes {mats_shapes} with q={q} are outside the INT8 engine's range")
            return True
        return ok and all(min(r, c) >= self.I8_MIN_DIM for r, c in mats_shapes)

    def _check_svd_type(self):
        if self.svd_type not in ("rsvd", "full-rank"):
            raise ValueError(f"Invalid svd_type: {self.svd_type} provided.")

    # ---- svd_type == "full-rank" (projectors.py:114-136, 138-142, 176-179) ------------------------------------------
    FULL_RANK_MAX = 3400      # size limit of the block-Jacobi core (shared-memory staging of row-block pairs)

    @staticmethod
    def full_svd(M):
        """torch.linalg.svd(M) (projectors.py:229-231) from the same kernels as the randomized path: M = Qb Rc (K4 + one
        K1 product), Rc = Jt^T diag(S) Wt (K5)  =>  U = Qb Jt^T, V = Wt^T.  Returns U (m,n), S (n) descending, V (n,n)."""
        m, n = M.shape
        if m < n or n > ProjectorCalculator.FULL_RANK_MAX:
            raise NotImplementedError(f"backend='b200': svd_type='full-rank' supports m >= n <= {ProjectorCalculator.FULL_RANK_MAX} "
                                      f"(got {m}x{n}); use svd_type='rsvd'")
        Qb = ops.orthonormalize(M.clone())
        Rc = ops.matmul(Qb, M, transpose_a=True)                  # (n, n)
        S, Wt, Jt, info = ops.jacobi_svd(Rc)
        U = ops.matmul(Qb, Jt.t().contiguous())
        return U, S, Wt.t().contiguous()

    def _full_rank_projectors(self, R1, R2, F, d1, d4, chi):
        """calculate_projectors (projectors.py:114-136) for explicitly formed R1 (m, .), R2 (., n) and F = R1 R2."""
        U, S, V = self.full_svd(F)
        s = S / S[0]
        keep = min(chi, int((s > self.svd_cutoff).sum().item()))
        if self.spectra is not None:
            self.spectra.append(s.detach().cpu())
        w = 1.0 / torch.sqrt(s[:keep])
        Us = (U[:, :keep] * w).contiguous()
        Vs = (V[:, :keep] * w).contiguous()
        p1 = ops.matmul(R1, Us, transpose_a=True)                 # einsum("xedD,xz->edDz", R1, conj(U))
        p2 = ops.matmul(R2, Vs)                                   # einsum("cuUy,yz->cuUz", R2, V)
        return p1.view(*d1[3:], keep), p2.view(*d4[:3], keep)

    @staticmethod
    def _normalized(M):
        mx = torch.zeros(1, dtype=M.dtype, device=M.device)
        ops.absmax(M, mx)
        return M / mx

    def _full_rank_half_system(self, ipeps, sites, k):
        """contract_half_system + calculate_projectors (projectors.py:62-82, 138-142)."""
        Q1, d1 = self.make_quarter_tensor(ipeps[sites[0]], k)
        Q4, d4 = self.make_quarter_tensor(ipeps[sites[3]], k + 3)
        R = self._normalized(ops.matmul(Q1, Q4))
        return self._full_rank_projectors(Q1, Q4, R, d1, d4, ipeps.dims["chi"])

    def _full_rank_full_system(self, ipeps, sites, k):
        """contract_full_system + calculate_projectors (projectors.py:84-112, 176-179)."""
        s1, s2, s3, s4 = sites
        Q1, d1 = self.make_quarter_tensor(ipeps[s1], k)
        Q2, _ = self.make_quarter_tensor(ipeps[s2], k + 1)
        Q3, _ = self.make_quarter_tensor(ipeps[s3], k + 2)
        Q4, d4 = self.make_quarter_tensor(ipeps[s4], k + 3)
        R1 = self._normalized(ops.matmul(Q2, Q1))
        R2 = self._normalized(ops.matmul(Q4, Q3))
        F = self._normalized(ops.matmul(R1, R2))
        return self._full_rank_projectors(R1, R2, F, d1, d4, ipeps.dims["chi"])

    # ---- phase 1 ------------------------------------------------------------------------------------------------
    def begin_half_system(self, ipeps, sites, k, stream=None, omega=None, bulk=None, slot=None, info=None):
        """projectors.py:138-161 : Q1, Q4, rSVD of Q1 @ Q4 (never formed).
        bulk: optional second CUDA stream for the throughput-bound first stage (quarter tensors + their K7 encodings); the rSVD
        chain (K7 products interleaved with ~500 latency-bound TSQR / Jacobi launches) then runs on `stream`, ordered behind
        the first stage by an event.  DirectionalMover.move_pair passes one low-priority bulk stream for all tasks of a phase
        and a high-priority `stream` per task.
        slot: optional TaskSlot whose grow-only buffers receive the two quarter tensors and their K7 encodings, so that a phase
        with more tasks than slots re-uses them instead of holding every task's 12 GiB until the phase ends."""
        self._check_svd_type()
        if self.svd_type == "full-rank":
            return {"kind": "done", "result": self._full_rank_half_system(ipeps, sites, k), "stream": None}
        s1, s4 = sites[0], sites[3]
        st1, st4 = ipeps[s1], ipeps[s4]
        chi = ipeps.dims["chi"]
        if omega is None:
            omega = self.draw_omega(ipeps, sites, k)
        sa = bulk if (bulk is not None and stream is not None) else stream
        mx = torch.empty(2, dtype=omega.dtype, device=omega.device)   # zeroed by the library on the side stream
        dev = omega.device
        n1, n4 = self.quarter_numel(st1, k), self.quarter_numel(st4, k + 3)
        o1 = slot.buffer("Q1", n1, omega.dtype, dev) if slot is not None else None
        o4 = slot.buffer("Q4", n4, omega.dtype, dev) if slot is not None else None
        m1, m4 = self.quarter_shape(st1, k), self.quarter_shape(st4, k + 3)
        encs = None
        if self._use_i8([m1, m4], omega.shape[1]):
            # K7: both quarter tensors are encoded once (16 int8 residue planes) and serve all 13 thin products; the encoding is
            # produced by the same library call that builds the tensor (column exponents from the producing kernel's epilogue)
            def storage(name, shape):
                nb = ops.i8_encoded_bytes(*shape)
                return slot.buffer(name, nb, torch.uint8, dev) if slot is not None else torch.empty(nb, dtype=torch.uint8, device=dev)
            Q1, q1D, e1 = self.make_quarter_tensor(st1, k, normalize=False, stream=sa, absmax=mx[0:1], out=o1, enc_storage=storage("enc1", m1))
            Q4, q4D, e4 = self.make_quarter_tensor(st4, k + 3, normalize=False, stream=sa, absmax=mx[1:2], out=o4, enc_storage=storage("enc4", m4))
            encs = [e1, e4]
        else:
            Q1, q1D = self.make_quarter_tensor(st1, k, normalize=False, stream=sa, absmax=mx[0:1], out=o1)
            Q4, q4D = self.make_quarter_tensor(st4, k + 3, normalize=False, stream=sa, absmax=mx[1:2], out=o4)
        if sa is not stream:
            built = torch.cuda.Event()
            built.record(sa)
            stream.wait_event(built)
        # U is never materialised: proj1 = Q1^T U = (Q1^T Qy) U_B reuses the first product of the final adjoint pass
        _, S, V, info, AtQ, Wt = ops.rsvd([Q1, Q4], omega, niter=self.rsvd_niter, reorth_adjoint=False, chi=chi,
                                          cutoff=self.svd_cutoff, stream=stream, want_u=False, want_atq=True, encs=encs, info=info)
        return {"kind": "half", "mx": mx, "Q1": Q1, "Q4": Q4, "q1D": q1D, "q4D": q4D, "U": None, "S": S, "V": V, "info": info,
                "AtQ": AtQ, "Wt": Wt, "omega": omega, "stream": stream, "encs": encs}

    def begin_full_system(self, ipeps, sites, k, stream=None, omega=None, bulk=None, slot=None):
        """projectors.py:176-201 : rSVD of (Q2 Q1)(Q4 Q3)."""
        self._check_svd_type()
        if self.svd_type == "full-rank":
            return {"kind": "done", "result": self._full_rank_full_system(ipeps, sites, k), "stream": None}
        s1, s2, s3, s4 = sites
        chi = ipeps.dims["chi"]
        if omega is None:
            omega = self.draw_omega(ipeps, sites, k)
        Q1, q1D = self.make_quarter_tensor(ipeps[s1], k, normalize=True, stream=stream)
        Q2, _ = self.make_quarter_tensor(ipeps[s2], k + 1, normalize=True, stream=stream)
        Q3, _ = self.make_quarter_tensor(ipeps[s3], k + 2, normalize=True, stream=stream)
        Q4, q4D = self.make_quarter_tensor(ipeps[s4], k + 3, normalize=True, stream=stream)
        encs = None
        if self._use_i8([tuple(Q.shape) for Q in (Q2, Q1, Q4, Q3)], omega.shape[1]):
            encs = [ops.i8_encode(Q, stream=stream) for Q in (Q2, Q1, Q4, Q3)]
        U, S, V, info = ops.rsvd([Q2, Q1, Q4, Q3], omega, niter=self.rsvd_niter, reorth_adjoint=True, chi=chi,
                                 cutoff=self.svd_cutoff, stream=stream, encs=encs)
        return {"kind": "full", "Q1": Q1, "Q2": Q2, "Q3": Q3, "Q4": Q4, "q1D": q1D, "q4D": q4D, "U": U, "S": S, "V": V,
                "info": info, "omega": omega, "stream": stream, "encs": encs}

    def draw_omega_shape(self, ipeps, sites, k):
        """An uninitialised tensor with the shape / dtype / device of draw_omega's result (no random draw)."""
        chi, D = ipeps.dims["chi"], ipeps.dims["bond"]
        st1, st4 = ipeps[sites[0]], ipeps[sites[3]]
        m = st1['E'][(0 + k) % 4].shape[1] * D * D
        n = st4['E'][(3 + k + 3) % 4].shape[0] * D * D
        A = st1['A']
        return torch.empty(n, min(chi + self.rsvd_oversampling, m, n), dtype=A.dtype, device=A.device)

    def draw_omega(self, ipeps, sites, k):
        """The Gaussian test matrix of this projector, drawn exactly where/how the reference draws it
        (torch.randn(n, q) on the tensors' device, fused_matmul_svd_lowrank.py:32): one draw per projector, in call order."""
        chi = ipeps.dims["chi"]
        D = ipeps.dims["bond"]
        if self.svd_type == "full-rank":
            ipeps[sites[0]], ipeps[sites[3]]      # same ValueError on unknown sites; no random draw in this mode
            return None
        if self.projectors == "half-system":
            st1, st4 = ipeps[sites[0]], ipeps[sites[3]]
            m = st1['E'][(0 + k) % 4].shape[1] * D * D            # rows of Q1: chi_c of E[k]
            n = st4['E'][(3 + k + 3) % 4].shape[0] * D * D        # cols of Q4: chi_e of E[(3+(k+3))%4]
            kdim = st1['E'][(3 + k) % 4].shape[0] * D * D
            q = min(chi + self.rsvd_oversampling, m, n)
        else:
            st2, st3 = ipeps[sites[1]], ipeps[sites[2]]
            m = st2['E'][(k + 1) % 4].shape[1] * D * D            # rows of Q2
            n = st3['E'][(3 + k + 2) % 4].shape[0] * D * D        # cols of Q3
            q = min(chi + self.rsvd_oversampling, m, n)
        A = ipeps[sites[0]]['A']
        return linalg._omega(n, q, A.dtype, A.device)

    # ---- phase 2 ------------------------------------------------------------------------------------------------
    def finish(self, pend):
        if pend["kind"] == "done":
            return pend["result"]
        stream = pend["stream"]
        if stream is not None:
            stream.synchronize()
        keep = int(pend["info"][0].item())   # projectors.py:163-164 : the one host sync per projector
        S = pend["S"]
        if self.spectra is not None:
            self.spectra.append((S / S[0]).detach().cpu())
        q1D, q4D = pend["q1D"], pend["q4D"]
        if pend["kind"] == "half":
            mx = pend["mx"]
            encs = pend.get("encs")
            p1, p2 = ops.projectors_from_usv(pend["Q1"], pend["Q4"], None, pend["V"], S, keep, stream=stream,
                                             qmax1=mx[0:1], qmax4=mx[1:2], AtQ=pend["AtQ"], Wt=pend["Wt"],
                                             enc4=encs[1] if encs is not None else None)
        else:
            # projectors.py:209-217 : proj1 = Q1^H (Q2^H conj(U)), proj2 = Q4 (Q3 V), columns scaled by s^-1/2
            # (runs on the main stream; finish() already synchronised the side stream)
            w = 1.0 / torch.sqrt(S[:keep] / S[0])
            Us = (pend["U"][:, :keep] * w).contiguous()
            Vs = (pend["V"][:, :keep] * w).contiguous()
            encs = pend.get("encs")
            if encs is not None and keep >= 1:
                p1 = ops.i8_matmul(encs[1], ops.i8_matmul(encs[0], Us, adjoint=True), adjoint=True)
                p2 = ops.i8_matmul(encs[2], ops.i8_matmul(encs[3], Vs))
            else:
                p1 = ops.matmul(pend["Q1"], ops.matmul(pend["Q2"], Us, transpose_a=True), transpose_a=True)
                p2 = ops.matmul(pend["Q4"], ops.matmul(pend["Q3"], Vs))
        if stream is not None:
            pend["ready"] = torch.cuda.Event()
            pend["ready"].record(stream)       # the projector pair is complete on the side stream
        return p1.view(*q1D[3:], keep), p2.view(*q4D[:3], keep)

    def finish_static(self, pend, keep):
        """finish() for a CAPTURED move (half-system rSVD only): the truncated rank is not read back -- the projector pair is formed
        for the rank `keep` the eager run of the same move produced, and the caller compares info[0] with it after the replay."""
        mx, encs = pend["mx"], pend.get("encs")
        p1, p2 = ops.projectors_from_usv(pend["Q1"], pend["Q4"], None, pend["V"], pend["S"], keep, stream=pend["stream"],
                                         qmax1=mx[0:1], qmax4=mx[1:2], AtQ=pend["AtQ"], Wt=pend["Wt"],
                                         enc4=encs[1] if encs is not None else None)
        return p1.view(*pend["q1D"][3:], keep), p2.view(*pend["q4D"][:3], keep)

    def calculate_half_system(self, ipeps, sites, k):
        """projectors.py:138-174."""
        return self.finish(self.begin_half_system(ipeps, sites, k))

    def calculate_full_system(self, ipeps, sites, k):
        """projectors.py:176-217 (rsvd branch)."""
        return self.finish(self.begin_full_system(ipeps, sites, k))


class TaskSlot:
    """Grow-only device buffers of one in-flight projector task (two quarter tensors + their K7 encodings = 12 GiB at D=8,
    chi=256).  A phase with more tasks than slots re-uses a slot once the previous task's projector pair is complete (`ready`),
    so peak memory is bounded by the window, not by the number of tasks of the phase (the reference holds one projector's
    quarter tensors at a time, projectors.py:138-161)."""

    def __init__(self):
        self.bufs = {}
        self.ready = None        # event: the last task that used this slot has finished reading its buffers

    def buffer(self, name, numel, dtype, device):
        buf = self.bufs.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != dtype or buf.device != device:
            if buf is not None:
                torch.cuda.synchronize(device)       # rare (chi still growing): kernels of a side stream may still read the old one
            self.bufs[name] = None
            buf = torch.empty(int(numel), dtype=dtype, device=device)
            self.bufs[name] = buf
        return buf


class _ArenaSite:
    """Fixed-address copies of one site's tensors (what captured moves read and, through `commit`, write).  One per site and
    mover, shared by all captured phases, so that in steady state no boundary tensor is copied between phases: the site's C / E
    list items ARE these tensors; only the 64 KiB site tensor `A` is copied in before every replay."""

    def __init__(self, st):
        self.A = st['A'].detach().clone()
        self.C = [c.detach().clone().contiguous() for c in st['C']]
        self.E = [e.detach().clone().contiguous() for e in st['E']]

    def __getitem__(self, key):
        return {'A': self.A, 'C': self.C, 'E': self.E}[key]

    def bond_permute(self, k):
        return self.A.permute([(i + k) % 4 for i in range(4)] + [4])

    def matches(self, st):
        return st['A'].shape == self.A.shape and all(st['C'][k].shape == self.C[k].shape and st['E'][k].shape == self.E[k].shape
                                                     for k in range(4))

    def adopt(self, st):
        """Make the arena hold the site's current tensors, and the site's C / E list items BE the arena tensors."""
        # `A` is the caller's object (the reference's setter clones, site_tensor.py:77): it cannot be adopted, and neither its address
        # (allocator reuse) nor a version counter (inference tensors have none) tells whether a bond update replaced it -> copy, 1 launch
        if st['A'].data_ptr() != self.A.data_ptr():
            self.A.copy_(st['A'])
        for k in range(4):
            for name, mine in (('C', self.C), ('E', self.E)):
                cur = st[name][k]
                if cur.data_ptr() != mine[k].data_ptr():
                    mine[k].copy_(cur)
                    st[name][k] = mine[k]


class MoveGraph:
    """One phase (a directional move, or a left+right / up+down pair) captured as a CUDA graph, for the launch-bound regime
    (BASELINE configs 1-2: a site-move at D=2, chi=20 is ~140 dependent launches of a few microseconds each; SURVEY.md 7 hard
    part 6, 8f-3).  Built only after an eager run of the same phase left every tensor it touches saturated (chi legs == chi)
    and every projector at full rank, so all shapes are static:

      * the site / boundary tensors of the lines involved live in an arena of fixed-address buffers; tensors the caller
        replaced since the last run (e.g. `A` after a bond update) are copied in before the replay;
      * Omega is drawn OUTSIDE the graph with the same torch.randn calls, in the same order, as the eager path;
      * the graph holds every kernel of the phase: quarter tensors, rSVD chains, projector GEMMs for the expected rank,
        absorptions into staging tensors -- no host interaction;
      * after the replay ONE host read checks that every projector came out with the expected rank (the eager path reads
        once per projector, projectors.py:164); then the staged C, C, E are committed into the arena (whose tensors ARE the
        ipeps list items).  If a rank differs the replay is discarded -- nothing was committed -- and the phase re-runs eagerly.
    Same kernels, launch parameters and order per tensor as the eager single-stream schedule: bit-identical results."""

    def __init__(self, mover, ipeps, groups):
        self.mover, self.groups = mover, groups
        self.tasks = [t for g in groups for t in g]
        pc = mover.projector_calculator
        self.chi = ipeps.dims["chi"]
        sites = []
        for t in self.tasks:
            for s in [t["s1"], t["s2"]] + list(t["plaq"]):
                if tuple(s) not in sites:
                    sites.append(tuple(s))
        self.sites = sites
        self.arena = {s: mover.arena_site(ipeps, s) for s in sites}
        for s in sites:
            self.arena[s].adopt(ipeps[s])
        dev = self.arena[sites[0]].A.device
        self.device = dev
        view = _ArenaCell(ipeps, self.arena)
        self.omega = [torch.empty_like(pc.draw_omega_shape(view, t["plaq"], t["k"])) for t in self.tasks]
        self.infos = torch.zeros(len(self.tasks), 2, dtype=torch.int32, device=dev)
        self.stream = mover.graph_stream(dev)
        ops.reserve_workspace(dev, self.stream, ops.workspace_high_water(dev))
        self.graph = torch.cuda.CUDAGraph()
        # garbage that owns CUDA resources (a dropped mover's graphs, streams) must not be finalised in the middle of the capture:
        # freeing device memory / destroying a graph is not allowed while a stream of this thread captures
        import gc
        gc.collect()
        # warm-up on the capture stream (results discarded): grows the scratch buffer to what this single-stream order of the phase
        # needs and makes every kernel resident, neither of which may happen inside a capture
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            self._enqueue(view)
        torch.cuda.synchronize(dev)
        self._keep = None
        gc.disable()
        n0 = ops.launch_count()
        try:
            with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
                self.staged = self._enqueue(view)
        finally:
            gc.enable()
        self.n_kernels = ops.launch_count() - n0      # kernel nodes of the library in the graph (recorded, not run)
        ops.note_replayed_launches(-self.n_kernels)   # ... which the library counted as launches while they were being captured
        self._keep = None            # capture is over: the graph's private pool keeps the addresses of the intermediates
        self.replays = 0

    def _enqueue(self, view):
        """Every kernel of the phase.  The independent tasks are forked onto side streams and joined again (captured as parallel
        branches of the graph), like the eager schedule overlaps their latency-bound chains.  All intermediates stay alive until the
        caller drops `self._keep`: a block freed in mid-capture could be handed to a kernel of a concurrent branch."""
        mover, pc, chi = self.mover, self.mover.projector_calculator, self.chi
        main = torch.cuda.current_stream(self.device)
        side = [st for st in mover._side_streams(self.device) if st is not None] or [main]
        keep, staged = [], []
        n = 0
        used = []
        proj = []
        # every projector task of the phase forks off first (the moves of a paired phase are independent: all of them read the
        # pre-phase state and the commit happens after the replay), one join, then every absorption
        for g in self.groups:
            p1, p2 = {}, {}
            for t in g:
                st = side[n % len(side)]
                if st is not main and st not in used:
                    st.wait_stream(main)
                    used.append(st)
                with torch.cuda.stream(st):
                    pend = pc.begin_half_system(view, t["plaq"], t["k"], stream=st if st is not main else None, omega=self.omega[n],
                                                info=self.infos[n])
                    p1[t["key"]], p2[t["key"]] = pc.finish_static(pend, chi)
                keep.append(pend)
                n += 1
            proj.append((p1, p2))
        for st in used:
            main.wait_stream(st)
        used = []
        i = 0
        for g, (p1, p2) in zip(self.groups, proj):
            for t in g:
                k, src = t["k"], view[t["s1"]]
                st = side[i % len(side)]
                if st is not main and st not in used:
                    st.wait_stream(main)
                    used.append(st)
                with torch.cuda.stream(st):
                    c1 = mover.renormalize_cj1(src['C'][(3 + k) % 4], src['E'][(2 + k) % 4], p1[t["i"]])
                    c2 = mover.renormalize_cj2(src['C'][k], src['E'][k], p2[t["j"]])
                    e = mover.renormalize_ej(src['E'][(3 + k) % 4], src.bond_permute(k), p2[t["i"]], p1[t["j"]])
                staged.append((t["s2"], k, c1, c2, e))
                i += 1
        for st in used:
            main.wait_stream(st)
        keep.append(proj)
        self._keep = keep
        return staged

    def usable(self, ipeps):
        return ipeps.dims["chi"] == self.chi and all(self.mover._arena.get(s) is self.arena[s] and self.arena[s].matches(ipeps[s])
                                                     for s in self.sites)

    def run(self, ipeps):
        """Replay; True when the result was committed, False when a projector rank differed (nothing changed)."""
        pc = self.mover.projector_calculator
        for s in self.sites:
            self.arena[s].adopt(ipeps[s])
        view = _ArenaCell(ipeps, self.arena)
        for n, t in enumerate(self.tasks):                  # the reference's draws, in the reference's order
            self.omega[n].copy_(pc.draw_omega(view, t["plaq"], t["k"]))
        self.graph.replay()
        ops.note_replayed_launches(self.n_kernels)
        if [int(v) for v in self.infos[:, 0].tolist()] != [self.chi] * len(self.tasks):      # the one host read of the phase
            return False
        for s2, k, c1, c2, e in self.staged:
            ar = self.arena[tuple(s2)]
            ar.C[(3 + k) % 4].copy_(c1)
            ar.C[k].copy_(c2)
            ar.E[(3 + k) % 4].copy_(e)
        self.replays += 1
        return True


class _ArenaCell:
    """ipeps-like view whose sites are the arena copies (dims / nx / ny from the real object)."""

    def __init__(self, ipeps, arena):
        self.dims, self.nx, self.ny, self._arena, self._ipeps = ipeps.dims, ipeps.nx, ipeps.ny, arena, ipeps

    def __getitem__(self, site):
        ar = self._arena.get(tuple(site))
        return ar if ar is not None else self._ipeps[site]


class DirectionalMover:
    """acetn/renormalization/directional_mover.py:5-366 (non-distributed moves).

    The ny (nx) projector computations of a move are independent (directional_mover.py:23-40 computes them in a plain
    loop); here they are enqueued round-robin on `n_streams` CUDA streams so that one site's latency-bound stages
    (TSQR panels, Jacobi core) overlap the other site's DGEMMs.  Omega is still drawn in the reference's order."""

    def __init__(self, config, n_streams=None):
        self.projector_calculator = ProjectorCalculator(config)
        self.calculate_projectors = self.projector_calculator.calculate
        self.n_streams = int(os.environ.get("ACETN_B200_STREAMS", "4")) if n_streams is None else n_streams
        self._streams = None
        self._bulk_stream = None
        # staggered phases (half-system rSVD): the throughput-bound stages of all tasks (quarter tensors + encodings, then the
        # absorptions of each move as soon as its projectors exist) are queued in task order on ONE low-priority "bulk" stream;
        # each task's rSVD chain runs on its own HIGH-priority stream.  The block scheduler then serves the ~500 small
        # latency-bound launches of a chain (TSQR, Jacobi, CRT) ahead of the pending CTAs of the next task's DGEMMs instead
        # of queueing them behind whole 4 ms kernels, and the tasks no longer go through their latency-bound stages in lockstep.
        self.stagger = os.environ.get("ACETN_B200_STAGGER", "1") != "0"
        # at most this many projector tasks are begun-but-unfinished at any time (each holds 2 quarter tensors + encodings)
        self.inflight = max(1, int(os.environ.get("ACETN_B200_INFLIGHT", "4")))
        # CUDA-graph replay of whole phases in the launch-bound regime (MoveGraph): "auto" = quarter tensors below GRAPH_MAX_DIM
        self.use_graphs = os.environ.get("ACETN_B200_GRAPHS", "auto")
        self._graphs, self._graph_ok, self._graph_stream, self._arena = {}, {}, None, {}
        self.graph_replays = 0
        self._slots = []
        self._config = config
        self._sharded = None
        self.split_edge = os.environ.get("ACETN_B200_SPLIT_EDGE", "1") != "0"     # piecewise absorptions (see _finish_and_absorb_piecewise)

    def _side_streams(self, device):
        if self.n_streams <= 1:
            return [None]
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=device) for _ in range(self.n_streams)]
        return self._streams

    def staggered(self):
        """True when the phase schedule with one bulk stream + one high-priority stream per rSVD chain applies."""
        pc = self.projector_calculator
        return self.stagger and self.n_streams > 1 and pc.projectors == "half-system" and pc.svd_type == "rsvd"

    def phase_streams(self, device, ntasks):
        """(bulk, [chain streams]): the low-priority stream of the throughput-bound stages and one high-priority stream per
        task (at most 16; finish() synchronises a whole stream, so chains do not share one unless there are more tasks)."""
        if self._bulk_stream is None:
            self._bulk_stream = torch.cuda.Stream(device=device, priority=0)
            self._hi_streams = []
        n = max(1, min(ntasks, 16))
        while len(self._hi_streams) < n:
            self._hi_streams.append(torch.cuda.Stream(device=device, priority=-1))
        return self._bulk_stream, self._hi_streams[:n]

    def release(self):
        """Drop the task-slot buffers and captured graphs (they are re-created on demand)."""
        self._slots = []
        self._graphs, self._graph_ok, self._arena = {}, {}, {}

    # chi * D^2 up to which a phase is replayed as a CUDA graph ("auto").  Measured (tools/small_config_bench.py / bench.py, 2x2 cell,
    # eager -> graph, after all tasks of a paired phase fork as parallel branches): D=2 chi=20 (chi D^2 = 80) 149 -> 312 sweeps/s,
    # D=4 chi=64 (1024) 76 -> 83, D=6 chi=36 (1296) 87 -> 103, D=5 chi=80 (2000) 42.3 -> 42.9, D=6 chi=64 (2304) 41.3 -> 42.2,
    # D=8 chi=48 (3072) 36.6 -> 38.8.  From 4096 on the thin products move to K7 and the sweep is throughput-bound: eager schedule.
    GRAPH_MAX_DIM = 3072

    def arena_site(self, ipeps, site):
        """The fixed-address buffers of `site` (created from its current, saturated tensors; rebuilt when a shape changed, which
        also drops every captured phase that used the old buffers -- MoveGraph.usable)."""
        ar = self._arena.get(site)
        if ar is None or not ar.matches(ipeps[site]):
            ar = self._arena[site] = _ArenaSite(ipeps[site])
        return ar

    def graph_stream(self, device):
        if self._graph_stream is None:
            self._graph_stream = torch.cuda.Stream(device=device)
        return self._graph_stream

    def _graph_candidate(self, ipeps, groups):
        """True when the phase may run as a captured graph: half-system rSVD, launch-bound size, every tensor of the lines involved
        saturated, source and target line of every move disjoint, no recorder.  (A captured phase reads the pre-phase state
        throughout and commits afterwards, which equals the eager order because no task reads a slot another task writes.)"""
        pc = self.projector_calculator
        if self.use_graphs == "0" or pc.spectra is not None or pc.projectors != "half-system" or pc.svd_type != "rsvd":
            return False
        chi, D = ipeps.dims["chi"], ipeps.dims["bond"]
        if self.use_graphs != "1" and chi * D * D > self.GRAPH_MAX_DIM:
            return False
        if min(chi + pc.rsvd_oversampling, chi * D * D) <= chi:          # q must exceed chi for the expected rank to be chi
            return False
        tasks = [t for g in groups for t in g]
        for g in groups:      # per move: source line != target line (across the moves of a pair the tensor SLOTS are disjoint, see move_pair)
            if not {tuple(t["s1"]) for t in g}.isdisjoint({tuple(t["s2"]) for t in g}):
                return False
        for t in tasks:
            for s in [t["s1"], t["s2"]] + list(t["plaq"]):
                st = ipeps[s]
                if st['A'].device.type != "cuda":
                    return False
                for k in range(4):
                    if tuple(st['C'][k].shape) != (chi, chi) or tuple(st['E'][k].shape) != (chi, chi, D, D):
                        return False
        return True

    def _run_graphed(self, ipeps, moves):
        """Replay the phase `moves` from its captured graph when possible.  Returns True when done."""
        key = tuple(moves)
        if not self._graph_ok.get(key):
            return False
        groups = [self.move_tasks(ipeps, k, line) for k, line in moves]
        if not self._graph_candidate(ipeps, groups):
            return False
        g = self._graphs.get(key)
        if g is not None and not g.usable(ipeps):
            g = None
        if g is None:
            g = self._graphs[key] = MoveGraph(self, ipeps, groups)
        if g.run(ipeps):
            self.graph_replays += 1
            return True
        self._graph_ok[key] = False          # a projector lost rank: back to the eager path until a full-rank eager run re-arms it
        return False

    def _note_eager(self, ipeps, moves, ranks):
        """After an eager run of a phase: arm the graph path when every projector had full rank chi."""
        if self.use_graphs == "0":
            return
        chi = ipeps.dims["chi"]
        self._graph_ok[tuple(moves)] = bool(ranks) and all(r == chi for r in ranks)

    def _pipeline(self, ipeps, specs, omegas, streams, retired, bulk=None):
        """Generator over the projector tasks specs = [(sites, k)]: yields (n, pending, (proj1, proj2)) in task order while at most
        `inflight` tasks are begun-but-unfinished.  Task n + W is begun (on the slot of task n) when the consumer asks for the next
        item, i.e. after it has queued whatever it derives from task n -- same kernels in the same per-tensor order for any W.
        retired: caller-owned list that keeps the small per-task tensors (Omega, S, V, AtQ, ...) alive until the caller has
        ordered its main stream behind the side streams; only the 12 GiB of a task slot are recycled inside a phase."""
        pc = self.projector_calculator
        W = max(1, min(self.inflight, len(specs)))
        while len(self._slots) < W:
            self._slots.append(TaskSlot())
        pend = {}

        def begin(n):
            sites, k = specs[n]
            slot = self._slots[n % W]
            stream = streams[n % len(streams)]
            first = bulk if (bulk is not None and stream is not None) else stream
            if slot.ready is not None and first is not None:
                first.wait_event(slot.ready)
            pend[n] = pc.begin(ipeps, sites, k, stream=stream, omega=omegas[n], bulk=bulk, slot=slot)

        for n in range(W):
            begin(n)
        for n in range(len(specs)):
            pd = pend.pop(n)
            res = pc.finish(pd)
            self._slots[n % W].ready = pd.get("ready")
            retired.append(pd)
            yield n, pd, res
            del pd
            if n + W < len(specs):
                begin(n + W)

    def _projectors_of_line(self, ipeps, plaquettes, k):
        """All projector pairs of one move: {key: (proj1, proj2)} ; plaquettes = [(key, sites)]."""
        pc = self.projector_calculator
        device = ipeps[plaquettes[0][1][0]]['A'].device
        ops.require_cuda_device(device)
        streams = self._side_streams(device)
        omegas = [pc.draw_omega(ipeps, sites, k) for _, sites in plaquettes]     # reference order of the RNG draws
        main = torch.cuda.current_stream(device)
        for st in streams:
            if st is not None:
                st.wait_stream(main)
        out, retired = {}, []
        for n, _, res in self._pipeline(ipeps, [(sites, k) for _, sites in plaquettes], omegas, streams, retired):
            out[plaquettes[n][0]] = res
        for st in streams:
            if st is not None:
                main.wait_stream(st)
        del retired                  # per-task tensors are released only after the main stream is ordered behind the side streams
        return {key: v[0] for key, v in out.items()}, {key: v[1] for key, v in out.items()}

    # ---- task view of the moves -----------------------------------------------------------------------------------
    @staticmethod
    def move_tasks(ipeps, k, line):
        """The independent site tasks of one directional move (directional_mover.py:23-97 + pickers :99-181):
        dicts with the projector plaquette, source/target sites and the projector keys used by renormalize_boundary."""
        nx, ny = ipeps.nx, ipeps.ny
        tasks = []
        if k == 0:
            for yi in range(ny):
                xj, yj = (line + 1) % nx, (yi - 1 + ny) % ny
                tasks.append(dict(k=0, line=line, key=yi, plaq=[(line, yi), (xj, yi), (xj, yj), (line, yj)], s1=(line, yi),
                                  s2=(xj, yi), i=yi, j=(yi + 1) % ny))
        elif k == 2:
            for yi in range(ny):
                xj, yj = (line - 1 + nx) % nx, (yi + 1) % ny
                tasks.append(dict(k=2, line=line, key=yi, plaq=[(line, yi), (xj, yi), (xj, yj), (line, yj)], s1=(line, yi),
                                  s2=(xj, yi), i=yi, j=(yi - 1 + ny) % ny))
        elif k == 1:
            for xi in range(nx):
                xj, yj = (xi - 1 + nx) % nx, (line - 1 + ny) % ny
                tasks.append(dict(k=1, line=line, key=xi, plaq=[(xi, line), (xi, yj), (xj, yj), (xj, line)], s1=(xi, line),
                                  s2=(xi, yj), i=xi, j=(xi + 1) % nx))
        elif k == 3:
            for xi in range(nx):
                xj, yj = (xi + 1) % nx, (line + 1) % ny
                tasks.append(dict(k=3, line=line, key=xi, plaq=[(xi, line), (xi, yj), (xj, yj), (xj, line)], s1=(xi, line),
                                  s2=(xi, yj), i=xi, j=(xi - 1 + nx) % nx))
        else:
            raise ValueError(f"Invalid bond direction k={k}")
        return tasks

    def move_pair(self, ipeps, moves):
        """Several directional moves whose tasks are mutually independent, run as one phase: all projectors (on the
        side streams), then all absorptions.  In half-system mode a left and a right move (an up and a down move)
        read and write disjoint boundary tensors, which is what the reference's distributed schedule relies on
        (directional_mover.py:183-271); the Omega draws keep the sequential order (first move's sites, then the
        second's)."""
        if self._run_graphed(ipeps, moves):
            return
        groups = [self.move_tasks(ipeps, k, line) for k, line in moves]
        tasks = [t for g in groups for t in g]
        if not (self.staggered() and len(groups) > 1):
            p1, p2 = self._projectors_of_tasks(ipeps, tasks)
            for t in tasks:
                self._absorb_task(ipeps, t, p1, p2)
            self._note_eager(ipeps, moves, [p.shape[-1] for p in p1.values()])
            return
        # staggered schedule: the moves of the phase touch disjoint boundary tensors (see above), so the absorptions of a move
        # may run -- on their own stream -- while the projectors of the following moves are still being computed
        pc = self.projector_calculator
        device = ipeps[tasks[0]["s1"]]['A'].device
        ops.require_cuda_device(device)
        bulk, streams = self.phase_streams(device, min(len(tasks), self.inflight))
        omegas = [pc.draw_omega(ipeps, t["plaq"], t["k"]) for t in tasks]      # reference order of the RNG draws
        main = torch.cuda.current_stream(device)
        for st in streams + [bulk]:
            st.wait_stream(main)
        retired = []
        pipe = self._pipeline(ipeps, [(t["plaq"], t["k"]) for t in tasks], omegas, streams, retired, bulk)
        p1, p2 = {}, {}
        for g in groups:
            # the absorptions of a move read its source line and write the neighbouring line; when the two do not share a site
            # (every cell with at least two columns / rows) they can be issued piecewise, in any order
            piecewise = self.split_edge and {t["s1"] for t in g}.isdisjoint({t["s2"] for t in g})
            if piecewise:
                self._finish_and_absorb_piecewise(ipeps, g, pipe, p1, p2, bulk)
                continue
            for t in g:
                _, pd, (p1[(t["k"], t["key"])], p2[(t["k"], t["key"])]) = next(pipe)
                if pd.get("ready") is not None:
                    bulk.wait_event(pd["ready"])
            with torch.cuda.stream(bulk):          # outputs and scratch come from this stream's pool
                for t in g:
                    self._absorb_task(ipeps, t, p1, p2)
        pipe.close()
        for st in streams + [bulk]:
            main.wait_stream(st)
        del retired                  # per-task tensors are released only after the main stream is ordered behind the side streams
        self._note_eager(ipeps, moves, [p.shape[-1] for p in p1.values()])

    def _finish_and_absorb_piecewise(self, ipeps, g, pipe, p1, p2, bulk):
        """renormalize_boundary (directional_mover.py:293-303) of one move, issued on the bulk stream as soon as its inputs exist:
        the projector pair of task n is all that corner 1 of task n, corner 2 of its neighbour and the first stage of the
        neighbour's edge absorption need (directional_mover.py:295, 299, 301-303: proj1[i], proj2[j], proj1[j]); only the last
        GEMM of every edge absorption (proj2[i]) waits for the task's own chain.  After the last chain of a move, 2 x 4 ms of
        absorption work are left instead of 2 x 13.5 ms.  Same kernels in the same per-tensor order: bit-identical results."""
        t3 = {}
        for t in g:
            k, key = t["k"], t["key"]
            _, pd, (p1[(k, key)], p2[(k, key)]) = next(pipe)
            if pd.get("ready") is not None:
                bulk.wait_event(pd["ready"])
            with torch.cuda.stream(bulk):          # outputs and scratch come from this stream's pool
                src, dst = ipeps[t["s1"]], ipeps[t["s2"]]
                dst['C'][(3 + k) % 4] = self.renormalize_cj1(src['C'][(3 + k) % 4], src['E'][(2 + k) % 4], p1[(k, key)])
                for t2 in g:
                    if t2["j"] != key:
                        continue
                    src2, dst2 = ipeps[t2["s1"]], ipeps[t2["s2"]]
                    dst2['C'][k] = self.renormalize_cj2(src2['C'][k], src2['E'][k], p2[(k, key)])
                    t3[t2["key"]] = ops.absorb_edge_begin(src2['E'][(3 + k) % 4], src2.bond_permute(k), p1[(k, key)])
        with torch.cuda.stream(bulk):
            for t in g:
                k = t["k"]
                e_old = ipeps[t["s1"]]['E'][(3 + k) % 4]
                ipeps[t["s2"]]['E'][(3 + k) % 4] = ops.absorb_edge_finish(t3.pop(t["key"]), p2[(k, t["i"])], e_old.shape[2])

    def _absorb_task(self, ipeps, t, p1, p2):
        k = t["k"]
        self.renormalize_boundary(ipeps, {t["i"]: p1[(k, t["i"])], t["j"]: p1[(k, t["j"])]},
                                  {t["i"]: p2[(k, t["i"])], t["j"]: p2[(k, t["j"])]}, t["s1"], t["s2"], t["i"], t["j"], k)

    def left_right_move(self, ipeps, x1, x2):
        """Single-process counterpart of left_right_move_dist (directional_mover.py:183-226)."""
        self.move_pair(ipeps, [(0, x1), (2, x2)])

    def up_down_move(self, ipeps, y1, y2):
        """Single-process counterpart of up_down_move_dist (directional_mover.py:228-271)."""
        self.move_pair(ipeps, [(1, y1), (3, y2)])

    # ---- the moves `ipeps.evolve` issues after every bond update (acetn/evolution/fast_full_update.py:72-129) -----------------------
    def absorb_bond(self, ipeps, bond):
        """FastFullUpdater.absorb_bond (fast_full_update.py:72-99): the two opposite moves that absorb the updated bond, in the
        reference's order (so the Omega draws keep their sequence); with half-system projectors they touch disjoint boundary
        tensors and are run as one phase."""
        s1, s2, k = bond
        moves = {0: [(2, s1[0]), (0, s2[0])], 1: [(3, s1[1]), (1, s2[1])],
                 2: [(0, s1[0]), (2, s2[0])], 3: [(1, s1[1]), (3, s2[1])]}[k]
        pc = self.projector_calculator
        if pc.projectors == "half-system" and os.environ.get("ACETN_B200_PAIR_MOVES", "1") != "0":
            self.move_pair(ipeps, moves)
            return
        do = {0: self.left_move, 1: self.up_move, 2: self.right_move, 3: self.down_move}
        for kk, line in moves:
            do[kk](ipeps, line)

    def _sharded_ctmrg(self, ipeps):
        if self._sharded is None or self._sharded.ipeps is not ipeps:
            from .distributed import B200Compute, ShardedCtmrg
            self._sharded = ShardedCtmrg(ipeps, self._config, ipeps.rank, ipeps.world_size, compute=B200Compute(self._config, mover=self))
        return self._sharded

    def left_right_move_dist(self, ipeps, x1, x2):
        """directional_mover.py:183-226 : left move on column x1 and right move on column x2 as one site-sharded phase."""
        self._sharded_ctmrg(ipeps).phase([(0, x1), (2, x2)])

    def up_down_move_dist(self, ipeps, y1, y2):
        """directional_mover.py:228-271 : up move on row y1 and down move on row y2 as one site-sharded phase."""
        self._sharded_ctmrg(ipeps).phase([(1, y1), (3, y2)])

    def _projectors_of_tasks(self, ipeps, tasks):
        pc = self.projector_calculator
        device = ipeps[tasks[0]["s1"]]['A'].device
        ops.require_cuda_device(device)
        streams = self._side_streams(device)
        omegas = [pc.draw_omega(ipeps, t["plaq"], t["k"]) for t in tasks]      # reference order of the RNG draws
        main = torch.cuda.current_stream(device)
        for st in streams:
            if st is not None:
                st.wait_stream(main)
        p1, p2, retired = {}, {}, []
        for n, _, res in self._pipeline(ipeps, [(t["plaq"], t["k"]) for t in tasks], omegas, streams, retired):
            t = tasks[n]
            p1[(t["k"], t["key"])], p2[(t["k"], t["key"])] = res
        for st in streams:
            if st is not None:
                main.wait_stream(st)
        del retired
        return p1, p2

    # ---- the four moves (directional_mover.py:23-97) -----------------------------------------------------------
    def left_move(self, ipeps, xi):
        if self._run_graphed(ipeps, [(0, xi)]):
            return
        nx, ny = ipeps.nx, ipeps.ny
        plaq = []
        for yi in range(ny):
            xj, yj = (xi + 1) % nx, (yi - 1 + ny) % ny
            plaq.append((yi, [(xi, yi), (xj, yi), (xj, yj), (xi, yj)]))
        proj1, proj2 = self._projectors_of_line(ipeps, plaq, 0)
        for yi in range(ny):
            xj = (xi + 1) % nx
            yj = (yi + 1) % ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xj, yi), yi, yj, k=0)
        self._note_eager(ipeps, [(0, xi)], [p.shape[-1] for p in proj1.values()])

    def up_move(self, ipeps, yi):
        if self._run_graphed(ipeps, [(1, yi)]):
            return
        nx, ny = ipeps.nx, ipeps.ny
        plaq = []
        for xi in range(nx):
            xj, yj = (xi - 1 + nx) % nx, (yi - 1 + ny) % ny
            plaq.append((xi, [(xi, yi), (xi, yj), (xj, yj), (xj, yi)]))
        proj1, proj2 = self._projectors_of_line(ipeps, plaq, 1)
        for xi in range(nx):
            xj = (xi + 1) % nx
            yj = (yi - 1 + ny) % ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xi, yj), xi, xj, k=1)
        self._note_eager(ipeps, [(1, yi)], [p.shape[-1] for p in proj1.values()])

    def right_move(self, ipeps, xi):
        if self._run_graphed(ipeps, [(2, xi)]):
            return
        nx, ny = ipeps.nx, ipeps.ny
        plaq = []
        for yi in range(ny):
            xj, yj = (xi - 1 + nx) % nx, (yi + 1) % ny
            plaq.append((yi, [(xi, yi), (xj, yi), (xj, yj), (xi, yj)]))
        proj1, proj2 = self._projectors_of_line(ipeps, plaq, 2)
        for yi in range(ny):
            xj = (xi - 1 + nx) % nx
            yj = (yi - 1 + ny) % ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xj, yi), yi, yj, k=2)
        self._note_eager(ipeps, [(2, xi)], [p.shape[-1] for p in proj1.values()])

    def down_move(self, ipeps, yi):
        if self._run_graphed(ipeps, [(3, yi)]):
            return
        nx, ny = ipeps.nx, ipeps.ny
        plaq = []
        for xi in range(nx):
            xj, yj = (xi + 1) % nx, (yi + 1) % ny
            plaq.append((xi, [(xi, yi), (xi, yj), (xj, yj), (xj, yi)]))
        proj1, proj2 = self._projectors_of_line(ipeps, plaq, 3)
        for xi in range(nx):
            xj = (xi - 1 + nx) % nx
            yj = (yi + 1) % ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xi, yj), xi, xj, k=3)
        self._note_eager(ipeps, [(3, yi)], [p.shape[-1] for p in proj1.values()])

    # ---- plaquette pickers (directional_mover.py:99-181) ----------------------------------------------------------
    def calculate_left_projectors(self, ipeps, xi, yi):
        xj, yj = (xi + 1) % ipeps.nx, (yi - 1 + ipeps.ny) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xj, yi), (xj, yj), (xi, yj)], k=0)

    def calculate_right_projectors(self, ipeps, xi, yi):
        xj, yj = (xi - 1 + ipeps.nx) % ipeps.nx, (yi + 1) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xj, yi), (xj, yj), (xi, yj)], k=2)

    def calculate_up_projectors(self, ipeps, xi, yi):
        xj, yj = (xi - 1 + ipeps.nx) % ipeps.nx, (yi - 1 + ipeps.ny) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xi, yj), (xj, yj), (xj, yi)], k=1)

    def calculate_down_projectors(self, ipeps, xi, yi):
        xj, yj = (xi + 1) % ipeps.nx, (yi + 1) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xi, yj), (xj, yj), (xj, yi)], k=3)

    # ---- absorption (directional_mover.py:273-366) ------------------------------------------------------------------
    def renormalize_boundary(self, ipeps, proj1, proj2, s1, s2, i, j, k):
        src, dst = ipeps[s1], ipeps[s2]
        dst['C'][(3 + k) % 4] = self.renormalize_cj1(src['C'][(3 + k) % 4], src['E'][(2 + k) % 4], proj1[i])
        dst['C'][k] = self.renormalize_cj2(src['C'][k], src['E'][k], proj2[j])
        dst['E'][(3 + k) % 4] = self.renormalize_ej(src['E'][(3 + k) % 4], src.bond_permute(k), proj2[i], proj1[j])

    @staticmethod
    def renormalize_cj1(ci, ei, proj):
        return ops.absorb_corner1(ci, ei, proj)

    @staticmethod
    def renormalize_cj2(ci, ei, proj):
        return ops.absorb_corner2(ci, ei, proj)

    @staticmethod
    def renormalize_ej(ei, ai, proj2, proj1):
        return ops.absorb_edge(ei, ai, proj2, proj1)


def ctmrg(ipeps, config, mover=None):
    """acetn/renormalization/ctmrg.py:4-31 (non-distributed ordering; the sharded schedule lives in
    acetn_b200.distributed)."""
    mover = mover or DirectionalMover(config)
    # half-system: the left/right (up/down) moves of one iteration are independent and are run as one phase
    paired = config.projectors == "half-system" and os.environ.get("ACETN_B200_PAIR_MOVES", "1") != "0"
    for _ in range(config.steps):
        for xi in range(ipeps.nx):
            if paired:
                mover.left_right_move(ipeps, xi, (ipeps.nx - xi + 1) % ipeps.nx)
            else:
                mover.left_move(ipeps, xi)
                mover.right_move(ipeps, (ipeps.nx - xi + 1) % ipeps.nx)
        for yi in range(ipeps.ny):
            if paired:
                mover.up_down_move(ipeps, (ipeps.ny - yi + 1) % ipeps.ny, yi)
            else:
                mover.up_move(ipeps, (ipeps.ny - yi + 1) % ipeps.ny)
                mover.down_move(ipeps, yi)
    return mover
