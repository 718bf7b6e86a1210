"""Site-sharded CTMRG over the GPUs of one box (replaces the reference's left_right_move_dist /
up_down_move_dist + all_gather_tensor, acetn/renormalization/directional_mover.py:183-271,
acetn/utils/distributed.py:85-107).

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).  The state is replicated, like in the
reference.  A *phase* is a left+right (or up+down) move pair on one line pair: its 2*ny (2*nx) site tasks are
independent in half-system mode.  Task t of a phase is owned by rank t % world:

  1. every rank draws ALL Omega matrices of the phase in the canonical (sequential) order and keeps its own
     (SURVEY.md App. D5), so results do not depend on the number of ranks;
  2. owners compute their projector pairs (quarter tensors, rSVD, truncation, projector GEMMs);
  3. the truncated ranks chi' are all-gathered (tiny), then each owner broadcasts its projector pair
     (2 x 32 MiB at D=8, chi=256) -- the absorption of task i needs the pair of its neighbour j;
  4. owners run the three absorptions of their tasks and broadcast the new (C, C, E) (about 33 MiB) so that every
     replica is current before the next phase.  Unlike the reference, absorptions are NOT replicated on every rank.

Communication is < 1 % of a phase at these sizes (about 100 MiB per task against >= 100 ms of compute), so plain NCCL
broadcasts are used; there is no compute step immediately followed by a collective on the same data that would
justify a fused kernel here (the GEMM epilogue producing a projector is followed by its host-side truncation read).

The compute backend is injected (`compute`), which lets the CPU/gloo tests drive this scheduler with the oracle's
functions and compare against the sequential sweep.
"""
import torch
import torch.distributed as dist


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


class B200Compute:
    """Default backend: libacetn_b200.so through acetn_b200.renormalization."""

    def __init__(self, config, mover=None):
        self.config = config
        from .renormalization import DirectionalMover
        self.mover = mover if mover is not None else DirectionalMover(config)
        self.pc = self.mover.projector_calculator

    def tasks(self, ipeps, k, line):
        return self.mover.move_tasks(ipeps, k, line)

    def draw_omega(self, ipeps, task):
        return self.pc.draw_omega(ipeps, task["plaq"], task["k"])

    def projectors(self, ipeps, tasks, omegas, group=None, group_rank=0, group_size=1):
        """Projector pairs of the given tasks -> list of (proj1, proj2).  With group_size > 1 every task is computed
        cooperatively by the ranks of `group` (row-sharded quarter tensors and rSVD products, sharded_projector.py)."""
        if not tasks:
            return []
        device = ipeps[tasks[0]["s1"]]['A'].device
        main = torch.cuda.current_stream(device)
        if group_size <= 1 and len(tasks) > 1 and self.mover.staggered():
            # several tasks on this rank: quarter tensors + encodings in task order on the low-priority bulk stream, every rSVD
            # chain on its own high-priority stream (renormalization.DirectionalMover.move_pair)
            bulk, streams = self.mover.phase_streams(device, min(len(tasks), self.mover.inflight))
            for st in streams + [bulk]:
                st.wait_stream(main)
            retired = []
            out = [res for _, _, res in self.mover._pipeline(ipeps, [(t["plaq"], t["k"]) for t in tasks], omegas, streams, retired, bulk)]
            for st in streams + [bulk]:
                main.wait_stream(st)
            del retired
            return [(a.contiguous(), b.contiguous()) for a, b in out]
        streams = self.mover._side_streams(device)
        for st in streams:
            if st is not None:
                st.wait_stream(main)
        pend = None
        if group_size <= 1:
            pend = []
            out = [res for _, _, res in self.mover._pipeline(ipeps, [(t["plaq"], t["k"]) for t in tasks], omegas, streams, pend)]
        else:
            from .sharded_projector import ShardedHalfSystemProjector
            sp = ShardedHalfSystemProjector(self.mover.projector_calculator, group, group_rank, group_size)
            sp.cfg = self.config
            sp.spectra = self.pc.spectra
            ctx = [torch.cuda.stream(streams[n % len(streams)]) if streams[0] is not None else _Null() for n in range(len(tasks))]
            pend = []
            for n, t in enumerate(tasks):
                with ctx[n]:
                    pend.append(sp.begin(ipeps, t["plaq"], t["k"], omegas[n]))
            out = []
            for n, pd in enumerate(pend):
                with (torch.cuda.stream(streams[n % len(streams)]) if streams[0] is not None else _Null()):
                    out.append(sp.finish(pd))
        for st in streams:
            if st is not None:
                main.wait_stream(st)
        del pend
        return [(a.contiguous(), b.contiguous()) for a, b in out]

    def absorb_coop(self, ipeps, task, p1i, p2i, p1j, p2j, group, g, G):
        """Cooperative absorption by the G ranks of a group: the two corners are tiny and replicated; the edge
        contraction is linear in the leg `a` shared by ei and proj2, so rank g contracts its a-block and the
        un-normalised partial results are all-reduced, then normalised identically on every rank."""
        k = task["k"]
        src = ipeps[task["s1"]]
        m = self.mover
        c1 = m.renormalize_cj1(src['C'][(3 + k) % 4], src['E'][(2 + k) % 4], p1i)
        c2 = m.renormalize_cj2(src['C'][k], src['E'][k], p2j)
        ei = src['E'][(3 + k) % 4]
        a0, a1 = _a_block(ei.shape[0], G, g)
        from . import ops
        if a1 > a0:
            e = ops.absorb_edge(ei[a0:a1].contiguous(), src.bond_permute(k), p2i[a0:a1].contiguous(), p1j, normalize=False)
        else:
            e = torch.zeros(p2i.shape[3], p1j.shape[3], ei.shape[2], ei.shape[3], dtype=ei.dtype, device=ei.device)
        dist.all_reduce(e, op=dist.ReduceOp.SUM, group=group)
        ops.frob_normalize(e)
        return c1, c2, e

    def absorb(self, ipeps, task, p1i, p2i, p1j, p2j):
        """The three absorptions of one task (directional_mover.py:293-303) -> (C[(3+k)%4], C[k], E[(3+k)%4]) of site s2."""
        k = task["k"]
        src = ipeps[task["s1"]]
        m = self.mover
        c1 = m.renormalize_cj1(src['C'][(3 + k) % 4], src['E'][(2 + k) % 4], p1i)
        c2 = m.renormalize_cj2(src['C'][k], src['E'][k], p2j)
        e = m.renormalize_ej(src['E'][(3 + k) % 4], src.bond_permute(k), p2i, p1j)
        return c1, c2, e


def _a_block(n, G, g):
    base, rem = divmod(n, G)
    start = g * base + min(g, rem)
    return start, start + base + (1 if g < rem else 0)


def phase_moves(ipeps):
    """The phases of one sweep in the reference's order (ctmrg.py:20-24)."""
    out = []
    for xi in range(ipeps.nx):
        out.append([(0, xi), (2, (ipeps.nx - xi + 1) % ipeps.nx)])
    for yi in range(ipeps.ny):
        out.append([(1, (ipeps.ny - yi + 1) % ipeps.ny), (3, yi)])
    return out


class ShardedCtmrg:
    """group_size G > 1: ranks [iG, (i+1)G) form a group that computes each of its projector tasks cooperatively
    (row-sharded, see sharded_projector.py); tasks are dealt to the world/G groups, absorptions to single ranks."""

    def __init__(self, ipeps, config, rank, world, compute=None, group=None, group_size=1):
        if getattr(config, "svd_type", "rsvd") != "rsvd":
            raise ValueError("site-sharded CTMRG supports svd_type='rsvd' only")
        if config.projectors != "half-system":
            raise ValueError("site-sharded CTMRG needs half-system projectors (full-system moves of a pair are not independent, "
                             "SURVEY.md App. D4)")
        self.ipeps, self.config, self.rank, self.world, self.group = ipeps, config, rank, world, group
        self.compute = compute if compute is not None else B200Compute(config)
        self.bytes_exchanged = 0
        if group_size < 1 or world % group_size != 0:
            raise ValueError(f"group_size {group_size} must divide the world size {world}")
        self.G = group_size
        self.ngroups = world // group_size
        self.gi, self.g = rank // group_size, rank % group_size
        self.pair_group = None
        if group_size > 1:
            for i in range(self.ngroups):         # every rank creates every group (torch.distributed requirement)
                pg = dist.new_group(ranks=list(range(i * group_size, (i + 1) * group_size)))
                if i == self.gi:
                    self.pair_group = pg

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def _bcast(self, tensor, src):
        if self.world > 1:
            dist.broadcast(tensor, src=src, group=self.group)
            self.bytes_exchanged += tensor.numel() * tensor.element_size()
        return tensor

    def owner_group(self, n):
        return n % self.ngroups

    def owner(self, n):
        """Rank that broadcasts the projector pair of task n and runs its absorptions."""
        return self.owner_group(n) * self.G + (n // self.ngroups) % self.G

    # ---- one phase ---------------------------------------------------------------------------------------------------
    def phase(self, moves):
        ip, cp, rank = self.ipeps, self.compute, self.rank
        tasks = []
        for k, line in moves:
            tasks += cp.tasks(ip, k, line)
        omegas = [cp.draw_omega(ip, t) for t in tasks]                    # every rank replays every draw
        coop = [n for n in range(len(tasks)) if self.owner_group(n) == self.gi]      # tasks my group computes
        if self.G > 1:
            # the ranks of a group multiply their row blocks by the SAME test matrix: do not rely on equal generator states, take the
            # group leader's draw (34 MB per task over NVLink)
            for n in coop:
                omegas[n] = omegas[n].contiguous()
                dist.broadcast(omegas[n], src=self.gi * self.G, group=self.pair_group)
            coop_pairs = cp.projectors(ip, [tasks[n] for n in coop], [omegas[n] for n in coop], group=self.pair_group,
                                       group_rank=self.g, group_size=self.G)
        else:
            coop_pairs = cp.projectors(ip, [tasks[n] for n in coop], [omegas[n] for n in coop])
        mine = [n for n in coop if self.owner(n) == rank]
        pairs = [pr for n, pr in zip(coop, coop_pairs) if self.owner(n) == rank]
        A0 = ip[tasks[0]["s1"]]['A']
        device, dtype = A0.device, A0.dtype
        D = ip.dims["bond"]
        # truncated ranks of all tasks
        keep = torch.zeros(len(tasks), dtype=torch.int64, device=device)
        for n, (p1, _) in zip(mine, pairs):
            keep[n] = p1.shape[-1]
        if self.world > 1:
            dist.all_reduce(keep, op=dist.ReduceOp.SUM, group=self.group)
        keep = [int(v) for v in keep.tolist()]
        # projector exchange
        P1, P2 = {}, {}
        own = dict(zip(mine, pairs))
        for n, t in enumerate(tasks):
            if n in own:
                p1, p2 = own[n]
            else:
                chi1 = ip[t["plaq"][0]]['E'][(3 + t["k"]) % 4].shape[0]       # leg e of Q1  (projectors.py:52-59)
                chi2 = ip[t["plaq"][3]]['E'][(t["k"] + 3) % 4].shape[1]       # leg c of Q4
                p1 = torch.empty(chi1, D, D, keep[n], dtype=dtype, device=device)
                p2 = torch.empty(chi2, D, D, keep[n], dtype=dtype, device=device)
            P1[(t["k"], t["key"])] = self._bcast(p1, self.owner(n))
            P2[(t["k"], t["key"])] = self._bcast(p2, self.owner(n))
        # owners absorb (G > 1: every rank of the owning group takes part in each of the group's absorptions)
        results = {}
        if self.G > 1 and hasattr(cp, "absorb_coop"):
            for n in coop:
                t = tasks[n]
                k = t["k"]
                res = cp.absorb_coop(ip, t, P1[(k, t["i"])], P2[(k, t["i"])], P1[(k, t["j"])], P2[(k, t["j"])], self.pair_group,
                                     self.g, self.G)
                if self.owner(n) == rank:
                    results[n] = res
        else:
            for n in mine:
                t = tasks[n]
                k = t["k"]
                results[n] = cp.absorb(ip, t, P1[(k, t["i"])], P2[(k, t["i"])], P1[(k, t["j"])], P2[(k, t["j"])])
        # publish the new boundary tensors (all reads of this phase are done: the writes touch tensors no task reads)
        for n, t in enumerate(tasks):
            k = t["k"]
            if n in results:
                c1, c2, e = (x.contiguous() for x in results[n])
            else:
                xa = ip[t["s1"]]['E'][(2 + k) % 4].shape[0]
                xc = ip[t["s1"]]['E'][k].shape[1]
                ki, kj = keep[self._index(tasks, k, t["i"])], keep[self._index(tasks, k, t["j"])]
                c1 = torch.empty(xa, ki, dtype=dtype, device=device)
                c2 = torch.empty(kj, xc, dtype=dtype, device=device)
                e = torch.empty(ki, kj, D, D, dtype=dtype, device=device)
            src = self.owner(n)
            dst = ip[t["s2"]]
            dst['C'][(3 + k) % 4] = self._bcast(c1, src)
            dst['C'][k] = self._bcast(c2, src)
            dst['E'][(3 + k) % 4] = self._bcast(e, src)

    @staticmethod
    def _index(tasks, k, key):
        for n, t in enumerate(tasks):
            if t["k"] == k and t["key"] == key:
                return n
        raise KeyError((k, key))

    def sweep(self):
        for moves in phase_moves(self.ipeps):
            self.phase(moves)

    def run(self, steps=None):
        for _ in range(self.config.steps if steps is None else steps):
            self.sweep()
