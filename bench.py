#!/usr/bin/env python
"""bench.py -- CTMRG sweeps/s at D=8, chi=256 (FP64) on 1/2/4/8 B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--D 8 --chi 256 --d 2 --nx 2 --ny 2]

One "step" = one CTMRG sweep (acetn/renormalization/ctmrg.py:18-31 = 4*nx*ny site-moves: projector pair + three
absorptions each) over synthetic random-init iPEPS tensors.  `value` is reported in sweeps of the
2x2 reference cell (16 site-moves) per second so that runs on different unit cells are comparable.
Prints ONE JSON line (rank 0).  --impl reference times the CPU restatement of the reference path (oracle/), which is
the reference's own torch code path on the host cores (the reference is a Python package and cannot travel to the box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "CTMRG sweeps/sec at D=8 chi=256 (FP64), 16 site-moves per sweep"
UNIT = "sweeps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--D", type=int, default=8)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--d", type=int, default=2)
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--group-size", type=int, default=0, help="ranks cooperating on one projector (0 = auto)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' times the reference torch path (cuBLAS/cuSOLVER) on the GPU instead of the host cores")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_site_move_seconds(args, reps=1):
    """One site-move (left projector pair + renormalize_boundary) of the reference path on the host cores."""
    import torch
    from oracle import ctmrg_oracle as orc
    torch.manual_seed(args.seed)
    cell = orc.random_cell(2, 2, args.D, args.chi, args.d, seed=args.seed)
    cfg = orc.CtmrgConfig()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        p1, p2 = orc.half_system_projectors(cell, orc.plaquette(cell, 0, 0, 0), 0, cfg)
        orc.renormalize_boundary(cell, {0: p1, 1: p1}, {0: p2, 1: p2}, (0, 0), (1, 0), 0, 1, 0)
        best = min(best, time.perf_counter() - t0)
    return best


def gpu_torch_site_move_seconds(args, dev):
    """The reference's torch path (oracle port: torch.einsum / @ / linalg.qr / linalg.svd -> cuBLAS + cuSOLVER) on the SAME
    GPU: the comparator BASELINE.md asks for besides the CPU baseline.  One site-move, best of 2 after one warm-up."""
    import torch
    from oracle import ctmrg_oracle as orc
    cell = orc.random_cell(2, 2, args.D, args.chi, args.d, seed=args.seed)
    for s in cell.site_list:
        st = cell[s]
        st.A = st.A.to(dev)
        st.C = [c.to(dev) for c in st.C]
        st.E = [e.to(dev) for e in st.E]
    cfg = orc.CtmrgConfig()
    omega_fn = lambda n, q, dtype=torch.float64, device=dev: torch.randn(n, q, dtype=dtype, device=dev)   # noqa: E731
    best = 1e30
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p1, p2 = orc.half_system_projectors(cell, orc.plaquette(cell, 0, 0, 0), 0, cfg, omega_fn=omega_fn)
        work = cell.clone()
        orc.renormalize_boundary(work, {0: p1, 1: p1}, {0: p2, 1: p2}, (0, 0), (1, 0), 0, 1, 0)
        torch.cuda.synchronize()
        if it > 0:
            best = min(best, time.perf_counter() - t0)
        del p1, p2, work
    torch.cuda.empty_cache()
    return best


def run_reference(args):
    """Reference arm: the reference's CPU torch path (oracle port), one bounded sample (= one site-move) per step."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is the CPU path on ALL host cores of the box
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    times = []
    on_gpu = args.ref_device == "cuda"
    for i in range(args.warmup + args.steps):
        t = gpu_torch_site_move_seconds(args, torch.device("cuda", 0)) if on_gpu else cpu_site_move_seconds(args)
        if i >= args.warmup:
            times.append(t)
    t_move = sum(times) / len(times)
    value = 1.0 / (16.0 * t_move)
    cores = torch.get_num_threads()
    sample = f"one site-move (half-system projector pair + renormalize_boundary) at D={args.D} chi={args.chi} per step; sweep = 16 site-moves"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 16.0 * t_move * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"CTMRG sweep, D={args.D} chi={args.chi} d={args.d}, {'2x2' if args.gpus <= 1 else '4x4'} cell "
                                   f"({16 if args.gpus <= 1 else 64} site-moves/sweep), half-system rsvd niter=2 p=2",
                       "value_unit": "sweeps of 16 site-moves per second",
                       "device": "cuda (reference torch path = cuBLAS/cuSOLVER via torch, oracle port)" if on_gpu else "cpu (reference torch path, oracle port)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist

    from acetn_b200 import _lib, ops
    from acetn_b200.ipeps import CTMRGConfig, SiteTensor
    from acetn_b200.renormalization import DirectionalMover, ctmrg
    from acetn_b200.synthetic import flops_sweep, random_ipeps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the b200 backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    real_stdout = None
    if world > 1:
        # exactly ONE line on stdout: native libraries (NCCL's version banner, NCCL_DEBUG output) write to fd 1 directly, so fd 1
        # is pointed at stderr for the duration of the run and the JSON line goes to the saved descriptor at the end
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    n = max(world, 1)
    gsz = 1
    nx = args.nx or (2 if n == 1 else 4)
    ny = args.ny or (2 if n == 1 else 4)
    D, chi, d = args.D, args.chi, args.d

    cfg = CTMRGConfig(steps=1)
    ip = random_ipeps(nx, ny, D, chi, d, seed=args.seed, ctmrg=cfg, device=dev)
    mover = DirectionalMover(cfg)
    torch.manual_seed(args.seed + 1)      # Omega stream (device generator), identical on every rank

    if world > 1:
        from acetn_b200.distributed import ShardedCtmrg
        # ranks per projector: 1 while there are at least as many site tasks per phase (2*min(nx,ny)) as ranks; beyond
        # that, pairs of ranks compute one projector cooperatively (row-sharded, acetn_b200/sharded_projector.py)
        gsz = args.group_size
        if not gsz:
            gsz = 1
            while world // gsz > 2 * min(nx, ny) and world % (2 * gsz) == 0:
                gsz *= 2
        sharded = ShardedCtmrg(ip, cfg, rank, world, group_size=gsz)
        sweep = sharded.sweep
    else:
        def sweep():
            ctmrg(ip, cfg, mover)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sweep()
    barrier()
    for s in ip.site_list:
        for k in range(4):
            assert tuple(ip[s]['C'][k].shape) == (chi, chi), f"chi not saturated after warm-up: {tuple(ip[s]['C'][k].shape)}"

    # ---- timed region: K sweeps, inputs resident in HBM ------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        sweep()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count()
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    site_moves = 4 * nx * ny
    value = args.steps * (site_moves / 16.0) / (ms * 1e-3)

    # ---- e2e: the same sweeps with the state living in pinned HOST buffers (H2D + sweep + D2H every step) -----------
    host = {s: {"A": ip[s]['A'].cpu().pin_memory(), "C": [c.cpu().pin_memory() for c in ip[s]['C']],
                "E": [e.cpu().pin_memory() for e in ip[s]['E']]} for s in ip.site_list}
    h2d = sum(h["A"].numel() + sum(c.numel() for c in h["C"]) + sum(e.numel() for e in h["E"]) for h in host.values()) * 8
    d2h = sum(sum(c.numel() for c in h["C"]) + sum(e.numel() for e in h["E"]) for h in host.values()) * 8
    # N > 1: the state is replicated on every rank (like the reference's distributed mode), but it crosses PCIe only once:
    # site s is uploaded / downloaded by rank (index of s) % world, and replicated to the other ranks over NVLink (NCCL
    # broadcast).  h2d / d2h count the bytes of the whole job per step.
    site_owner = {s: i % world for i, s in enumerate(ip.site_list)}

    def e2e_step():
        for s in ip.site_list:
            h = host[s]
            if site_owner[s] == rank:
                st = SiteTensor(h["A"].to(dev, non_blocking=True), [c.to(dev, non_blocking=True) for c in h["C"]],
                                [e.to(dev, non_blocking=True) for e in h["E"]])
            else:
                mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev)     # noqa: E731
                st = SiteTensor(mk(h["A"]), [mk(c) for c in h["C"]], [mk(e) for e in h["E"]])
            if world > 1:
                for t in [st['A']] + list(st['C']) + list(st['E']):
                    dist.broadcast(t, src=site_owner[s])
            ip[s] = st
        sweep()
        for s in ip.site_list:
            if site_owner[s] != rank:
                continue
            for k in range(4):
                host[s]["C"][k].copy_(ip[s]['C'][k], non_blocking=True)
                host[s]["E"][k].copy_(ip[s]['E'][k], non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = args.steps * (site_moves / 16.0) / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        # ---- rooflines, measured live (CUDA events on the launching stream; operands of 2 GiB+ exceed L2) --------------
        import ctypes
        m = chi * D * D
        q = chi + 2
        nrep = 20
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(nrep):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / nrep * 1e-3

        scratch = torch.empty(1 << 20, dtype=torch.float64, device=dev)
        lib.acetn_b200_fp64_peak_probe(ctypes.c_void_p(scratch.data_ptr()), 2000, stream)
        torch.cuda.synchronize()
        e0.record()
        fl = lib.acetn_b200_fp64_peak_probe(ctypes.c_void_p(scratch.data_ptr()), 20000, stream)
        e1.record()
        torch.cuda.synchronize()
        peak = fl / (e0.elapsed_time(e1) * 1e-3) * 1e-12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))
            except Exception:
                traffic = {}
        step_s = ms * 1e-3 / args.steps
        st = ip[(0, 0)]
        Q, _ = ops.quarter_tensor(st['C'][0], st['E'][0], st['E'][3], st['A'])
        X = torch.randn(m, q, dtype=torch.float64, device=dev)
        use_i8 = mover.projector_calculator._use_i8([(m, m), (m, m)], q)
        # K1 (FP64 DMMA GEMM) on the contraction class that dominates its time: (chi D^2) x (chi D^2) x chi, the 2 chi^3 D^4
        # steps of make_quarter_tensor (projectors.py:53) and renormalize_ej (directional_mover.py:362,365): 4 per site-move
        Am = torch.randn(m, chi, dtype=torch.float64, device=dev)
        Bm = torch.randn(chi, m, dtype=torch.float64, device=dev)
        t_med = timed(lambda: ops.matmul(Am, Bm, out=Q))
        med_flops = 2.0 * m * m * chi
        k1 = {"bound": "tensor", "achieved": med_flops / t_med * 1e-12, "peak": peak, "unit": "TFLOP/s",
              "frac": med_flops / t_med * 1e-12 / peak, "traffic": traffic.get("k1_medium_dram_bytes_per_launch"),
              "kernel": "dgemm_dmma_kernel (K1), %dx%dx%d (2 chi^3 D^4 contraction class)" % (m, m, chi),
              "peak_source": "live DMMA.8x8x4 issue-rate probe in this run (FP64 tensor pipe; MEASURED_PEAKS.json has no FP64 entry)",
              "launches_per_step": 4 * site_moves // n, "share_of_step": (4 * site_moves / n) * t_med / step_s}
        Q, _ = ops.quarter_tensor(st['C'][0], st['E'][0], st['E'][3], st['A'])
        del Am, Bm
        # the thin product (chi D^2)^2 x (chi+2): 13 per site-move.  K1 (DMMA) and, when selected, K7 (INT8 tensor cores, exact)
        t_thin_dmma = timed(lambda: ops.matmul(Q, X))
        thin_flops = 2.0 * m * m * q
        thin = {"dmma_ms": t_thin_dmma * 1e3, "dmma_tflops": thin_flops / t_thin_dmma * 1e-12, "dmma_frac_of_fp64_peak": thin_flops / t_thin_dmma * 1e-12 / peak}
        roofline = k1
        k7 = None
        if use_i8:
            enc = ops.i8_encode(Q)
            t_enc = timed(lambda: ops.i8_encode(Q, storage=enc.storage))
            out = torch.empty(m, q, dtype=torch.float64, device=dev)
            t_i8 = timed(lambda: ops.i8_matmul(enc, X, out=out))
            # algorithmic HBM bytes of one K7 product: 16 residue planes of Q (int8) + thin operand in (FP64) + result out (FP64)
            i8_bytes = 16.0 * m * m + 2.0 * 8.0 * m * q
            k7 = {"bound": "hbm", "achieved": i8_bytes / t_i8 * 1e-9, "peak": hbm_peak, "unit": "GB/s", "frac": i8_bytes / t_i8 * 1e-9 / hbm_peak,
                  "traffic": traffic.get("k7_product_dram_bytes"), "peak_source": hbm_src,
                  "kernel": "K7 thin product %dx%dx%d = thin_encode + i8_gemm_kernel (tcgen05.mma kind::i8) + crt_kernel" % (m, q, m),
                  "ms": t_i8 * 1e3, "fp64_equivalent_tflops": thin_flops / t_i8 * 1e-12, "x_fp64_dmma_peak": thin_flops / t_i8 * 1e-12 / peak,
                  "int8_tops": 16.0 * 2.0 * m * m * (((q + 15) // 16) * 16) / t_i8 * 1e-12,
                  "encode_ms_per_quarter_tensor": t_enc * 1e3, "encode_gbs": (8.0 + 16.0) * m * m / t_enc * 1e-9,
                  "launches_per_step": 13 * site_moves // n, "share_of_step": ((13 * t_i8 + 2 * t_enc) * site_moves / n) / step_s}
            thin["speedup_k7_over_dmma"] = t_thin_dmma / t_i8
            del enc, out
        roofline["whole_sweep_tflops"] = flops_sweep(nx, ny, D, chi, d) * args.steps / (ms * 1e-3) * 1e-12 / n
        roofline["whole_sweep_frac_of_fp64_peak"] = roofline["whole_sweep_tflops"] / peak
        roofline["thin_product"] = thin
        del Q, X
        ops.release_workspace()
        torch.cuda.empty_cache()
        cpu = None
        if n == 1 and not args.no_cpu_baseline:
            t_move = cpu_site_move_seconds(args)
            cpu = {"value": 1.0 / (16.0 * t_move), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"one site-move (projector pair + renormalize_boundary) at D={D} chi={chi}, {t_move:.2f} s, scaled x16 to a 2x2 sweep"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"CTMRG sweep, D={D} chi={chi} d={d}, {nx}x{ny} cell ({site_moves} site-moves/sweep), half-system rsvd niter=2 p=2",
                           "cell": f"{nx}x{ny}", "value_unit": "sweeps of 16 site-moves per second", "parallelism": f"site-sharded x{n}" + (f", {gsz} ranks per projector (row-sharded)" if n > 1 and gsz > 1 else ""),
                           "l2": "inputs (2 GiB quarter tensors) exceed the 126 MB L2; no flush between iterations"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roofline}
        line["config"]["thin_engine"] = "i8 (K7: exact integer products on the INT8 tensor cores)" if use_i8 else "dmma (K1)"
        if k7 is not None:
            line["roofline_k7"] = k7
        if cpu is not None:
            line["cpu_baseline"] = cpu

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
