#!/usr/bin/env python
"""Headline benchmark: grid-cells x levels per second of the FV transport step on C384x72.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the configuration the headline target is quoted on):
FV3-style horizontal finite-volume flux/advection stencil (fv_tp2d, 3-cell halo) on the C384 cubed
sphere (6 tiles x 384 x 384 columns) x 72 levels, synthetic fields.  One STEP = halo update of q from
the neighbouring sub-domains (library-owned exchange: same-GPU copies, and at N > 1 a neighbour handshake + pull over NVLink
peer memory, overlapped with the cells of the stencil that read no halo) followed by fv_tp2d on every sub-domain the
GPU hosts.  The domain is fixed, so scaling is STRONG.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "geosongpu-ci_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

CUBE_N, NK, HALO = 384, 72, 3
METRIC = "grid-cells x levels per second, fv_tp2d transport step (C384x72, 6 tiles)"
UNIT = "points/s"


def AUTO_STEP_MODE(n_gpus: int) -> str:
    """How --step auto launches the device-exchange step: exchange kernel, then the plain stencil, at every N.

    Measured in round 2 (profiles/README.md): on one GPU every link is a same-GPU copy that competes with the stencil for
    HBM, so forking it only slows both (545 against 533 us); on two GPUs the gated stencil starts later than the serial one
    (307 against 292 us: the forked exchange gets two blocks per SM and sub-domain 0's third of it takes longer than the
    whole exchange alone); on eight the two tie (112.6 against 112.7 us)."""
    return "serial"


def algorithmic_bytes_per_point(es: int) -> float:
    """40 R (q, crx, xfx, cry, yfx) + 8 W + 8/nk (rarea) in fp64; scales with the element size."""
    return 6 * es + es / NK


# ---------------------------------------------------------------------------------------------------
# CPU side: the oracle port, timed on the host cores (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------


def host_cores() -> int:
    """Cores this process may run on (the affinity mask, not os.cpu_count(): a cgroup / taskset can be narrower)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_inputs(ni, nj, nk, dtype, seed=20240728):
    import numpy as np

    from oracle import inputs as gen

    rng = np.random.default_rng(seed)

    def rnd(shape, lo, hi):
        a = gen.ifirst_empty(shape, dtype)
        a[...] = rng.uniform(lo, hi, size=shape)
        return a

    q = rnd((ni + 6, nj + 6, nk), 0.5, 1.5)
    crx, cry = rnd((ni + 1, nj, nk), -0.9, 0.9), rnd((ni, nj + 1, nk), -0.9, 0.9)
    xfx, yfx = crx * 1.05, cry * 0.95
    xfx, yfx = gen.as_ifirst(xfx), gen.as_ifirst(yfx)
    rarea = rnd((ni, nj), 0.9, 1.1)
    out = gen.ifirst_empty((ni, nj, nk), dtype)
    out[...] = 0
    return q, crx, xfx, cry, yfx, rarea, out


class CpuStep:
    """The GPU arm's step on the host cores: halo fill of q on all six tiles (oracle/c halo_move over the same link
    table) followed by oracle/c fv_tp2d on each of the six 384 x 384 x 72 tiles -- the whole C384x72 workload."""

    def __init__(self, dtype_name: str):
        import numpy as np
        import torch

        from b200stencil.halo.partitioner import CubedSpherePartitioner
        from b200stencil.halo.updater import FieldGeometry, HaloPlan
        from oracle.c_oracle import COracle

        self.dtype = np.float64 if dtype_name == "f64" else np.float32
        try:
            self.orc = COracle(native=True)  # -O3 -march=native, built on this machine
            self.build = "gcc -O3 -march=native -fopenmp"
        except Exception:
            self.orc = COracle(native=False)
            self.build = "gcc -O2 -fopenmp (portable build)"
        # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports 1)
        self.cores = host_cores()
        self.orc.set_threads(self.cores)
        self.threads = self.orc.threads
        if self.cores > 1 and self.threads <= 1:
            raise RuntimeError(f"CPU baseline would run on 1 of {self.cores} cores: refusing to time it")
        n = CUBE_N
        self.tiles = [cpu_sample_inputs(n, n, NK, self.dtype, seed=20240728 + t) for t in range(1)]
        # six tiles: q is one batch storage (the halo fill crosses tiles); the other inputs are tile 0's, reused --
        # their values do not change the work and generating 2.5 GB of random numbers would dominate the leg
        self.q = np.empty((6, NK, n + 6, n + 6), dtype=self.dtype)
        for t in range(6):
            self.q[t] = self.tiles[0][0].transpose(2, 1, 0)
        self.q_view = [self.q[t].transpose(2, 1, 0) for t in range(6)]  # [i, j, k], i-fastest
        self.outs = [self.tiles[0][6]] + [np.zeros_like(self.tiles[0][6]) for _ in range(1)]
        part = CubedSpherePartitioner(n, (1, 1), HALO)
        tq = torch.from_numpy(self.q).permute(0, 3, 2, 1)
        self.links = HaloPlan(part, 1, 0).tables(FieldGeometry(tq, HALO))["local"]
        self.points = 6 * n * n * NK

    def __call__(self):
        flat = self.q.reshape(-1)
        self.orc.halo_move(self.links, NK, flat, flat)
        _, crx, xfx, cry, yfx, rarea, _ = self.tiles[0]
        for t in range(6):
            self.orc.fv_tp2d(self.q_view[t], crx, xfx, cry, yfx, rarea, self.outs[t & 1])


def time_cpu_port(dtype_name: str, budget_s: float = 20.0, steps=None, warmup: int = 1):
    """Time the C/OpenMP restatement (oracle/c) on the WHOLE workload step (halo fill + six tiles), all host threads.

    Returns (points_per_second, description dict, per-step times)."""
    step = CpuStep(dtype_name)
    for _ in range(warmup):
        step()
    times = []
    t_begin = time.perf_counter()
    n = 0
    while True:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        n += 1
        if steps is not None:
            if n >= steps:
                break
        elif n >= 3 and time.perf_counter() - t_begin > budget_s or n >= 50:
            break
    best = min(times)
    med = statistics.median(times)
    info = {
        "value": step.points / med, "unit": UNIT, "cores": step.threads, "kind": "port",
        "sample": f"oracle/c halo_move + fv_tp2d ({step.build}, {step.threads} OpenMP threads of {step.cores} usable cores) on the whole "
                  f"step: halo fill + 6 tiles of 384x384x{NK} ({dtype_name}), {n} steps, median {med * 1e3:.1f} ms, best {best * 1e3:.1f} ms",
        "ms_per_sample": med * 1e3,
    }  # fmt: skip
    return step.points / med, info, times


def time_numpy_port(dtype_name: str):
    """The NumPy restatement (stand-in for gt4py's numpy backend), single thread, on a 96x96x72 sample."""
    import numpy as np

    from oracle import numpy_oracle as orc

    dtype = np.float64 if dtype_name == "f64" else np.float32
    n = 96
    args = cpu_sample_inputs(n, n, NK, dtype)
    orc.fv_tp2d(*args)
    t0 = time.perf_counter()
    orc.fv_tp2d(*args)
    dt = time.perf_counter() - t0
    return {"value": n * n * NK / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle/numpy_oracle.fv_tp2d on 96x96x{NK} ({dtype_name}), 1 run {dt * 1e3:.0f} ms"}  # fmt: skip


def run_reference(ns) -> int:
    """--impl reference: the reference's CPU implementation of the path.  gt4py/NDSL cannot be installed
    here, so this is the oracle port (C/OpenMP, every host thread) on the same config/metric:
    one step = halo fill + fv_tp2d on all six tiles, like the GPU arm's."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, info, times = time_cpu_port(ns.dtype, steps=max(1, min(ns.steps, 30)), warmup=max(1, min(ns.warmup, 3)))
    ms = statistics.median(times) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ns.gpus, "steps": len(times),
        "warmup": max(1, min(ns.warmup, 3)), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": ns.dtype, "data": "synthetic",
        "config": {"workload": "fv_tp2d C384x72 (CPU port): halo fill of q + PPM flux-form update on all 6 tiles, the GPU arm's step",
                   "grid": f"C{CUBE_N}", "tiles": 6, "levels": NK, "halo": HALO, "same_config": True,
                   "steps_note": "at most 30 timed steps (bounded CPU leg)"},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--dtype", choices=["f64", "f32"], default="f64")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=["fv", "chain", "patterns"], default="fv",
                    help="fv (default, the contract line): fv_tp2d transport step on C384x72; chain: BASELINE configs[4], "
                         "fv_tp2d + pe_prefix + remap on C720x137; patterns: BASELINE configs[1], the dsl_patterns column "
                         "stencils on C96x72 (separate reports, see run_chain / run_patterns)")
    ap.add_argument("--step", choices=["auto", "fused", "overlap", "serial"], default="auto",
                    help="how the device-exchange step is launched: fused = ONE kernel (b2s_halo_fv_tp2d: the stencil grid shares the "
                         "exchange among its CTAs first, then computes behind per-sub-domain gates); overlap = exchange kernel forked "
                         "beside one gated stencil launch; serial = exchange kernel, then stencil; auto = serial (AUTO_STEP_MODE)")
    ap.add_argument("--no-overlap", action="store_true", help="same as --step serial (and no interior/frame overlap on the NCCL baseline)")
    ap.add_argument("--halo", choices=["auto", "device", "nccl"], default="auto",
                    help="halo exchange: device (= auto) the library-owned exchange (neighbour handshake "
                         "+ pull over NVLink peer memory, b2s_halo_*), forked next to a gated stencil launch; nccl = the "
                         "portable baseline: packed strips + grouped NCCL send/recv overlapped with an interior launch")
    ap.add_argument("--option", action="append", default=[], help="libb200stencil tuning option name=value (b2s_set_option), repeatable")
    ap.add_argument("--push", choices=["staged", "inplace", "off"], default="off",
                    help="[serial step, N > 1] strips that cross NVLink: off (default) = every strip pulled in place; staged = pushed "
                         "PACKED by their owner into a staging area behind the destination's field and unpacked there; inplace = "
                         "pushed straight into the halo cells (both measured slower than the pull: profiles/README.md, round 2)")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly (default: the step is replayed from a CUDA graph at every N)")
    ap.add_argument("--regions", type=int, default=0, help="timed regions of K steps each (median reported); 0 = auto")
    ap.add_argument("--fused-remap", action="store_true", help="[chain] fold pe_prefix into the remap kernel (remap_delp)")
    ap.add_argument("--hws-dump", default=None, help="write the hws sampler record of the run (npz) to this path")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true")
    ns = ap.parse_args(argv)
    if ns.impl == "reference":
        return run_reference(ns)
    if ns.workload == "chain":
        return run_chain(ns)
    if ns.workload == "patterns":
        return run_patterns(ns)

    import torch
    import torch.distributed as dist

    from b200stencil import _abi, fields, hostio, stencils
    from b200stencil.bench import harness
    from b200stencil.halo.device import HaloContext
    from b200stencil.halo.partitioner import CubedSpherePartitioner, expected_halo, global_id_field, layout_for
    from b200stencil.halo.transport import FvTransport
    from b200stencil.hws import Sampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = ns.gpus
    if world != n_gpus:
        if world == 1 and n_gpus > 1:
            raise SystemExit(f"--gpus {n_gpus} needs torchrun with {n_gpus} ranks (one process per GPU)")
        n_gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = hostio.bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # torch.distributed: timing plumbing (barrier, max over ranks), the session-name bootstrap and the NCCL
        # baseline exchange; the product exchange is libb200stencil's
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if ns.dtype == "f64" else torch.float32
    es = 8 if ns.dtype == "f64" else 4

    part = CubedSpherePartitioner(CUBE_N, layout_for(n_gpus), HALO)
    nsub = part.subdomains_per_gpu(n_gpus)
    ni, nj = part.nx, part.ny

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic fields, resident in HBM before the timed region (SURVEY.md 8d recipe, device RNG)
    g = torch.Generator(device=dev)
    g.manual_seed(20240724 + 4 + 1000 * rank)
    mk = lambda s, lo, hi: fields.empty(s, dtype, dev, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    use_device = ns.halo in ("auto", "device")
    ctx = HaloContext(rank, world, local_rank)
    for opt in ns.option:
        name, val = opt.split("=")
        _abi.set_option(name, int(val))
    step_mode = "serial" if ns.no_overlap else (AUTO_STEP_MODE(n_gpus) if ns.step == "auto" else ns.step)
    stage_part = part if ns.push == "staged" else None  # reserve the staging area behind the exchanged fields

    # ---- halo_check: the exchange the timed loop uses, on a global-id field, every halo cell against geometry ----
    halo_check = None
    if use_device:
        nk_chk = 2
        idf = ctx.field((ni + 6, nj + 6, nk_chk), nsub, torch.float64, part=stage_part)
        for b in range(nsub):
            idf[b].copy_(torch.from_numpy(global_id_field(part, rank * nsub + b, nk_chk)))
        barrier()
        ex_chk = ctx.plan(idf, part, push=ns.push != "off")
        for _ in range(2):
            ex_chk.update()
        torch.cuda.synchronize()
        bad = 0
        for b in range(nsub):
            want = torch.from_numpy(expected_halo(part, rank * nsub + b, nk_chk)).to(dev)
            bad += int((idf[b] != want).sum().item())
        epoch, status = ctx.status()
        flag = torch.tensor([bad + status], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(flag)
        if int(flag.item()) != 0:
            raise SystemExit(f"[bench] halo_check FAILED on rank {rank}: {bad} wrong halo cells, device status {status}")
        halo_check = "ok"
        barrier()

    if use_device:
        q = ctx.field((ni + 6, nj + 6, NK), nsub, dtype, part=stage_part)
        q.uniform_(0.5, 1.5, generator=g)
        ex = ctx.plan(q, part, push=ns.push != "off")
        tr = FvTransport(part, n_gpus, rank, exchange="device", halo_exchange=ex, overlap=step_mode != "serial", fused=step_mode == "fused")
    else:
        q = mk((ni + 6, nj + 6, NK), 0.5, 1.5)
        ex = None
        tr = FvTransport(part, n_gpus, rank, overlap=not ns.no_overlap, exchange="nccl")
    crx, cry = mk((ni + 1, nj, NK), -0.9, 0.9), mk((ni, nj + 1, NK), -0.9, 0.9)
    xfx = mk((ni + 1, nj, NK), 0.9, 1.1).mul_(crx)
    yfx = mk((ni, nj + 1, NK), 0.9, 1.1).mul_(cry)
    rarea = mk((ni, nj), 0.9, 1.1)
    q_out = fields.empty((ni, nj, NK), dtype, dev, batch=nsub)
    args = (q, crx, xfx, cry, yfx, rarea, q_out)
    total_points = 6 * CUBE_N * CUBE_N * NK
    local_points = nsub * ni * nj * NK
    input_mb = sum(t.numel() * es for t in args[:6]) / 1e6
    barrier()

    # ---- the product step against the NCCL baseline step (same fields): bit-identical, or the run stops ----
    nccl_equal = None
    if use_device and world > 1:
        tr.step(*args)
        q_b = fields.empty((ni + 6, nj + 6, NK), dtype, dev, batch=nsub)
        q_b.copy_(q)
        q_b[:, :3, 3:-3] = -1.0
        q_b[:, -3:, 3:-3] = -1.0  # scrub west / east halos: the NCCL exchange has to refill them
        out_b = fields.empty((ni, nj, NK), dtype, dev, batch=nsub)
        FvTransport(part, n_gpus, rank, overlap=False, exchange="nccl").step(q_b, crx, xfx, cry, yfx, rarea, out_b)
        torch.cuda.synchronize()
        same = torch.equal(out_b, q_out) and torch.equal(q_b[:, :, 3:-3], q[:, :, 3:-3]) and torch.equal(q_b[:, 3:-3], q[:, 3:-3])
        flag = torch.tensor([0 if same else 1], device=dev, dtype=torch.int64)
        dist.all_reduce(flag)
        if int(flag.item()) != 0:
            raise SystemExit(f"[bench] rank {rank}: the device-exchange step differs from the NCCL baseline step")
        nccl_equal = True
        del q_b, out_b
        barrier()

    # ---- warm-up ----
    for _ in range(max(3, ns.warmup)):
        tr.step(*args)
    barrier()
    if ex is not None:
        ctx.check()

    full_call = tr.calls(*args)[0]
    launches_before = _abi.launch_count()
    tr.step(*args)
    launches_eager_step = _abi.launch_count() - launches_before  # kernels of libb200stencil per step (NCCL's own are not counted)
    barrier()

    # The whole step (both streams, exchange included) is captured ONCE into a CUDA graph and replayed at every N
    # (N = 1 included: one launch mode for the whole scaling curve), so the host never paces the device.
    graph = None
    if not ns.no_graph:
        try:
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap):
                tr.step(*args)
            for _ in range(3):
                graph.replay()
            barrier()
        except Exception as exc:  # capture not possible on this stack: run eagerly, say so
            sys.stderr.write(f"[bench] CUDA-graph capture of the step failed ({exc!r}); timing eager launches\n")
            graph = None
            barrier()

    # ---- timed regions: K steps between two CUDA events on the launching stream, barrier + sync on both sides.
    #      The K-step region is repeated and the MEDIAN region reported (a 20-step region is ~10 ms at N = 1 and
    #      ~1.5 ms at N = 8: one region is at the mercy of a single clock excursion) ----
    sampler = None
    if rank == 0:
        try:
            sampler = Sampler(dt=0.02).start()
        except Exception:
            sampler = None
    K = ns.steps
    regions = ns.regions if ns.regions > 0 else 5
    launches0 = _abi.launch_count()
    t_wall0 = time.time()
    region_ms = []
    for _ in range(regions):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is not None:
            for _ in range(K):
                graph.replay()
        else:
            for _ in range(K):
                tr.step(*args)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        region_ms.append(ms)
    t_wall1 = time.time()
    halo_trace = None
    if ex is not None:
        ctx.check()  # no device-side wait timed out during the timed regions
        halo_trace = ctx.trace()  # device timeline (ns) of the last timed step
    launches_per_step = launches_eager_step  # counted by the library (b2s_launch_count) over one eager step before the capture
    launches = regions * K * launches_per_step if graph is not None else _abi.launch_count() - launches0
    elapsed_ms = statistics.median(region_ms)
    ms_per_step = elapsed_ms / K
    value = total_points * K / (elapsed_ms * 1e-3)

    # ---- the dominant launch (fv_tp2d on the GPU's whole batch), timed in an eager pass of exchange-then-stencil
    #      steps right after the timed regions: CUDA events on its stream, not under a profiler ----
    k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
    halo_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
    for hv, ev in zip(halo_events, k_events):
        hv[0].record()
        if ex is not None:
            ex.update()
        else:
            tr.updater.update(q)
        hv[1].record()
        ev[0].record()
        full_call()
        ev[1].record()
    barrier()
    kernel_ms = statistics.median(a.elapsed_time(b) for a, b in k_events)
    halo_ms = statistics.median(a.elapsed_time(b) for a, b in halo_events)
    hws_summary = None
    if sampler is not None:
        time.sleep(0.05)
        sampler.stop()
        clocks = sampler.clocks_summary(local_rank, since=t_wall0, until=t_wall1 + 0.05)
        if not clocks["samples"]:
            clocks = sampler.clocks_summary(local_rank)
        # the repo's hardware sampler (b200stencil.hws, successor of tcn.hws): every visible GPU, 50 Hz
        d = sampler.dump_dict()
        sel = [i for i, t in enumerate(d["timestamps"]) if t_wall0 <= t <= t_wall1 + 0.05] or list(range(len(d["timestamps"])))
        if sel:
            ng = len(d["gpu_indices"])
            hws_summary = {
                "samples": len(sel), "dt_s": sampler.dt, "gpus": ng,
                "gpu_power_w_mean": [round(sum(d["gpu_psu"][i][g] for i in sel) / len(sel), 1) for g in range(ng)],
                "gpu_util_pct_mean": [round(sum(d["gpu_exe_utl"][i][g] for i in sel) / len(sel), 1) for g in range(ng)],
                "gpu_mem_used_mib_max": [round(max(d["gpu_mem"][i][g] for i in sel)) for g in range(ng)],
            }  # fmt: skip
        if ns.hws_dump:
            from b200stencil.hws import server as hws_server

            hws_server.dump(sampler, os.path.splitext(ns.hws_dump)[0], "npz")
    else:
        clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: NVML sampler failed to start"]}

    # ---- roofline of the dominant kernel ----
    peaks = harness.measured_peaks(ROOT)
    kernel_points = local_points
    small = kernel_points < 12000000
    kernel_name = ("k_fv_tma" if small else "k_fv_stream") + " (fv_tp2d on the GPU's whole batch of sub-domains)"
    kbytes = kernel_points * algorithmic_bytes_per_point(es)
    achieved = kbytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "fv_stream_traffic.json")  # ncu --set full capture of the same launch
    if os.path.exists(prof) and n_gpus == 1:  # the capture is of the N = 1 launch (6 x 384 x 384 x 72)
        with open(prof) as f:
            traffic = json.load(f).get(ns.dtype, {}).get("dram_bytes_per_launch")
    roofline = {
        "bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic, "kernel": kernel_name,
        "kernel_ms": round(kernel_ms, 4), "algorithmic_bytes_per_launch": kbytes, "peak_source": peaks["source"],
        "frac_of_nominal_8TBs": round(achieved / harness.NOMINAL_HBM_GBS, 4),
        "timed": "CUDA events around the ungated launch in 60 eager exchange-then-stencil steps after the timed regions",
        "halo_exchange_ms": round(halo_ms, 4),
        "step_frac": round(total_points / n_gpus * algorithmic_bytes_per_point(es) / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
    }  # fmt: skip

    # ---- e2e: the same step through the host-buffer API (pinned host fields; H2D, halo update, stencil and D2H
    #      inside the timed region) ----
    e2e = None
    if not ns.skip_e2e:
        pcie = hostio.measure_pcie(dev)
        pipe = hostio.FvTp2dHost(ni, nj, NK, dtype, dev)
        host = pipe.host_fields(nsub)
        for name, t in zip(pipe.NAMES, args[:6]):
            host[name].copy_(t)  # the synthetic inputs, now living on the host
        torch.cuda.synchronize()
        hx = ex.update if ex is not None else (lambda: tr.updater.update(q))
        hsync = ctx.barrier if world > 1 else None
        pipe(host, q_dev=q, exchange=hx, sync=hsync)  # warm-up
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        h0.record()
        for _ in range(ns.e2e_steps):
            pipe(host, q_dev=q, exchange=hx, sync=hsync)
        h1.record()
        barrier()
        e2e_ms = max(h0.elapsed_time(h1), (time.perf_counter() - t0) * 1e3)  # pipe() returns host-synchronised
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        # parity of the host path with the resident path (same inputs, same halo update) on this rank
        tr.step(*args)
        torch.cuda.synchronize()
        ok = bool(torch.equal(host["q_out"].to(dev), q_out))
        step_s = e2e_ms * 1e-3 / ns.e2e_steps
        e2e = {
            "value": total_points * ns.e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": pipe.h2d_bytes * world, "d2h_bytes_per_step": pipe.d2h_bytes * world,
            "ms_per_step": e2e_ms / ns.e2e_steps, "steps": ns.e2e_steps, "matches_resident_path": ok,
            "includes_halo_update": True,
            "api": "b200stencil.hostio.FvTp2dHost (pinned host fields; q uploaded, halo update, then a 3-stream upload|compute|download pipeline over the sub-domains)",
            "pcie_gbs": pcie,
            # the step moves h2d + d2h bytes per GPU over a full-duplex link: its floor is the longer of the two directions
            "frac_of_pcie": round(max(pipe.h2d_bytes / (pcie["h2d"] * 1e9), pipe.d2h_bytes / (pcie["d2h"] * 1e9)) / step_s, 4),
            "numa_node": numa,
        }  # fmt: skip
        del pipe, host

    # ---- cpu_baseline: rank 0, N = 1 only ----
    cpu = None
    cpu_numpy = None
    if rank == 0 and n_gpus == 1 and not ns.skip_cpu:
        _, cpu, _ = time_cpu_port(ns.dtype, budget_s=12.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu_numpy = time_numpy_port(ns.dtype)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": max(3, ns.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": ns.dtype, "data": "synthetic",
            "config": {
                "workload": "fv_tp2d transport step on C384x72 (BASELINE configs[3]): halo update of q + PPM flux-form update",
                "grid": f"C{CUBE_N}", "tiles": 6, "levels": NK, "halo": HALO, "layout": list(layout_for(n_gpus)),
                "subdomains_per_gpu": nsub, "subdomain": [ni, nj], "overlap_exchange": bool(tr.overlap),
                "launch": "cuda-graph replay of the whole step" if graph is not None else "eager launches",
                "l2": f"inputs larger than L2: {input_mb:.0f} MB of inputs per GPU per step vs 126 MB L2, no flush needed",
                "step_launch": (step_mode if ex is not None else "nccl"),
                "halo_exchange": (("device, fused: ONE kernel per step (b2s_halo_fv_tp2d) -- neighbour handshake + pull over NVLink peer memory "
                                   "shared among the CTAs of the stencil grid, then fv_tp2d behind per-sub-domain gates" if tr.fused else
                                   "device: library-owned exchange (b2s_halo_*: a one-block handshake kernel -- announce, await the neighbours -- then the strip copies; same-GPU strips pulled, strips that cross NVLink "
                                   + {"staged": "pushed packed into the destination's staging area and unpacked there", "inplace": "pushed into the halo cells",
                                      "off": "pulled in place"}[ns.push if not tr.overlap else "off"] + ")"
                                   + (", forked beside one gated fv_tp2d launch (sub-domain b computed while the halos of b+1.. arrive)" if tr.overlap else ", then fv_tp2d"))
                                  if ex is not None else
                                  "nccl baseline: pack kernel + grouped NCCL send/recv + unpack kernel" if n_gpus > 1 else
                                  "local halo_move kernel"),
                "halo_bytes_over_nvlink_per_gpu_per_step": (ex.remote_bytes if ex is not None else tr.updater.bytes_sent_per_update),
                "timed_regions": regions, "region_ms": [round(x, 4) for x in region_ms], "reported": "median region",
            },
            "halo_check": halo_check, "device_step_equals_nccl_step": nccl_equal, "halo_trace_ns": halo_trace,
            "roofline": roofline, "clocks": clocks, "hws": hws_summary, "gpu_launches": int(launches),
            "e2e": e2e, "cpu_baseline": cpu,
        }  # fmt: skip
        if cpu_numpy is not None:
            line["cpu_baseline_numpy"] = cpu_numpy
        print(json.dumps(line), flush=True)
    graph = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ctx.finalize()
    if world > 1:
        # A CUDA graph that held NCCL kernels must die before the communicator, or the destroy blocks:
        # the graph is gone, the device drained; leave without the (hang-prone) communicator teardown.
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def run_chain(ns) -> int:
    """BASELINE configs[4]: combined dycore-step chain (horizontal FV + vertical remap scan) on C720x137.

    Not the contract line (that is the C384x72 transport step): prints one JSON line with the same timing
    protocol (barrier + sync, CUDA events, max over ranks) and the unfused algorithmic byte count 96.1 B/pt.
    """
    import torch
    import torch.distributed as dist

    from b200stencil import fields, stencils
    from b200stencil.bench import harness
    from b200stencil.halo.partitioner import CubedSpherePartitioner, layout_for
    from b200stencil.halo.transport import DycoreChain, FvTransport

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, nk = 720, 137
    dtype = torch.float64 if ns.dtype == "f64" else torch.float32
    es = 8 if ns.dtype == "f64" else 4
    part = CubedSpherePartitioner(n, layout_for(world), HALO)
    nsub, ni, nj = part.subdomains_per_gpu(world), part.nx, part.ny
    g = torch.Generator(device=dev)
    g.manual_seed(20240724 + 5 + 1000 * rank)
    mk = lambda s, lo, hi: fields.empty(s, dtype, dev, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    from b200stencil.halo.device import HaloContext

    ctx = HaloContext(rank, world, local_rank)
    exchange = "device"
    q = ctx.field((ni + 6, nj + 6, nk), nsub, dtype, part=part)
    q.uniform_(0.5, 1.5, generator=g)
    ex = ctx.plan(q, part)
    crx, cry = mk((ni + 1, nj, nk), -0.9, 0.9), mk((ni, nj + 1, nk), -0.9, 0.9)
    xfx = mk((ni + 1, nj, nk), 0.9, 1.1).mul_(crx)
    yfx = mk((ni, nj + 1, nk), 0.9, 1.1).mul_(cry)
    rarea = mk((ni, nj), 0.9, 1.1)
    delp = mk((ni, nj, nk), 0.5e5 / nk, 1.5e5 / nk)
    pe1 = fields.empty((ni, nj, nk + 1), dtype, dev, batch=nsub)
    stencils.pe_prefix(delp, 1.0, pe1)
    sig = (torch.arange(nk + 1, device=dev, dtype=torch.float64) / nk).to(dtype)
    pe2 = fields.empty((ni, nj, nk + 1), dtype, dev, batch=nsub)
    pe2[...] = pe1[..., :1] + (pe1[..., -1:] - pe1[..., :1]) * sig
    pe2[..., -1] = pe1[..., -1]
    q_adv = fields.empty((ni, nj, nk), dtype, dev, batch=nsub)
    q_new = fields.empty((ni, nj, nk), dtype, dev, batch=nsub)
    step_mode = "serial" if ns.no_overlap else ("fused" if ns.step == "auto" else ns.step)
    chain = DycoreChain(FvTransport(part, world, rank, overlap=step_mode != "serial", fused=step_mode == "fused", exchange=exchange,
                                    halo_exchange=ex), fused=ns.fused_remap)
    args = (q, crx, xfx, cry, yfx, rarea, delp, pe2, q_adv, pe1, q_new)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, ns.warmup)):
        chain.step(*args)
    barrier()
    graph = None
    if not ns.no_graph:
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            chain.step(*args)
        graph.replay()
        barrier()
    K = ns.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        graph.replay() if graph is not None else chain.step(*args)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_points = 6 * n * n * nk
    bpp = (6 * es + es / nk) + 2 * es + 4 * es
    peaks = harness.measured_peaks(ROOT)
    if rank == 0:
        gbs = total_points / world * bpp / (ms / K * 1e-3) / 1e9
        print(json.dumps({
            "metric": "grid-cells x levels per second, dycore chain fv_tp2d + pe_prefix + remap (C720x137, 6 tiles)",
            "value": total_points * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, ns.warmup),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": ns.dtype,
            "data": "synthetic",
            "config": {"workload": "dycore chain on C720x137 (BASELINE configs[4])", "subdomains_per_gpu": nsub,
                       "subdomain": [ni, nj], "halo_exchange": "device (k_halo_exchange, b2s_halo_*)",
                       "step_launch": step_mode,
                       "launch": "cuda-graph replay" if graph is not None else "eager launches",
                       "vertical": "remap_delp (pe_prefix fused into remap)" if ns.fused_remap else "pe_prefix + remap"},
            "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(gbs / peaks["hbm_gbs"], 4), "traffic": None,
                         "kernel": "whole chain (unfused algorithmic bytes 96.1 B/pt in fp64), per GPU"},
        }), flush=True)  # fmt: skip
    graph = None
    torch.cuda.synchronize()
    ctx.check()
    if world > 1:
        dist.barrier()
    ctx.finalize()
    if world > 1:
        sys.stdout.flush()
        os._exit(0)
    return 0


def run_patterns(ns) -> int:
    """BASELINE configs[1]: the dsl_patterns column stencils at their own size, C96 (6 tiles) x 72 levels, fp64 and fp32.

    Not the contract line.  Each stencil is timed as CUDA-graph replays over rotating buffer sets (32 MB fields are
    smaller than L2: the sets rotate so that the working set is > 2.5x L2), CUDA events around the replays; one JSON
    line with a roofline entry per stencil (algorithmic bytes of SURVEY.md 8d)."""
    import torch

    from b200stencil.bench import harness, sweep

    torch.cuda.set_device(0)
    rows = []
    for stencil in ("top_of_column", "while_in_function", "hybrid_index_2dout"):
        for dt in (torch.float64, torch.float32):
            rows.append(sweep.run_one(stencil, "C96x72", dt, iters=max(5, min(ns.steps, 30)), warmup=max(3, min(ns.warmup, 10)), graph=True))
    peaks = harness.measured_peaks(ROOT)
    pts = sum(r["points"] for r in rows)
    ms = sum(r["median_ms"] for r in rows)
    worst = min(rows, key=lambda r: r["frac_measured_peak"])
    print(json.dumps({
        "metric": "grid-cells x levels per second, dsl_patterns column stencils (C96x72, 6 tiles), fp64 + fp32",
        "value": pts / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": max(5, min(ns.steps, 30)), "warmup": max(3, min(ns.warmup, 10)),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": {"workload": "dsl_patterns S1 top_of_column, S2 while_in_function, S3 hybrid_index_2dout on C96x72 (BASELINE configs[1])",
                   "launch": "cuda-graph replay, rotating buffer sets (working set > 2.5x L2)", "step": "one launch of each stencil in each precision"},
        "roofline": {"bound": "hbm", "achieved": worst["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": worst["frac_measured_peak"],
                     "traffic": None, "kernel": f"{worst['stencil']} {worst['dtype']} (the lowest of the six)"},
        "stencils": {f"{r['stencil']}_{r['dtype']}": {"median_us": round(r["median_ms"] * 1e3, 2), "GBps": r["GBps"], "frac": r["frac_measured_peak"],
                                                      "frac_of_nominal_8TBs": r["frac_nominal_8TBs"], "gpts_per_s": r["gpts_per_s"]} for r in rows},
    }), flush=True)  # fmt: skip
    return 0


if __name__ == "__main__":
    sys.exit(main())
