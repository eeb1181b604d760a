#!/usr/bin/env python
"""Headline benchmark: grid-cells x levels per second of the FV transport step on C384x72.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the configuration the headline target is quoted on):
FV3-style horizontal finite-volume flux/advection stencil (fv_tp2d, 3-cell halo) on the C384 cubed
sphere (6 tiles x 384 x 384 columns) x 72 levels, synthetic fields.  One STEP = halo update of q from
the neighbouring sub-domains (same-GPU copies at N = 1, + NCCL exchange over NVLink at N > 1) followed
by fv_tp2d on every sub-domain the GPU hosts.  The domain is fixed, so scaling is STRONG.

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


def algorithmic_bytes_per_point(es: int) -> float:
    """40 R (q, crx, xfx, cry, yfx) + 8 W + 8/nk (rarea) in fp64; scales with the element size."""
    return 6 * es + es / NK


# ---------------------------------------------------------------------------------------------------
# CPU side: the oracle port, timed on the host cores (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------


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


def time_cpu_port(dtype_name: str, budget_s: float = 20.0, steps=None, warmup: int = 1):
    """Time the C/OpenMP restatement (oracle/c) on ONE tile of the workload (384 x 384 x 72), all host threads.

    Returns (points_per_second, description dict).  The sample is 1/6 of a step of the real workload.
    """
    import numpy as np

    from oracle.c_oracle import COracle

    dtype = np.float64 if dtype_name == "f64" else np.float32
    try:
        orc = COracle(native=True)  # -O3 -march=native, built on this machine
        build = "gcc -O3 -march=native -fopenmp"
    except Exception:
        orc = COracle(native=False)
        build = "gcc -O2 -fopenmp (portable build)"
    threads = orc.threads
    ni = nj = CUBE_N
    args = cpu_sample_inputs(ni, nj, NK, dtype)
    pts = ni * nj * NK
    for _ in range(warmup):
        orc.fv_tp2d(*args)
    times = []
    t_begin = time.perf_counter()
    n = 0
    while True:
        t0 = time.perf_counter()
        orc.fv_tp2d(*args)
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
        "value": pts / med, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"oracle/c fv_tp2d ({build}) on 1 of the 6 tiles (384x384x{NK}, {dtype_name}), "
                  f"{n} runs, median {med * 1e3:.1f} ms, best {best * 1e3:.1f} ms",
        "ms_per_sample": med * 1e3,
    }  # fmt: skip
    return pts / med, info, times


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
    here, so this is the oracle port (C/OpenMP, every host thread) on the same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, info, times = time_cpu_port(ns.dtype, steps=max(1, ns.steps), warmup=max(1, ns.warmup))
    ms = statistics.median(times) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ns.gpus, "steps": len(times),
        "warmup": max(1, ns.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": ns.dtype, "data": "synthetic",
        "config": {"workload": "fv_tp2d C384x72 (CPU port; each step = 1 of the 6 tiles, 384x384x72)",
                   "grid": f"C{CUBE_N}", "levels": NK, "halo": HALO},
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
    ap.add_argument("--workload", choices=["fv", "chain"], default="fv",
                    help="fv (default, the contract line): fv_tp2d transport step on C384x72; chain: BASELINE configs[4], "
                         "fv_tp2d + pe_prefix + remap on C720x137 (a separate report, see run_chain)")
    ap.add_argument("--no-overlap", action="store_true", help="exchange, then compute (no interior/frame split)")
    ap.add_argument("--halo", choices=["auto", "p2p", "p2p-fused", "nccl"], default="auto",
                    help="multi-GPU halo exchange: p2p = device barrier + one peer-memory pull kernel over NVLink "
                         "(torch symmetric memory); nccl = packed strips + grouped NCCL send/recv overlapped with the "
                         "interior; auto = p2p when it can be set up, else nccl; p2p-fused = EXPERIMENTAL one-launch "
                         "handshake + pull (halo_pull_sync), not part of the default path")
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph (default when --gpus > 1)")
    ap.add_argument("--no-graph", action="store_true", help="always launch eagerly")
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

    import torch
    import torch.distributed as dist

    from b200stencil import _abi, fields, hostio, stencils
    from b200stencil.bench import harness
    from b200stencil.halo.partitioner import CubedSpherePartitioner, layout_for
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if ns.dtype == "f64" else torch.float32
    es = 8 if ns.dtype == "f64" else 4

    part = CubedSpherePartitioner(CUBE_N, layout_for(n_gpus), HALO)
    nsub = part.subdomains_per_gpu(n_gpus)
    ni, nj = part.nx, part.ny

    # synthetic fields, resident in HBM before the timed region (SURVEY.md 8d recipe, device RNG)
    g = torch.Generator(device=dev)
    g.manual_seed(20240724 + 4 + 1000 * rank)
    mk = lambda s, lo, hi: fields.empty(s, dtype, dev, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    exchange, sym_q = "nccl", None
    if world > 1 and ns.halo in ("auto", "p2p", "p2p-fused"):
        try:
            from b200stencil.halo.p2p import SymmetricField

            sym_q = SymmetricField((ni + 6, nj + 6, NK), nsub, dtype, dev)
            exchange = "p2p"
        except Exception as exc:
            if ns.halo in ("p2p", "p2p-fused"):
                raise
            sys.stderr.write(f"[bench] symmetric memory unavailable ({exc!r}); using the NCCL exchange\n")
    if sym_q is not None:
        q = sym_q.field.uniform_(0.5, 1.5, generator=g)
    else:
        q = mk((ni + 6, nj + 6, NK), 0.5, 1.5)
    tr = FvTransport(part, n_gpus, rank, overlap=not ns.no_overlap, exchange=exchange, symmetric_q=sym_q,
                     fused_signal=ns.halo == "p2p-fused")
    crx, cry = mk((ni + 1, nj, NK), -0.9, 0.9), mk((ni, nj + 1, NK), -0.9, 0.9)
    xfx = mk((ni + 1, nj, NK), 0.9, 1.1).mul_(crx)
    yfx = mk((ni, nj + 1, NK), 0.9, 1.1).mul_(cry)
    rarea = mk((ni, nj), 0.9, 1.1)
    q_out = fields.empty((ni, nj, NK), dtype, dev, batch=nsub)
    args = (q, crx, xfx, cry, yfx, rarea, q_out)
    total_points = 6 * CUBE_N * CUBE_N * NK
    local_points = nsub * ni * nj * NK
    input_mb = sum(t.numel() * es for t in args[:6]) / 1e6

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(3, ns.warmup)):
        tr.step(*args)
    barrier()

    full_call, interior_call, frame_calls = tr.calls(*args)
    dominant = interior_call if tr.overlap else full_call

    # Multi-GPU steps are a handful of ~50 us kernels plus an NCCL group: the whole step (both streams,
    # NCCL included) is captured ONCE into a CUDA graph and replayed, so the host never paces the device.
    use_graph = (n_gpus > 1 and not ns.no_graph) or ns.graph
    graph = None
    if use_graph:
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

    # ---- timed region: K steps between two CUDA events on the launching stream, barrier + sync on
    #      both sides; eager mode also brackets the dominant fv_tp2d launch of every step ----
    sampler = None
    if rank == 0:
        try:
            sampler = Sampler(dt=0.02).start()
        except Exception:
            sampler = None
    K = ns.steps
    # the dominant launch is bracketed on (up to) 200 steps spread evenly over the timed region, so a clock
    # that sags under the power cap late in a long loop is seen by the kernel time as it is by the step time
    k_stride = max(1, K // 200)
    k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                for _ in range(len(range(0, K, k_stride)))]

    def eager_step_with_events(ev):
        """tr.step() with two events around the dominant launch (same kernels, same order)."""
        if tr.p2p is not None:
            tr.p2p.update()
            ev[0].record()
            full_call()
            ev[1].record()
        elif tr.overlap:
            tr.updater.start(q)
            ev[0].record()
            interior_call()
            ev[1].record()
            tr.updater.wait()
            for c in frame_calls:
                c()
        else:
            tr.updater.update(q)
            ev[0].record()
            full_call()
            ev[1].record()

    launches0 = _abi.launch_count()
    t_wall0 = time.time()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is not None:
        for _ in range(K):
            graph.replay()
    else:
        for it in range(K):
            if it % k_stride == 0:
                eager_step_with_events(k_events[it // k_stride])
            else:
                tr.step(*args)
    e1.record()
    barrier()
    t_wall1 = time.time()
    if tr.p2p is not None:
        launches_per_step = 2  # halo_pull + fv_tp2d (the barrier kernel is torch's, not counted)
    else:
        launches_per_step = (1 if not tr.updater.plan.peers else 3) + (1 if not tr.overlap else 1 + len(frame_calls))
    launches = K * launches_per_step if graph is not None else _abi.launch_count() - launches0
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / K
    value = total_points * K / (elapsed_ms * 1e-3)
    if graph is not None:
        # a replayed graph leaves no place for events: the dominant launch is timed in an eager pass of the
        # same step loop right after the timed region (CUDA events on its stream, not under a profiler)
        for ev in k_events[:60]:
            eager_step_with_events(ev)
        barrier()
        kernel_ms = statistics.median(a.elapsed_time(b) for a, b in k_events[:60])
    else:
        kernel_ms = statistics.median(a.elapsed_time(b) for a, b in k_events)
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
    if tr.overlap:
        r = tr.interior
        kernel_points = nsub * (r[1] - r[0]) * (r[3] - r[2]) * NK
        kernel_name = "k_fv_stream (interior rectangle launch)"
    else:
        kernel_points = local_points
        kernel_name = "k_fv_stream (full-domain launch)"
    kbytes = kernel_points * algorithmic_bytes_per_point(es)
    achieved = kbytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "fv_stream_traffic.json")  # ncu --set full capture of the same launch
    if os.path.exists(prof) and not tr.overlap and n_gpus == 1:  # the capture is of the N = 1 launch (6 x 384 x 384 x 72)
        with open(prof) as f:
            traffic = json.load(f).get(ns.dtype, {}).get("dram_bytes_per_launch")
    roofline = {
        "bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic, "kernel": kernel_name,
        "kernel_ms": round(kernel_ms, 4), "algorithmic_bytes_per_launch": kbytes, "peak_source": peaks["source"],
        "frac_of_nominal_8TBs": round(achieved / harness.NOMINAL_HBM_GBS, 4),
    }  # fmt: skip

    # ---- e2e: same step through the host-buffer API (pinned host fields, H2D + D2H inside the timed region) ----
    e2e = None
    if not ns.skip_e2e:
        pipe = hostio.FvTp2dHost(ni, nj, NK, dtype, dev)
        host = pipe.host_fields(nsub)
        for name, t in zip(pipe.NAMES, args[:6]):
            host[name].copy_(t)  # the synthetic inputs, now living on the host
        torch.cuda.synchronize()
        pipe(host)  # warm-up
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        h0.record()
        for _ in range(ns.e2e_steps):
            pipe(host)
        h1.record()
        barrier()
        e2e_ms = max(h0.elapsed_time(h1), (time.perf_counter() - t0) * 1e3)  # pipe() returns host-synchronised
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        # parity of the host path with the resident path on this rank's first sub-domain
        ok = bool(torch.equal(host["q_out"][0].to(dev), _fv_reference_of(stencils, fields, args, dev)))
        e2e = {
            "value": total_points * ns.e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": pipe.h2d_bytes * world, "d2h_bytes_per_step": pipe.d2h_bytes * world,
            "ms_per_step": e2e_ms / ns.e2e_steps, "steps": ns.e2e_steps, "matches_resident_path": ok,
            "api": "b200stencil.hostio.FvTp2dHost (pinned host fields, 3-stream upload|compute|download pipeline)",
        }  # fmt: skip
        del pipe, host

    # ---- cpu_baseline: rank 0, N = 1 only ----
    cpu = None
    cpu_numpy = None
    if rank == 0 and n_gpus == 1 and not ns.skip_cpu:
        _, cpu, _ = time_cpu_port(ns.dtype, budget_s=15.0)
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
                "subdomains_per_gpu": nsub, "subdomain": [ni, nj], "overlap_exchange": tr.overlap,
                "launch": "cuda-graph replay of the whole step" if graph is not None else "eager launches",
                "l2": f"inputs larger than L2: {input_mb:.0f} MB of inputs per GPU per step vs 126 MB L2, no flush needed",
                "halo_exchange": ("none (all neighbours on this GPU: one local halo_move kernel)" if n_gpus == 1 else
                                  "p2p-fused (experimental): one halo_pull_sync kernel, handshake inside" if tr.p2p is not None and tr.p2p.fused_signal else
                                  "p2p: device barrier + one halo_pull kernel over NVLink peer memory" if tr.p2p is not None else
                                  "nccl: pack kernel + grouped NCCL send/recv + unpack kernel"),
                "halo_bytes_over_nvlink_per_gpu_per_step": (tr.p2p.remote_bytes if tr.p2p is not None
                                                            else tr.updater.bytes_sent_per_update),
            },
            "roofline": roofline, "clocks": clocks, "hws": hws_summary, "gpu_launches": int(launches),
            "e2e": e2e, "cpu_baseline": cpu,
        }  # fmt: skip
        if cpu_numpy is not None:
            line["cpu_baseline_numpy"] = cpu_numpy
        print(json.dumps(line), flush=True)
    if world > 1:
        # A CUDA graph that holds NCCL kernels must die before the communicator, or the destroy blocks:
        # drop it, drain the device, and leave without the (hang-prone) communicator teardown.
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
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
    exchange, sym_q = "nccl", None
    if world > 1:
        from b200stencil.halo.p2p import SymmetricField

        sym_q = SymmetricField((ni + 6, nj + 6, nk), nsub, dtype, dev)
        exchange = "p2p"
        q = sym_q.field.uniform_(0.5, 1.5, generator=g)
    else:
        q = mk((ni + 6, nj + 6, nk), 0.5, 1.5)
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
    chain = DycoreChain(FvTransport(part, world, rank, overlap=False, exchange=exchange, symmetric_q=sym_q),
                        fused=ns.fused_remap)
    args = (q, crx, xfx, cry, yfx, rarea, delp, pe2, q_adv, pe1, q_new)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, ns.warmup)):
        chain.step(*args)
    barrier()
    graph = None
    if world > 1 and not ns.no_graph:
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
                       "subdomain": [ni, nj], "halo_exchange": exchange if world > 1 else "local",
                       "launch": "cuda-graph replay" if graph is not None else "eager launches",
                       "vertical": "remap_delp (pe_prefix fused into remap)" if ns.fused_remap else "pe_prefix + remap"},
            "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(gbs / peaks["hbm_gbs"], 4), "traffic": None,
                         "kernel": "whole chain (unfused algorithmic bytes 96.1 B/pt in fp64), per GPU"},
        }), flush=True)  # fmt: skip
    if world > 1:
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)
    return 0


def _fv_reference_of(stencils, fields, args, dev):
    """fv_tp2d of sub-domain 0 recomputed on resident fields WITHOUT a halo update (what the host path computes)."""
    import torch

    q, crx, xfx, cry, yfx, rarea, q_out = args
    out = fields.empty(tuple(q_out.shape[1:]), q_out.dtype, dev)
    stencils.fv_tp2d(q[0], crx[0], xfx[0], cry[0], yfx[0], rarea[0], out)
    torch.cuda.synchronize()
    return out


if __name__ == "__main__":
    sys.exit(main())
