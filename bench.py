#!/usr/bin/env python
"""Benchmark of the SSG loss hot path (BASELINE.json metric: edge-pixels/sec, SSG fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the whole SSL block (edge list, SSG rows of SR and GT, L1, backward into
SR) over one batch of synthetic crops.  Workload = BASELINE.json configs[1]: batch 16 of 256x256
fp32 crops, k_search=25, k_window=9, sigma=0.004, Bernoulli(0.114) edge mask, per GPU
(weak scaling: rank r owns its own 16 crops, seed 1+r; the only exchange is the 24-byte all-reduce
of [sum|d|, sumKL, n_rows] at the end of the forward).

Reported on ONE JSON line (rank 0):
  value      edge-px/s with inputs resident in HBM, CUDA-event timed per step, max over ranks;
             L2 is flushed (256 MiB write) between timed steps, outside the timed events
  e2e        the same metric through the C ABI's host entry (ssl_b200_loss_step_host): pinned host
             sr/gt/mask -> device, step, loss + gradient -> host, all inside the timed region
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration (events recorded by
             the library around the kernel, on its stream, during the timed steps) vs the measured HBM
             peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle's restatement of the reference ssl_pytorch (oracle/ssl_oracle.py) timed on
             this box's host cores on a bounded sample of the same workload
  gpu_baseline  the reference's own CUDA operator (similarity.cu, compiled unmodified by oracle/build_ref.py
             into oracle/_ref) timed on this GPU on the same batch, per image as the reference does
             (similaritywrapper.py:25-69): fwd(SR) + fwd(GT) + bwd -- the GPU bar to beat (rank 0, N=1)
  config3    BASELINE configs[2] measured in the same run: ONE batch of 64 bf16 crops (seed 2) sharded over
             the N ranks by ssl_b200.dist.shard_range, parity="global" (24-byte all-reduce inside the timed
             step) -- strong scaling; its loss is checked against the fp64 oracle value of the whole batch
The loss of every timed workload is asserted against tests/golden/bench_loss.json (fp64 oracle, 1e-5).
`--impl reference` times the CPU restatement as the whole run (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KS, KW, SIGMA, EPS = 25, 9, 0.004, 1e-10
BATCH_PER_GPU, HEIGHT, WIDTH, DENSITY = 16, 256, 256, 0.114
L = KS * KS
# algorithmic HBM reads per edge pixel, fp32: three search-tile gathers of
# C*k_s^2*4 = 7,500 B (SR fwd, GT fwd, SR bwd) + 8 B position = 22,508 B
TILE_BYTES = 3 * L * 4
ALGO_BYTES_STEP = 3 * TILE_BYTES + 8
ALGO_BYTES_FWD = 2 * TILE_BYTES + 8      # the forward launch gathers the SR and the GT tile
ALGO_BYTES_BWD = TILE_BYTES              # the backward launch gathers the SR tile again
FALLBACK_HBM_GBS = 6650.0                # B200_PROFILING.md fallback


def baseline_metric():
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            return json.load(f)["metric"]
    except Exception:
        return "edge-pixels/sec SSG fwd+bwd, 256×256 k_s=25 k_w=9, 1/2/4/8 B200"


def workload_config(n_gpus):
    return {"workload": "configs[1]: batch=16 256x256 fp32 crops per GPU, k_search=25 k_window=9 sigma=0.004, "
                        "Bernoulli(0.114) edge mask (~7.5k edge px per crop), fwd(SR)+fwd(GT)+L1+bwd",
            "global_batch": BATCH_PER_GPU * n_gpus, "crop": [HEIGHT, WIDTH], "k_search": KS, "k_window": KW,
            "mask_density": DENSITY, "parallelism": f"dp{n_gpus} (images sharded, 24-byte all-reduce)",
            "l2": "flushed between timed steps (256 MiB write, outside the timed events)"}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while a region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                index = int(visible.split(",")[index]) if visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference ssl_pytorch
# ---------------------------------------------------------------------------------------------

CPU_SAMPLE_PX = 4096   # >= 4,096 edge pixels, fixed (never resized from a cold call)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_sample(n_px=CPU_SAMPLE_PX, seed=1):
    """THE bounded sample of the workload, the same for cpu_baseline and --impl reference: crop 0 of the
    config-2 batch (seed 1), mask cut to the first row bands that hold >= n_px edge pixels."""
    import torch
    from ssl_b200 import synth
    sr, gt, mask = synth.make_case(1, HEIGHT, WIDTH, seed=seed, density=DENSITY)
    per_row = mask[0, 0].sum(dim=1).cumsum(0)
    rows = int((per_row < n_px).sum().item()) + 1
    m = torch.zeros_like(mask)
    m[:, :, :rows] = mask[:, :, :rows]
    return sr, gt, m, int(m.sum().item())


def cpu_step(sr, gt, mask):
    from oracle import ssl_oracle
    t0 = time.perf_counter()
    loss, grad, n = ssl_oracle.ssl_step_pytorch_port(sr, gt, mask, KS, KW, SIGMA, True, EPS, 1.0, max_px=512)
    return time.perf_counter() - t0, n


def cpu_sample_text(n):
    return (f"{n} edge px = crop 0 of the workload (seed 1) cut to its first row bands holding >= {CPU_SAMPLE_PX} "
            f"edge px; fwd(SR)+fwd(GT)+L1+bwd with oracle.ssl_step_pytorch_port (op-for-op restatement of the "
            f"reference ssl_pytorch, 512-px chunks; the reference itself is Python under /root/reference and does "
            f"not travel to the GPU box); full-batch figures are a linear extrapolation")


def cpu_baseline():
    """Edge-px/s of the CPU restatement on the fixed sample: one warm-up call, then the best of 3."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    sr, gt, m, n = cpu_sample()
    cpu_step(sr, gt, m)
    times = [cpu_step(sr, gt, m)[0] for _ in range(3)]
    return {"value": n / min(times), "unit": "edge-pixels/s", "cores": torch.get_num_threads(), "kind": "port",
            "cpu_model": cpu_model(), "host_cpus": os.cpu_count(), "best_of": 3,
            "times_s": [round(t, 3) for t in times], "sample": cpu_sample_text(n)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    sr, gt, m, n = cpu_sample()
    for _ in range(args.warmup):
        cpu_step(sr, gt, m)
    times = [cpu_step(sr, gt, m)[0] for _ in range(args.steps)]
    t = sum(times)
    value = n * args.steps / t
    out = {"impl": "reference", "metric": baseline_metric(), "value": value, "unit": "edge-pixels/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.gpus),
           "cpu_baseline": {"value": value, "unit": "edge-pixels/s", "cores": torch.get_num_threads(),
                            "kind": "port", "cpu_model": cpu_model(), "host_cpus": os.cpu_count(),
                            "best_step_value": n / min(times), "sample": "each step = " + cpu_sample_text(n)},
           "e2e": {"value": value, "unit": "edge-pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU baseline-to-beat: the reference's own similarity.cu on the same device
# ---------------------------------------------------------------------------------------------

def gpu_baseline(sr_d, gt_d, mask_h, n_edges, steps=3):
    """Times oracle/_ref/libsimilarity_ref.so (the reference's unmodified CUDA op) on the batch the B200
    arm just processed, the way the reference drives it: per image, fwd(SR) + fwd(GT) into zero-filled
    outputs and one backward into a zero-filled padded gradient (similaritywrapper.py:25-57).  Padding,
    `nonzero` and the exp / normalise / L1 tail are left OUT of the timed region (in the reference's favour).
    The reference launches on the legacy default stream; so are the events."""
    import torch
    from oracle import build_ref
    lib = build_ref.load()
    if lib is None:
        return {"unavailable": "oracle/_ref/libsimilarity_ref.so not built (python -m oracle.build_ref)"}
    dev = sr_d.device
    P = KS // 2
    B = sr_d.shape[0]
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    pads, poss = [], []
    for i in range(B):
        sp = torch.nn.functional.pad(sr_d[i].float(), (P, P, P, P), mode="reflect").contiguous()
        gp = torch.nn.functional.pad(gt_d[i].float(), (P, P, P, P), mode="reflect").contiguous()
        mp = torch.nn.functional.pad(mask_h[i, 0].to(dev), (P, P, P, P))
        pos = torch.nonzero(mp == 1).to(torch.int32).contiguous()
        pads.append((sp, gp))
        poss.append(pos)
    gen = torch.Generator(device="cpu").manual_seed(3)
    grads = [(1e-6 * torch.randn(p.shape[0], L, generator=gen)).to(dev) for p in poss]
    c, hp, wp = pads[0][0].shape
    assert torch.cuda.current_stream().cuda_stream == 0, "gpu_baseline must run on the legacy default stream"

    def one_pass():
        for i in range(B):
            mc = poss[i].shape[0]
            if mc == 0:
                continue
            for img in pads[i]:
                out = torch.zeros(mc, KS, KS, device=dev)
                rc = lib.ref_compute_similarity(vp(img), vp(poss[i]), vp(out), mc, KS, KW, hp, wp, c)
                assert rc == 0
            gi = torch.zeros(c, hp, wp, device=dev)
            rc = lib.ref_compute_similarity_backward(vp(pads[i][0]), vp(grads[i]), vp(poss[i]), vp(gi), mc, KS, KW,
                                                     hp, wp, c)
            assert rc == 0

    one_pass()
    torch.cuda.synchronize()
    times = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one_pass()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) * 1e-3)
    t = min(times)
    return {"value": n_edges / t, "unit": "edge-pixels/s", "ms_per_step": 1e3 * t, "steps": steps, "kind": "reference",
            "what": "reference similarity.cu (unmodified, nvcc sm_100, 16-thread blocks, global fp32 atomics), per "
                    "image: _compute_similarity(SR) + _compute_similarity(GT) + _compute_similarity_backward; "
                    "best of %d passes over the same %d-crop batch; pad/nonzero/exp/normalise/L1 not timed" % (steps, B)}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------

def bench_golden():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_loss.json")) as f:
            return json.load(f)
    except Exception:
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU path; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the communicator is created; stdout must carry exactly
        # one JSON line, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    import ssl_b200
    from ssl_b200 import _lib, synth
    from ssl_b200.dist import shard_range
    lib = _lib.load()
    golden = bench_golden()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_steps(fn, k, w):
        for _ in range(w):
            fn()
        evs = []
        barrier()
        for _ in range(k):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs) * 1e-3

    # ---- workload (BASELINE configs[1], per rank) ------------------------------------------------
    sr_h, gt_h, mask_h = synth.make_case(BATCH_PER_GPU, HEIGHT, WIDTH, seed=1 + rank, density=DENSITY)
    sr_h, gt_h, mask_h = sr_h.pin_memory(), gt_h.pin_memory(), mask_h.pin_memory()
    sr_d, gt_d, mask_d = sr_h.to(dev), gt_h.to(dev), mask_h.to(dev)
    n_edges = int(mask_h.sum().item())
    state = {}

    def step():
        # max_edges = rows capacity known to the caller: the whole step is enqueued without a host sync
        x = sr_d.detach().requires_grad_(True)
        loss = ssl_b200.ssl(x, gt_d, mask_d, KS, KW, SIGMA, True, EPS, loss_weight=1.0, parity="global",
                            max_edges=n_edges)
        loss.backward()
        state["loss"], state["grad"] = loss.detach(), x.grad
        return loss, x.grad

    # ---- value: inputs resident in HBM ------------------------------------------------------
    # The library brackets each of its stages with CUDA events on the launching stream while
    # profiling is on (include/ssl_b200.h: ssl_b200_profile_*), so the per-kernel durations below
    # come from the very launches of the timed region.
    with ClockSampler(local) as clocks:
        for _ in range(args.warmup):
            step()
        _lib.profile_enable(True)
        launches0 = lib.ssl_b200_launch_count()
        t_local = timed_steps(step, args.steps, 0)
        launches = lib.ssl_b200_launch_count() - launches0
        stages = _lib.profile_read()
        _lib.profile_enable(False)
    t_step = max_over_ranks(t_local)
    total_edges = sum_over_ranks(float(n_edges))
    value = total_edges * args.steps / t_step
    total_launches = int(sum_over_ranks(float(launches)))
    # the loss the timed step computed (global mean over all ranks' rows) against the fp64 oracle
    loss_check = None
    if golden is not None and all(str(1 + r) in golden["config2_fp32"] for r in range(world)):
        num = sum(golden["config2_fp32"][str(1 + r)] * golden["config2_rows"][str(1 + r)] for r in range(world))
        want = num / sum(golden["config2_rows"][str(1 + r)] for r in range(world))
        got = float(state["loss"])
        loss_check = {"loss": got, "oracle_fp64": want, "rel_err": abs(got - want) / want}
        assert loss_check["rel_err"] <= 1e-5, f"bench.py: timed step computed loss {got}, fp64 oracle says {want}"
        assert bool(torch.isfinite(state["grad"]).all()) and float(state["grad"].abs().max()) > 0

    # ---- e2e: host buffers through the C ABI --------------------------------------------------
    grad_h = torch.empty_like(sr_h).pin_memory()

    def host_step():
        return ssl_b200.ssl_step_host(sr_h, gt_h, mask_h, KS, KW, SIGMA, True, EPS, 1.0, out_grad=grad_h)

    for _ in range(max(args.warmup, 3)):
        loss_host, _, _ = host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    torch.cuda.synchronize()
    t_host = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": total_edges * args.steps / t_host, "unit": "edge-pixels/s",
           "h2d_bytes_per_step": int(world * (sr_h.numel() + gt_h.numel() + mask_h.numel()) * 4),
           "d2h_bytes_per_step": int(world * (grad_h.numel() * 4 + 12 + 8)),
           "ms_per_step": 1e3 * t_host / args.steps,
           "call": "ssl_b200_loss_step_host (pinned host sr/gt/mask in, loss[3] + d loss/d sr out, wall clock)"}
    if golden is not None and str(1 + rank) in golden["config2_fp32"]:
        want = golden["config2_fp32"][str(1 + rank)]
        assert abs(float(loss_host[0]) - want) <= 1e-5 * want, "bench.py: host step loss disagrees with the oracle"

    # ---- config 3: ONE 64-crop bf16 batch sharded over the ranks (strong scaling) ---------------
    config3 = None
    if not args.no_config3 and 64 % world == 0:
        sr3, gt3, mask3 = synth.make_case(64, HEIGHT, WIDTH, seed=2, density=DENSITY)
        idx = list(shard_range(64, rank, world))
        sr3_d, gt3_d = sr3[idx].bfloat16().to(dev), gt3[idx].bfloat16().to(dev)
        mask3_d = mask3[idx].to(dev)
        n3_local, n3_total = int(mask3[idx].sum().item()), int(mask3.sum().item())
        del sr3, gt3
        st3 = {}

        def step3():
            x = sr3_d.detach().requires_grad_(True)
            loss = ssl_b200.ssl(x, gt3_d, mask3_d, KS, KW, SIGMA, True, EPS, loss_weight=1.0, parity="global",
                                max_edges=n3_local)
            loss.backward()
            st3["loss"], st3["grad"] = loss.detach(), x.grad

        k3 = max(3, args.steps // 2)
        t3 = max_over_ranks(timed_steps(step3, k3, 3))
        config3 = {"workload": "configs[2]: ONE batch of 64 256x256 bf16 crops (seed 2) sharded by image over the "
                               "ranks, fp32 arithmetic, parity=global (NCCL all-reduce of [sum|d|, sumKL, n_rows] "
                               "inside the timed step)",
                   "value": n3_total * k3 / t3, "unit": "edge-pixels/s", "ms_per_step": 1e3 * t3 / k3, "steps": k3,
                   "scaling": "strong", "dtype": "bf16 storage, f32 arithmetic", "crops_per_gpu": 64 // world,
                   "edge_px_per_step": n3_total, "loss": float(st3["loss"])}
        if golden is not None and golden.get("config3_bf16"):
            want = golden["config3_bf16"]["loss"]
            config3["oracle_fp64"] = want
            config3["rel_err"] = abs(config3["loss"] - want) / want
            assert config3["rel_err"] <= 1e-5, f"bench.py: config-3 loss {config3['loss']} vs fp64 oracle {want}"
        del sr3_d, gt3_d, mask3_d, st3

    # ---- roofline of the dominant kernel (rank 0's GPU, from the timed region's own launches) ---
    roof = None
    kernels = {}
    if rank == 0:
        algo = {"ssg_plane_fwd": ALGO_BYTES_FWD, "ssg_point_fwd": ALGO_BYTES_FWD,
                "ssg_plane_bwd": ALGO_BYTES_BWD, "ssg_point_bwd": ALGO_BYTES_BWD}
        for name, (ms, cnt) in stages.items():
            per_step_ms = ms / args.steps
            kernels[name] = {"ms": per_step_ms, "launches_per_step": cnt / args.steps}
            if name in algo:
                kernels[name]["algo_gbs"] = n_edges * algo[name] / (per_step_ms * 1e-3) / 1e9
        peak, peak_src = hbm_peak()
        name = max(algo.keys() & kernels.keys(), key=lambda k: kernels[k]["ms"])
        t_dom = kernels[name]["ms"] * 1e-3
        achieved = n_edges * algo[name] / t_dom / 1e9
        step_gbs = n_edges * ALGO_BYTES_STEP / (t_local / args.steps) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": name, "kernel_ms": 1e3 * t_dom, "peak_source": peak_src,
                "algorithmic_bytes_per_edge_px": algo[name],
                "timing": "CUDA events recorded by the library around this kernel on its launching stream, "
                          "averaged over the launches of the timed region",
                "step": {"achieved": step_gbs, "frac": step_gbs / peak,
                         "algorithmic_bytes_per_edge_px": ALGO_BYTES_STEP,
                         "note": "whole fwd+bwd step (the figure north_star's 70% target is stated on)"}}
        traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    roof["traffic"] = json.load(f).get(name)
            except Exception:
                pass

    # ---- baselines (rank 0, N=1 only) -----------------------------------------------------------
    cpu = gpu_ref = None
    if rank == 0 and world == 1:
        if not args.no_gpu_baseline:
            gpu_ref = gpu_baseline(sr_d, gt_d, mask_h, n_edges)
            if "value" in gpu_ref:
                gpu_ref["speedup_device_timed"] = value / gpu_ref["value"]
        if not args.no_cpu_baseline:
            cpu = cpu_baseline()

    if rank == 0:
        out = {"metric": baseline_metric(), "value": value, "unit": "edge-pixels/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": workload_config(world), "edge_px_per_step": int(total_edges),
               "host_syncs_per_step": 0, "loss_check": loss_check,
               "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "gpu_baseline": gpu_ref, "e2e": e2e,
               "config3": config3, "gpu_launches": total_launches, "clocks": clocks.summary()}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            import subprocess
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
                   os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
                   str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []) + \
                  (["--no-gpu-baseline"] if args.no_gpu_baseline else []) + \
                  (["--no-config3"] if args.no_config3 else [])
            raise SystemExit(subprocess.call(cmd))
        run_b200(args)


if __name__ == "__main__":
    main()
