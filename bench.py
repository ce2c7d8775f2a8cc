#!/usr/bin/env python
"""Headline benchmark: CplxLinearVD forward + KL, samples/s at B=4096, d=4096 per GPU
(BASELINE.json `metric`, configs[2]); weak scaling over N GPUs (batch rows sharded, KL
row-sharded + ONE scalar all-reduce).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.
A "step" = one training-mode forward of the layer on one batch (fused local
reparameterisation, in-kernel Philox noise) + sum(penalties(model)).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B = 4096          # batch rows per GPU
D = 4096          # in = out features
METRIC = "CplxLinearVD forward+KL samples/sec (B=4096, d=4096)"
FLOPS_PER_STEP = 10.0 * B * D * D          # 8 (complex mean GEMM, 4-multiply form) + 2 (variance GEMM)
FWD_ALGO_BYTES = 4.0 * (2 * B * D + 2 * D * D + D * D + 2 * B * D)   # x, W, log_sigma2, y  (fp32)
KL_ALGO_BYTES = 4.0 * 3 * D * D                                       # U, V, log_sigma2
# operand pre-pass (fp32 planes): reads x (2 planes), W (2 planes), log_sigma2; writes 3 + 3 planes
# of 16-bit operands (fp16 re/im + bf16 |x|^2, fp16 U/V + bf16 exp(log_sigma2)) and the row scales
PREP_ALGO_BYTES = 4.0 * (2 * B * D + 3 * D * D) + 2.0 * (3 * B * D + 3 * D * D) + 4.0 * (B + D)
# the GEMM kernel then streams those six 16-bit planes and writes y (fp32)
GEMM_ALGO_BYTES = 2.0 * (3 * B * D + 3 * D * D) + 4.0 * 2 * B * D


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = str(index), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8 and parts[0] == self.index:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# dram bytes (read + write) per launch from the committed ncu --set full captures (profiles/)
NCU_TRAFFIC = {
    "gemm_f32": 631.48e6 + 186.73e6,     # profiles/prof_tc3_f32_r1.raw.csv  (fwd_tc3_kernel<float, cplx>)
    "gemm_bf16": 625.43e6 + 120.13e6,    # profiles/prof_tc3_bf16_r1.raw.csv
    "prepass_f32": 336.51e6 + 172.87e6,  # profiles/prof_prep_f32_r1.raw.csv (vd_prepare_f16_kernel<cplx>)
}


# ----------------------------------------------------------------------------- CPU side
def oracle_step_inputs(rows_x, rows_w, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / (2 * D) ** 0.5                  # reference default init of each weight plane
    w_re = torch.empty(rows_w, D).uniform_(-bound, bound, generator=g)
    w_im = torch.empty(rows_w, D).uniform_(-bound, bound, generator=g)
    bb = 1.0 / rows_w ** 0.5
    b_re = torch.empty(rows_w).uniform_(-bb, bb, generator=g)
    b_im = torch.empty(rows_w).uniform_(-bb, bb, generator=g)
    ls2 = torch.full((rows_w, D), -10.0)
    x_re = torch.randn(rows_x, D, generator=g) / 2 ** 0.5
    x_im = torch.randn(rows_x, D, generator=g) / 2 ** 0.5
    return x_re, x_im, w_re, w_im, b_re, b_im, ls2


def time_oracle(frac_den, repeats=1):
    """Seconds for a 1/frac_den sample of the headline step on the host cores: B/frac_den
    input rows through the full 4096-wide layer's forward, and the KL over 1/frac_den of
    the weight rows (both parts of the step are linear in their row count)."""
    import torch
    from oracle import cplx_oracle as orc
    x_re, x_im, w_re, w_im, b_re, b_im, ls2 = oracle_step_inputs(B // frac_den, D)
    kl_rows = D // frac_den
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            M = x_re.shape[0]
            er, ei = orc.cplx_randn(M, D)
            orc.cplx_linear_vd(x_re, x_im, w_re, w_im, b_re, b_im, ls2, er, ei)
            orc.layer_penalty("cplx_vd", w_re[:kl_rows], w_im[:kl_rows], ls2[:kl_rows], "sum")
            best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    time_oracle(64)                                    # MKL / allocator warm-up
    probe = time_oracle(16)
    total_steps = args.steps + args.warmup
    den = 1
    while den < 16 and probe * 16 / den * total_steps > 150.0:
        den *= 2
    for _ in range(args.warmup):
        time_oracle(den)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_oracle(den)
    dt = time.perf_counter() - t0
    value = (B // den) * args.steps / dt
    sample = (f"per step: {B // den} of {B} batch rows through the 4096x4096 forward + KL over "
              f"{D // den} of {D} weight rows (1/{den} of the headline step), fp32, torch "
              f"{torch.__version__} CPU + scipy expi")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps * den, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CplxLinearVD 4096->4096 forward(train)+KL, batch=4096 (configs[2])",
                   "per_gpu_batch": B, "global_batch": B * args.gpus, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


_ORIG_AFFINITY = []


def unbind_all_threads():
    """Give every thread of this process its original CPU set back (the CPU baseline leg must see
    all host cores, including worker threads spawned while the process was bound)."""
    if not _ORIG_AFFINITY:
        return
    for tid in os.listdir("/proc/self/task"):
        try:
            os.sched_setaffinity(int(tid), _ORIG_AFFINITY[0])
        except OSError:
            pass


def bind_near_gpu(local_rank):
    """Run this process (and first-touch its pinned staging buffers) on the NUMA node the GPU's
    PCIe root port hangs off: host<->device copies that cross the socket interconnect lose a
    large part of the PCIe bandwidth.  Placement only; best effort, returns a short description."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa: single node"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"numa: node {node} has no allowed cpus"
        _ORIG_AFFINITY.append(os.sched_getaffinity(0))
        os.sched_setaffinity(0, allowed)
        return f"numa: bound to node {node} ({len(allowed)} cpus)"
    except Exception as exc:  # noqa: BLE001 -- placement is optional
        return f"numa: not bound ({type(exc).__name__})"


# ----------------------------------------------------------------------------- GPU side
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import cplxmodule_b200 as cb
    from cplxmodule_b200 import cplx
    from cplxmodule_b200.distributed import sharded_penalties
    from cplxmodule_b200.nn.relevance import CplxLinearVD, penalties

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    placement = bind_near_gpu(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cb.set_noise_mode(args.noise)
    if world > 1:
        # N > 1: the pre-pass by-product covers this rank's block of weight rows; ONE all-reduce
        # combines the partial sums on a side stream while the GEMM runs on an SM pair less
        cb.set_kl_shard(rank, world)
        cb.set_sm_reserve(int(os.environ.get("BENCH_SM_RESERVE", "0")))
    torch.manual_seed(0)                               # identical (replicated) parameters
    layer = CplxLinearVD(D, D).to(dev).train()
    if args.dtype == "bf16":
        layer = layer.bfloat16()
    dt = torch.float32 if args.dtype != "bf16" else torch.bfloat16
    torch.manual_seed(1000 + rank)                     # per-rank batch
    host_x = [torch.randn(B, D).div_(2 ** 0.5).to(dt).pin_memory() for _ in range(2)]
    x = cplx.Cplx(host_x[0].to(dev), host_x[1].to(dev))
    host_y = [torch.empty(B, D, dtype=dt).pin_memory() for _ in range(2)]
    host_kl = torch.empty((), dtype=torch.float32).pin_memory()

    def kl_term():
        if world > 1:
            return sharded_penalties(layer)[1].sum()
        return sum(penalties(layer))

    fwd_ms, kl_ms = [], []
    # high priority: when the pre-pass retires, the collective's CTA is placed before the
    # persistent GEMM grid takes every SM
    kl_stream = torch.cuda.Stream(dev, priority=int(os.environ.get("BENCH_KL_PRIO", "-1"))) if world > 1 else None

    def step(record=False):
        """N > 1: this rank's partial KL sum is final right after the forward's pre-pass (its own
        CUDA event); the scalar all-reduce is issued on a side stream and overlaps the GEMM."""
        main_s = torch.cuda.current_stream(dev)
        if record:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if kl_stream is not None:
            kl_stream.wait_stream(main_s)      # parameters of the previous step are settled
        if record:
            e[0].record()
        y = layer(x)
        if record:
            e[1].record()
        if kl_stream is not None:
            with torch.cuda.stream(kl_stream):
                if record:
                    e[2].record()
                kl = sharded_penalties(layer, stream=kl_stream)[1].sum()
                if record:
                    e[3].record()
            kl.record_stream(main_s)
            main_s.wait_stream(kl_stream)
        else:
            if record:
                e[2].record()
            kl = kl_term()
            if record:
                e[3].record()
        if record:
            fwd_ms.append((e[0], e[1])); kl_ms.append((e[2], e[3]))
        return y, kl

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        t.record()
        sync_all()
        ms = torch.tensor([s.elapsed_time(t)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        ms = timed(lambda: step(record=True), args.steps)
        # the timed loop may be shorter than nvidia-smi's sampling period: keep the identical
        # loop running (untimed) until ~0.4 s of load has been sampled
        # (`ms` is the all-reduced maximum, so every rank runs the same number of extra steps)
        extra = int(max(0.0, 400.0 - ms) / max(ms / args.steps, 1e-3)) + 1
        for _ in range(min(extra, 5000)):
            step()
        torch.cuda.synchronize(dev)
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks["window"] = "timed loop + identical untimed continuation, >= 0.4 s under load"

        # ---- end to end: host (pinned) inputs in, host outputs back, EVERY step.
        # Double-buffered: the H2D copy of step i+1 and the D2H copy of step i-1 run on their
        # own streams while step i computes (PCIe is full duplex); every byte still moves
        # inside the timed region.
        main = torch.cuda.current_stream(dev)
        h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        xbuf = [cplx.Cplx(torch.empty_like(x.real), torch.empty_like(x.imag)) for _ in range(2)]
        x_free = [None, None]
        y_done = [None, None]

        def e2e_loop(steps):
            for i in range(steps):
                b = i & 1
                with torch.cuda.stream(h2d):
                    if x_free[b] is not None:
                        h2d.wait_event(x_free[b])
                    xbuf[b].real.copy_(host_x[0], non_blocking=True)
                    xbuf[b].imag.copy_(host_x[1], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(h2d)
                main.wait_event(ready)
                y = layer(xbuf[b])
                kl = kl_term()
                done = torch.cuda.Event()
                done.record(main)
                x_free[b] = done
                with torch.cuda.stream(d2h):
                    d2h.wait_event(done)
                    if y_done[b] is not None:
                        d2h.wait_event(y_done[b])
                    host_y[0].copy_(y.real, non_blocking=True)
                    host_y[1].copy_(y.imag, non_blocking=True)
                    host_kl.copy_(kl.float().reshape(()), non_blocking=True)
                    y.real.record_stream(d2h); y.imag.record_stream(d2h); kl.record_stream(d2h)
                    out_done = torch.cuda.Event()
                    out_done.record(d2h)
                    y_done[b] = out_done
            main.wait_stream(d2h)
            main.wait_stream(h2d)

        e2e_loop(3)
        e2e_steps = max(4, min(args.steps, 20))
        e2e_ms = timed(lambda: e2e_loop(e2e_steps), 1)

        # ---- same step with bf16 planes (BASELINE.json configs[1] precision class), reported
        # beside the fp32/tf32 headline, never instead of it
        alt = None
        if args.dtype == "f32" and not args.no_alt:
            layer16 = CplxLinearVD(D, D).to(dev).train().bfloat16()
            x16 = x.to(torch.bfloat16)
            ev = []

            def step16():
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                layer16(x16)
                b.record()
                ev.append((a, b))
                if world > 1:
                    sharded_penalties(layer16)
                else:
                    sum(penalties(layer16))

            for _ in range(3):
                step16()
            ev.clear()
            ms16 = timed(step16, args.steps)
            f16 = statistics.mean(a.elapsed_time(b) for a, b in ev)
            alt = {"dtype": "bf16", "value": B * world * args.steps / (ms16 / 1e3),
                   "unit": "samples/s", "ms_per_step": ms16 / args.steps,
                   "fwd_ms_per_launch": f16, "fwd_tflops": FLOPS_PER_STEP / (f16 / 1e3) / 1e12}
            del layer16, x16

        # ---- side measurements that explain the step (not part of `value`): the operand pre-pass
        # alone (CPLXK_DBG=4 returns before the GEMM launch) and the stand-alone KL pass that the
        # fused pre-pass replaces at N = 1
        prep_ms = kl_alone_ms = None

        def device_time(fn, n=20):
            """Mean device time of fn's kernels.  A short kernel timed from Python would measure
            the host's launch latency, so the device is first parked on a ~50 ms busy-wait and
            all n [event, fn, event] groups are queued behind it."""
            fn(); fn(); fn()
            sync_all()
            torch.cuda._sleep(100_000_000)
            evs = []
            for _ in range(n):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                evs.append((a, b))
            sync_all()
            return statistics.mean(a.elapsed_time(b) for a, b in evs)

        if args.dtype == "f32":
            os.environ["CPLXK_DBG"] = "4"
            prep_ms = device_time(lambda: layer(x))
            os.environ.pop("CPLXK_DBG")
            cb.set_kl_fusion(False)
            kl_alone_ms = device_time(lambda: sum(penalties(layer)))
            cb.set_kl_fusion(True)

    f_ms = statistics.mean(a.elapsed_time(b) for a, b in fwd_ms)
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kl_ms)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    esize = 4 if args.dtype != "bf16" else 2
    value = B * world * args.steps / (ms / 1e3)
    e2e_value = B * world * e2e_steps / (e2e_ms / 1e3)
    gemm_ms = f_ms - prep_ms if prep_ms is not None else f_ms
    achieved_tf = FLOPS_PER_STEP / (gemm_ms / 1e3) / 1e12
    peak_tf = peaks["bf16_tflops_sustained"]
    fused = world == 1 and args.dtype == "f32"
    if fused:
        kl_gbs = KL_ALGO_BYTES / (kl_alone_ms / 1e3) / 1e9
        roofline_kl = {
            "kernel": "kl_kernel<CPLX_VD> (stand-alone KL pass)", "bound": "hbm", "achieved": kl_gbs,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kl_gbs / peaks["hbm_gbs"],
            "ms_per_launch": kl_alone_ms,
            "note": "NOT launched inside the timed step at N=1: the forward's operand pre-pass "
                    "evaluates the same per-element penalty on the weight rows it converts and "
                    "penalties() returns that sum (cplxk_linear_vd_fwd_kl); measured here with the "
                    "fusion switched off, CUDA events with the launches queued behind a device busy-wait",
            "ms_in_step": k_ms,
        }
    elif world == 1:
        kl_gbs = KL_ALGO_BYTES * (esize / 4.0) / (k_ms / 1e3) / 1e9
        roofline_kl = {
            "kernel": "kl_kernel<CPLX_VD>", "bound": "hbm", "achieved": kl_gbs,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kl_gbs / peaks["hbm_gbs"],
            "ms_per_launch": k_ms,
            "note": "event pair around the Python-level penalties() call: includes launch latency",
        }
    else:
        roofline_kl = {
            "kernel": "pre-pass by-product on this rank's weight-row shard + NCCL all-reduce of the scalar",
            "note": "the all-reduce waits for the pre-pass event only and runs on a side stream under "
                    "the GEMM (which leaves one SM pair free); not separately timed",
        }
    f32 = args.dtype != "bf16"
    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if f32 else "bf16", "data": "synthetic",
        "config": {
            "workload": "CplxLinearVD 4096->4096 fused local-reparam forward(train) + KL, "
                        "batch=4096 per GPU (BASELINE.json configs[2])",
            "per_gpu_batch": B, "global_batch": B * world, "features": D,
            "parallelism": f"dp{world}" + (" (batch rows sharded, KL row-sharded, 1 scalar all-reduce)"
                                            if world > 1 else ""),
            "storage": "fp32 planes in and out" if f32 else "bf16 planes",
            "math": ("tcgen05 kind::f16 on per-row power-of-two scaled fp16 copies of the fp32 planes "
                     "(11-bit significand = tf32, round-to-nearest; variance GEMM operands bf16), "
                     "fp32 accumulate in TMEM, scales undone in the epilogue") if f32
                    else "tcgen05 kind::f16 (bf16), fp32 accumulate in TMEM",
            "noise": f"in-kernel Philox4x32-10, layout={args.noise}",
            "kl": "fused into the operand pre-pass" if fused else (
                "row shard fused into the operand pre-pass + 1 all-reduce" if world > 1 and f32 else "kl_kernel"),
            "l2": "no flush needed: each step streams 470 MB (fp32) of distinct operands, "
                  "larger than the 126 MB L2",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s",
                "h2d_bytes_per_step": 2 * B * D * esize,
                "d2h_bytes_per_step": 2 * B * D * esize + 4,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "note": "pinned host x -> device, forward+KL, y and KL -> pinned host; copies "
                        "double-buffered on side streams, all inside the timed region; " + placement},
        "gpu_launches": (2 if fused or world > 1 else 3) * args.steps,
        "roofline": {
            "kernel": "fwd_tc3_kernel (persistent CTA-pair: complex mean GEMM + variance GEMM + Philox "
                      "noise + epilogue)" + ("; ms_per_launch = event-timed forward call (inside the timed loop) "
                      "minus the device time of the pre-pass launch" if prep_ms is not None else
                      " timed together with its operand pre-pass"),
            "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf,
            "peak_source": f"{peaks['source']} bf16 cuBLAS sustained (MEASURED_PEAKS.json); kind::f16 "
                           "runs fp16 and bf16 operands at the same rate",
            "algorithmic_flops_per_launch": FLOPS_PER_STEP, "ms_per_launch": gemm_ms,
            "forward_call_ms": f_ms,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed
            # `ncu --set full` capture (profiles/README.md); None until re-captured for this kernel
            "traffic": NCU_TRAFFIC.get("gemm_" + args.dtype),
            "traffic_unit": "bytes/launch",
            "algorithmic_bytes_per_launch": GEMM_ALGO_BYTES if f32 else FWD_ALGO_BYTES * esize / 4.0,
        },
        "roofline_kl": roofline_kl,
    }
    if prep_ms is not None:
        out["roofline_prepass"] = {
            "kernel": "vd_prepare_f16_kernel (fp32 -> row-scaled fp16 operands, |x|^2, exp(log_sigma2), "
                      "row scales, KL sum)", "bound": "hbm",
            "achieved": PREP_ALGO_BYTES / (prep_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": PREP_ALGO_BYTES / (prep_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            "ms_per_launch": prep_ms, "algorithmic_bytes_per_launch": PREP_ALGO_BYTES,
            "traffic": NCU_TRAFFIC.get("prepass_f32"),
        }
    if alt is not None:
        alt["fwd_frac_of_peak"] = alt["fwd_tflops"] / peak_tf
        out["alt_bf16"] = alt
    if world == 1 and not args.no_cpu:
        unbind_all_threads()
        torch.set_num_threads(os.cpu_count() or 1)
        time_oracle(64)
        t = time_oracle(1)
        out["cpu_baseline"] = {
            "value": B / t, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "1 full headline step (B=4096 rows forward + KL over all 4096x4096 weights), "
                      f"fp32 torch CPU + scipy expi, {t:.2f} s",
        }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--noise", default="torch", choices=["torch", "fast"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the in-run CPU baseline leg")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra bf16 measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
