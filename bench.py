#!/usr/bin/env python
"""Headline benchmark: CplxLinearVD forward + KL, samples/s at B=4096, d=4096 per GPU
(BASELINE.json `metric`, configs[2]); weak scaling over N GPUs (batch rows sharded, KL
row-sharded + ONE scalar all-reduce).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3|5]

N > 1 is launched by torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.
A "step" = one training-mode forward of the layer on one batch (fused local
reparameterisation, in-kernel Philox noise) + sum(penalties(model)).

Legs of the default run (rank 0 prints them all in the one line):
  value / roofline   device-timed steps, inputs resident in HBM (W warm-up + K timed steps: the
                     burst regime, compared with the burst cuBLAS peak); roofline.sustained = the
                     same K steps after > 0.4 s under load, against the sustained cuBLAS peak
  e2e                the same step from pinned host buffers and back, every step, + the box's
                     measured duplex copy rate for the same bytes (e2e.roofline)
  parity             64 sampled rows of the timed layer's output + its KL against the oracle
  cpu_baseline       one full step of the reference's own CPU path on the host cores
  torch_eager_gpu    the UNMODIFIED reference executed on this GPU through torch eager
  extra.config5      CplxLinearARD 8192x8192 on 8192 rows per GPU (BASELINE.json configs[4])
  extra.config4      CplxConv2d 64->64 3x3 on 256x64x128x128, fp32 NCHW and bf16 channels-last (configs[3])
`--impl reference` times the unmodified reference (baseline/_ref) on the host cores, every step
the FULL headline step.
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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

METRIC = "CplxLinearVD forward+KL samples/sec (B=4096, d=4096)"

# workloads: BASELINE.json configs[2] (headline) and configs[4] (one GPU's shard)
WORKLOADS = {
    3: {"name": "CplxLinearVD 4096->4096 fused local-reparam forward(train) + KL, batch=4096 per GPU "
                "(BASELINE.json configs[2])", "layer": "CplxLinearVD", "B": 4096, "D": 4096, "kl": "cplx_vd"},
    5: {"name": "CplxLinearARD 8192->8192 fused local-reparam forward(train) + KL, batch=8192 per GPU "
                "(BASELINE.json configs[4]: global batch 65536 on 8 GPUs)", "layer": "CplxLinearARD",
        "B": 8192, "D": 8192, "kl": "cplx_ard"},
}


def flops_per_step(B, D):
    return 10.0 * B * D * D          # 8 (complex mean GEMM, 4-multiply form) + 2 (variance GEMM)


def prep_algo_bytes(B, D):
    # operand pre-pass (fp32 planes): reads x (2 planes), W (2 planes), log_sigma2; writes 3 + 3
    # planes of 16-bit operands (fp16 re/im + bf16 |x|^2, fp16 U/V + bf16 exp(log_sigma2)) + scales
    return 4.0 * (2 * B * D + 3 * D * D) + 2.0 * (3 * B * D + 3 * D * D) + 4.0 * (B + D)


def gemm_algo_bytes(B, D):
    # the GEMM kernel streams those six 16-bit planes and writes y (fp32)
    return 2.0 * (3 * B * D + 3 * D * D) + 4.0 * 2 * B * D


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
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# dram bytes (read + write) per launch from the committed ncu --set full captures (profiles/);
# the values are read from profiles/traffic.json (written by tools/ncu_traffic.py from the raw
# csv of the capture) so that the number in the line is the one of the committed capture
def ncu_traffic():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


# ----------------------------------------------------------------------------- CPU side
def import_reference():
    """The UNMODIFIED reference package installed under baseline/_ref (git-ignored, travels to
    the GPU box with the working tree).  Returns the module or None."""
    if not os.path.isdir(os.path.join(REF_DIR, "cplxmodule")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import cplxmodule  # noqa: F401
        return cplxmodule
    except Exception:  # noqa: BLE001
        return None


def make_reference_step(wl, device="cpu"):
    """(step_fn, kind, describe): one full step of the workload on `device` through the
    reference's own public API (`kind` "reference"), or -- if baseline/_ref cannot be imported --
    through the oracle's restatement of it (`kind` "port")."""
    import torch
    B, D = wl["B"], wl["D"]
    ref = import_reference()
    torch.manual_seed(0)
    if ref is not None:
        from cplxmodule import cplx as rcplx
        from cplxmodule.nn import relevance as rrel
        layer = getattr(rrel, wl["layer"])(D, D).to(device).train()
        z = rcplx.randn(B, D).to(device)

        def step():
            with torch.no_grad():
                y = layer(z)
                kl = sum(rrel.penalties(layer))
            return y, kl
        return step, "reference", f"cplxmodule {ref.__version__} (baseline/_ref), torch {torch.__version__}"
    from oracle import cplx_oracle as orc
    bound = 1.0 / (2 * D) ** 0.5
    w_re = torch.empty(D, D, device=device).uniform_(-bound, bound)
    w_im = torch.empty(D, D, device=device).uniform_(-bound, bound)
    b_re = torch.empty(D, device=device).uniform_(-1 / D ** 0.5, 1 / D ** 0.5)
    b_im = torch.empty(D, device=device).uniform_(-1 / D ** 0.5, 1 / D ** 0.5)
    ls2 = torch.full((D, D), -10.0, device=device)
    x_re, x_im = (torch.randn(B, D, device=device) / 2 ** 0.5 for _ in range(2))

    def step():
        with torch.no_grad():
            er, ei = orc.cplx_randn(B, D)
            y = orc.cplx_linear_vd(x_re, x_im, w_re, w_im, b_re, b_im, ls2, er.to(device), ei.to(device))
            kl = orc.layer_penalty(wl["kl"], w_re, w_im, ls2, "sum")
        return y, kl
    return step, "port", f"oracle/cplx_oracle.py restatement, torch {torch.__version__}"


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores;
    every step is the FULL workload (no sub-sampling, no extrapolation)."""
    if rank != 0:
        return
    import torch
    wl = WORKLOADS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, what = make_reference_step(wl)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = wl["B"] * args.steps / dt
    sample = (f"every step = the full workload: {wl['B']} rows through the {wl['D']}x{wl['D']} "
              f"training-mode forward + KL over all {wl['D']}x{wl['D']} weights, fp32, {what}")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "per_gpu_batch": wl["B"],
                   "global_batch": wl["B"] * args.gpus, "features": wl["D"],
                   "parallelism": f"cpu ({cores} threads), rank 0 only"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind,
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


def numa_nodes():
    nodes = {}
    base = "/sys/devices/system/node"
    try:
        names = [n for n in os.listdir(base) if n.startswith("node") and n[4:].isdigit()]
    except OSError:
        return nodes
    for n in names:
        try:
            with open(f"{base}/{n}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    if not part:
                        continue
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            nodes[int(n[4:])] = cpus
        except (OSError, ValueError):
            continue
    return nodes


def place_near_gpu(local_rank, dev):
    """Bind this process (and so the first touch of its pinned staging buffers) to the NUMA node
    from which host<->device copies are fastest.  sysfs' numa_node is used when it names a node;
    when it does not (-1, common in VMs) or there are several candidates, every node is PROBED
    with a short pinned H2D copy and the fastest wins -- ranks are never all piled on node 0 by
    default.  Placement only; best effort."""
    import torch
    allowed0 = os.sched_getaffinity(0)
    nodes = {k: v & allowed0 for k, v in numa_nodes().items() if v & allowed0}
    if len(nodes) <= 1:
        return "numa: single node"
    hinted = None
    try:
        p = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
        with open(path) as f:
            hinted = int(f.read().strip())
    except Exception:  # noqa: BLE001
        hinted = None
    _ORIG_AFFINITY.append(allowed0)
    if hinted is not None and hinted in nodes:
        os.sched_setaffinity(0, nodes[hinted])
        return f"numa: bound to node {hinted} (sysfs, {len(nodes[hinted])} cpus)"
    best, best_gbs, rates = None, 0.0, {}
    dst = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
    for node, cpus in sorted(nodes.items()):
        try:
            os.sched_setaffinity(0, cpus)
            src = torch.empty(32 << 20, dtype=torch.uint8).pin_memory()
            src.fill_(1)                              # first touch on this node
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            b.record()
            torch.cuda.synchronize(dev)
            gbs = 4 * src.numel() / (a.elapsed_time(b) / 1e3) / 1e9
            rates[node] = round(gbs, 1)
            if gbs > best_gbs:
                best, best_gbs = node, gbs
            del src
        except Exception:  # noqa: BLE001
            continue
    if best is None:
        os.sched_setaffinity(0, allowed0)
        return "numa: probe failed, not bound"
    os.sched_setaffinity(0, nodes[best])
    return f"numa: bound to node {best} (probed H2D GB/s per node: {rates})"


# ----------------------------------------------------------------------------- GPU side
class Harness:
    """Timing helpers shared by the legs (CUDA events on the current stream, max over ranks)."""

    def __init__(self, dev, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.dev, self.world = torch, dist, dev, world

    def sync_all(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps):
        torch = self.torch
        self.sync_all()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        t.record()
        self.sync_all()
        ms = torch.tensor([s.elapsed_time(t)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def device_time(self, fn, n=20):
        """Mean device time of fn's kernels.  A short kernel timed from Python would measure the
        host's launch latency, so the device is first parked on a ~50 ms busy-wait and all n
        [event, fn, event] groups are queued behind it."""
        torch = self.torch
        fn(); fn(); fn()
        self.sync_all()
        torch.cuda._sleep(100_000_000)
        evs = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        self.sync_all()
        return statistics.mean(a.elapsed_time(b) for a, b in evs)


def parity_block(layer, x, wl, world, kl_step):
    """Parity of the very layer that was timed: 64 sampled rows of a forward with INJECTED noise
    against the float64 oracle, the KL against the float64 closed form, and -- N > 1 -- the
    all-reduced row-sharded KL against the unsharded stand-alone kernel on this rank."""
    import torch
    import cplxmodule_b200 as cb
    from cplxmodule_b200 import cplx
    from cplxmodule_b200.nn.relevance import penalties
    from oracle import cplx_oracle as orc
    B, D = wl["B"], wl["D"]
    dev = x.real.device
    g = torch.Generator(device=dev).manual_seed(12345)
    rows = torch.randperm(B, generator=g, device=dev)[:64]
    xs = cplx.Cplx(x.real[rows].contiguous(), x.imag[rows].contiguous())
    # same kernel as the timed step needs M > 128 rows: run the full batch with injected noise
    eps = cplx.Cplx(torch.randn(B, D, device=dev, generator=g) / 2 ** 0.5,
                    torch.randn(B, D, device=dev, generator=g) / 2 ** 0.5)
    with torch.no_grad():
        y = layer(x, eps=eps)
        kl_fused = sum(penalties(layer)) if world == 1 else None
        cb.set_kl_fusion(False)
        shard = cb.ops._state["kl_shard"]
        cb.set_kl_shard()
        kl_alone = sum(penalties(layer))                      # unsharded kl_kernel
        cb.set_kl_shard(*shard) if shard else None
        cb.set_kl_fusion(True)
    c = lambda t: t.detach().float().cpu().double()
    w, b = layer.weight, layer.bias
    ls2 = c(layer.log_sigma2)
    want = orc.cplx_linear_vd(c(xs.real), c(xs.imag), c(w.real), c(w.imag), c(b.real), c(b.imag),
                              ls2, c(eps.real[rows]), c(eps.imag[rows]))
    scale = max(want[0].abs().max().item(), want[1].abs().max().item())
    err = max((c(y.real[rows]) - want[0]).abs().max().item(),
              (c(y.imag[rows]) - want[1]).abs().max().item())
    la = orc.log_alpha_cplx(c(w.real), c(w.imag), ls2)
    kl_ref = (orc.penalty_cplx_vd_exact64(la) if wl["kl"] == "cplx_vd"
              else orc.penalty_cplx_ard(la)).sum().item()
    out = {"fwd_rel_err": err / scale, "fwd_tol": 1e-3 if y.real.dtype == torch.float32 else 1e-2,
           "rows_checked": 64, "kl_rel_err": abs(kl_alone.item() - kl_ref) / abs(kl_ref),
           "kl_tol": 1e-3, "against": "float64 oracle (oracle/cplx_oracle.py), injected noise"}
    if kl_fused is not None:
        out["kl_fused_vs_standalone"] = abs(kl_fused.item() - kl_alone.item()) / abs(kl_ref)
    if kl_step is not None:                                    # N > 1: the value the timed step produced
        rel = abs(float(kl_step) - kl_alone.item()) / abs(kl_alone.item())
        out["kl_allreduced_vs_unsharded"] = rel
        out["kl_allreduce_ok"] = bool(rel < 1e-6)
    # reported, never fatal: a bench line with "ok": false is still a line the driver can read
    out["ok"] = bool(out["fwd_rel_err"] < out["fwd_tol"] and out["kl_rel_err"] < out["kl_tol"]
                     and out.get("kl_allreduce_ok", True))
    return out


def torch_eager_leg(wl, dev):
    """The UNMODIFIED reference executed on this GPU through torch eager (5 cuBLAS GEMMs + ~16
    elementwise kernels + host scipy Ei for the KL): the implementation to beat on the same box."""
    import torch
    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            step, kind, what = make_reference_step(wl, device=dev)
            step()
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                step()
                torch.cuda.synchronize(dev)
                ts.append(time.perf_counter() - t0)
            key = "tf32" if tf32 else "fp32"
            out[key] = {"ms_per_step": 1e3 * min(ts), "value": wl["B"] / min(ts), "unit": "samples/s"}
            out["kind"], out["what"] = kind, what + ", device=cuda, wall clock around a synchronised step (min of 2)"
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def copy_roofline(dev, world, h, host_in, host_out, dev_in, dev_out, steps):
    """Pinned host->device of one step's inputs and device->host of one step's outputs, issued
    concurrently on two streams with NO compute: the box's duplex copy rate for exactly the bytes
    the e2e step moves (all N ranks at once)."""
    import torch
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def loop():
        for _ in range(steps):
            with torch.cuda.stream(s_in):
                for d, s in zip(dev_in, host_in):
                    d.copy_(s, non_blocking=True)
            with torch.cuda.stream(s_out):
                for d, s in zip(host_out, dev_out):
                    d.copy_(s, non_blocking=True)
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(s_in)
        cur.wait_stream(s_out)

    loop()
    return h.timed(loop, 1) / steps


def run_workload(args, wl, h, rank, local_rank, world, placement, full):
    """Device-timed steps (+ e2e, parity, side measurements when `full`).  Returns a dict."""
    import torch
    import torch.distributed as dist
    import cplxmodule_b200 as cb
    from cplxmodule_b200 import _native as nv
    from cplxmodule_b200 import cplx
    from cplxmodule_b200.distributed import sharded_penalties
    from cplxmodule_b200.nn import relevance

    dev = h.dev
    B, D = wl["B"], wl["D"]
    torch.manual_seed(0)                               # identical (replicated) parameters
    layer = getattr(relevance, wl["layer"])(D, D).to(dev).train()
    if args.dtype == "bf16":
        layer = layer.bfloat16()
    dt = torch.float32 if args.dtype != "bf16" else torch.bfloat16
    torch.manual_seed(1000 + rank)                     # per-rank batch
    host_x = [torch.randn(B, D).div_(2 ** 0.5).to(dt).pin_memory() for _ in range(2)]
    x = cplx.Cplx(host_x[0].to(dev), host_x[1].to(dev))

    def kl_term():
        if world > 1:
            return sharded_penalties(layer)[1].sum()
        return sum(relevance.penalties(layer))

    fwd_ms, kl_ms = [], []
    # high priority: when the pre-pass retires, the collective's CTA is placed before the
    # persistent GEMM grid takes every SM
    kl_stream = torch.cuda.Stream(dev, priority=-1) if world > 1 else None
    last = {}

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
        last["y"], last["kl"] = y, kl
        return y, kl

    res = {"B": B, "D": D}
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step()
        sampler = ClockSampler(local_rank)
        if rank == 0 and full:
            sampler.start()
            time.sleep(0.3)
        # Every 4th step of the timed loop carries the event pairs of the per-call breakdown
        # (forward / penalties): an event record between two launches breaks their programmatic-
        # dependent-launch adjacency (GEMM -> KL guard), so the other steps run bare.
        counter = {"i": 0}

        def timed_step():
            counter["i"] += 1
            step(record=(counter["i"] & 3) == 0)

        ms = h.timed(timed_step, args.steps)
        res["ms"] = ms
        if not fwd_ms:                        # fewer than 4 timed steps: one instrumented step behind them
            step(record=True)
            torch.cuda.synchronize(dev)
        res["f_ms"] = statistics.mean(a.elapsed_time(b) for a, b in fwd_ms)
        res["k_ms"] = statistics.mean(a.elapsed_time(b) for a, b in kl_ms)
        res["sustained_ms"] = None
        if full:
            # the timed loop may be shorter than nvidia-smi's sampling period: keep the identical
            # loop running (untimed) until ~0.4 s of load has been sampled (`ms` is the all-reduced
            # maximum, so every rank runs the same number of extra steps)
            extra = int(max(0.0, 400.0 - ms) / max(ms / args.steps, 1e-3)) + 1
            for _ in range(min(extra, 5000)):
                step()
            # ... and the same K steps once more, now that the part has been under load for > 0.4 s
            # (the 1 kW power cap has pulled the SM clock down by then): the SUSTAINED figure next
            # to the contract's `value` (W warm-up steps, then K timed steps)
            res["sustained_ms"] = h.timed(step, args.steps)
            torch.cuda.synchronize(dev)
            clocks = sampler.stop() if rank == 0 else None
            if clocks is not None:
                clocks["window"] = "timed loop + identical continuation (>= 0.4 s under load) + sustained loop"
            res["clocks"] = clocks
        res["kl_value"] = float(last["kl"].item())

        # ---- side measurements that explain the step (not part of `value`): the operand pre-pass
        # alone (cplxk_linear_vd_prepare: the same launch the forward starts with) and the
        # stand-alone KL pass that the fused pre-pass replaces at N = 1
        res["prep_ms"] = res["kl_alone_ms"] = None
        if args.dtype == "f32":
            lib = nv.lib()
            code = nv.dtype_code(dt)
            ws_bytes = lib.cplxk_linear_vd_workspace_bytes(B, D, D, code)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            kl_sum = torch.empty((), dtype=torch.float32, device=dev)
            kl_ws = nv.kl_workspace(dev)
            w = layer.weight
            kind = layer._kl_kind

            def prep_only():
                # exactly what the forward call launches in front of the GEMM kernel
                nv.check(lib.cplxk_linear_vd_prepare(
                    nv.ptr(x.real), nv.ptr(x.imag), nv.ptr(w.real), nv.ptr(w.imag),
                    nv.ptr(layer.log_sigma2), B, D, D, code, nv.ptr(ws), ws_bytes,
                    kind, nv.ptr(kl_sum), nv.ptr(kl_ws), kl_ws.numel() * 8,
                    nv.stream_ptr(dev)))

            res["prep_ms"] = h.device_time(prep_only)
            del ws
            cb.set_kl_fusion(False)
            shard = cb.ops._state["kl_shard"]
            cb.set_kl_shard()
            res["kl_alone_ms"] = h.device_time(lambda: sum(relevance.penalties(layer)))
            if shard:
                cb.set_kl_shard(*shard)
            cb.set_kl_fusion(True)

        if not full:
            res["parity"] = parity_block(layer, x, wl, world, res["kl_value"] if world > 1 else None) \
                if rank == 0 else None
            return res

        # ---- end to end: host (pinned) inputs in, host outputs back, EVERY step.
        # Double-buffered: the H2D copy of step i+1 and the D2H copy of step i-1 run on their
        # own streams while step i computes (PCIe is full duplex); every byte still moves
        # inside the timed region.
        host_y = [torch.empty(B, D, dtype=dt).pin_memory() for _ in range(2)]
        host_kl = torch.empty((), dtype=torch.float32).pin_memory()
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
        res["e2e_steps"] = e2e_steps
        res["e2e_ms"] = h.timed(lambda: e2e_loop(e2e_steps), 1)
        y_dev = [last["y"].real, last["y"].imag]
        res["copy_ms_per_step"] = copy_roofline(dev, world, h, host_x, host_y, [xbuf[0].real, xbuf[0].imag],
                                                y_dev, e2e_steps)
        del host_y, xbuf

        # ---- same step with bf16 planes (BASELINE.json configs[1] precision class), reported
        # beside the fp32 headline, never instead of it
        res["alt"] = None
        if args.dtype == "f32" and not args.no_alt:
            layer16 = getattr(relevance, wl["layer"])(D, D).to(dev).train().bfloat16()
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
                    sum(relevance.penalties(layer16))

            for _ in range(3):
                step16()
            ev.clear()
            ms16 = h.timed(step16, args.steps)
            f16 = statistics.mean(a.elapsed_time(b) for a, b in ev)
            res["alt"] = {"dtype": "bf16", "value": B * world * args.steps / (ms16 / 1e3),
                          "unit": "samples/s", "ms_per_step": ms16 / args.steps,
                          "fwd_ms_per_launch": f16, "fwd_tflops": flops_per_step(B, D) / (f16 / 1e3) / 1e12}
            del layer16, x16

        res["parity"] = parity_block(layer, x, wl, world, res["kl_value"] if world > 1 else None) \
            if rank == 0 else None
    return res


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import cplxmodule_b200 as cb

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    placement = place_near_gpu(local_rank, dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cb.set_noise_mode(args.noise)
    if world > 1:
        # N > 1: the pre-pass by-product covers this rank's block of weight rows; ONE all-reduce
        # combines the partial sums on a side stream while the GEMM runs
        cb.set_kl_shard(rank, world)
        cb.set_sm_reserve(int(os.environ.get("BENCH_SM_RESERVE", "0")))
    h = Harness(dev, world)
    wl = WORKLOADS[args.config]
    B, D = wl["B"], wl["D"]
    res = run_workload(args, wl, h, rank, local_rank, world, placement, full=True)

    extra = {}
    if args.config == 3 and not args.no_extra and args.dtype == "f32":
        # BASELINE.json configs[4]: one GPU's shard of CplxLinearARD 8192^2, global batch 8192 * N
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup = max(5, min(args.steps, 10)), 3
        torch.cuda.empty_cache()
        r5 = run_workload(sub, WORKLOADS[5], h, rank, local_rank, world, placement, full=False)
        w5 = WORKLOADS[5]
        g5 = r5["f_ms"] - r5["prep_ms"] if r5["prep_ms"] is not None else r5["f_ms"]
        extra["config5"] = {
            "workload": w5["name"], "value": w5["B"] * world * sub.steps / (r5["ms"] / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": sub.steps, "ms_per_step": r5["ms"] / sub.steps,
            "global_batch": w5["B"] * world, "step_tflops": flops_per_step(w5["B"], w5["D"]) / (r5["ms"] / sub.steps / 1e3) / 1e12,
            "gemm_ms": g5, "prepass_ms": r5["prep_ms"], "parity": r5["parity"],
        }
        torch.cuda.empty_cache()

    if args.config == 3 and not args.no_extra and args.dtype == "f32" and world == 1:
        # BASELINE.json configs[3]: CplxConv2d 64 -> 64, 3x3, 256 x 64 x 128 x 128 on one GPU
        try:
            extra["config4"] = conv_config4_leg(dev)
        except Exception as exc:  # noqa: BLE001 -- an extra line, never the headline
            extra["config4"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.empty_cache()

    eager = None
    if world == 1 and not args.no_eager and rank == 0:
        try:
            eager = torch_eager_leg(wl, dev)
        except Exception as exc:  # noqa: BLE001 -- a reported baseline, never the product path
            eager = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    traffic = ncu_traffic()
    esize = 4 if args.dtype != "bf16" else 2
    ms, f_ms, k_ms, prep_ms, kl_alone_ms = res["ms"], res["f_ms"], res["k_ms"], res["prep_ms"], res["kl_alone_ms"]
    e2e_steps, e2e_ms = res["e2e_steps"], res["e2e_ms"]
    value = B * world * args.steps / (ms / 1e3)
    e2e_value = B * world * e2e_steps / (e2e_ms / 1e3)
    copy_value = B * world / (res["copy_ms_per_step"] / 1e3)
    gemm_ms = f_ms - prep_ms if prep_ms is not None else f_ms
    FLOPS = flops_per_step(B, D)
    achieved_tf = FLOPS / (gemm_ms / 1e3) / 1e12
    step_tf = FLOPS / (ms / args.steps / 1e3) / 1e12
    peak_tf = peaks["bf16_tflops"]                 # burst: the regime of W warm-up + K timed steps
    peak_sus = peaks["bf16_tflops_sustained"]
    fused = world == 1 and args.dtype == "f32"
    KL_BYTES = 4.0 * 3 * D * D
    if fused:
        kl_gbs = KL_BYTES / (kl_alone_ms / 1e3) / 1e9
        roofline_kl = {
            "kernel": "kl_kernel (stand-alone KL pass)", "bound": "hbm", "achieved": kl_gbs,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kl_gbs / peaks["hbm_gbs"],
            "ms_per_launch": kl_alone_ms,
            "note": "NOT launched inside the timed step at N=1: the forward's operand pre-pass "
                    "evaluates the same per-element penalty on the weight rows it converts and "
                    "penalties() returns that sum (cplxk_linear_vd_fwd_kl); measured here with the "
                    "fusion switched off, CUDA events with the launches queued behind a device busy-wait",
            "ms_in_step": k_ms,
        }
    elif world == 1:
        kl_gbs = KL_BYTES * (esize / 4.0) / (k_ms / 1e3) / 1e9
        roofline_kl = {
            "kernel": "kl_kernel", "bound": "hbm", "achieved": kl_gbs,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kl_gbs / peaks["hbm_gbs"],
            "ms_per_launch": k_ms,
            "note": "event pair around the Python-level penalties() call: includes launch latency",
        }
    else:
        roofline_kl = {
            "kernel": "pre-pass by-product on this rank's weight-row shard + NCCL all-reduce of the scalar",
            "note": "the all-reduce waits for the pre-pass event only and runs on a side stream under "
                    "the GEMM; not separately timed",
        }
    f32 = args.dtype != "bf16"
    h2d_b, d2h_b = 2 * B * D * esize, 2 * B * D * esize + 4
    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if f32 else "bf16", "data": "synthetic",
        "config": {
            "workload": wl["name"],
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
            "l2": f"no flush needed: each step streams {(4 * B * D + 3 * D * D) * esize / 1e6:.0f} MB of "
                  "distinct operands and outputs, larger than the 126 MB L2",
        },
        "clocks": res["clocks"],
        "e2e": {"value": e2e_value, "unit": "samples/s",
                "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "roofline": {
                    "bound": "pcie", "what": "the same pinned buffers copied H2D and D2H concurrently on two "
                    "streams with no compute, all ranks at once (measured in this run)",
                    "peak": copy_value, "unit": "samples/s", "copy_ms_per_step": res["copy_ms_per_step"],
                    "h2d_gbs_per_gpu": h2d_b / (res["copy_ms_per_step"] / 1e3) / 1e9,
                    "d2h_gbs_per_gpu": d2h_b / (res["copy_ms_per_step"] / 1e3) / 1e9,
                    "frac": e2e_value / copy_value},
                "note": "pinned host x -> device, forward+KL, y and KL -> pinned host; copies "
                        "double-buffered on side streams, all inside the timed region; " + placement},
        # our kernels per step: operand pre-pass (+ KL sum + parameter fingerprint), persistent GEMM,
        # KL guard -- or, bf16 planes: |x|^2 / exp pre-pass, GEMM, kl_kernel
        "gpu_launches": 3 * args.steps,
        "parity": res["parity"],
        "roofline": {
            "kernel": "fwd_tc3_kernel (persistent CTA-pair: complex mean GEMM + variance GEMM + Philox "
                      "noise + epilogue)" + ("; ms_per_launch = event-timed forward call (inside the timed loop) "
                      "minus the device time of the pre-pass launch" if prep_ms is not None else
                      " timed together with its operand pre-pass"),
            "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf,
            "step_achieved": step_tf, "step_frac": step_tf / peak_tf,
            "peak_source": f"{peaks['source']} bf16 cuBLAS BURST peak (MEASURED_PEAKS.json: best of 10) -- the "
                           "timed loop is W warm-up + K steps, i.e. tens of milliseconds, before the 1 kW "
                           "power cap settles the clocks; `sustained` below is the same loop after > 0.4 s "
                           "under load against the SUSTAINED cuBLAS peak; kind::f16 runs fp16 and bf16 "
                           "operands at the same rate",
            "frac_of_sustained_peak": achieved_tf / peak_sus,
            "sustained": None if res.get("sustained_ms") is None else {
                "ms_per_step": res["sustained_ms"] / args.steps,
                "value": B * world * args.steps / (res["sustained_ms"] / 1e3), "unit": "samples/s",
                "step_achieved": FLOPS / (res["sustained_ms"] / args.steps / 1e3) / 1e12,
                "peak": peak_sus,
                "step_frac": FLOPS / (res["sustained_ms"] / args.steps / 1e3) / 1e12 / peak_sus},
            "algorithmic_flops_per_launch": FLOPS, "ms_per_launch": gemm_ms,
            "forward_call_ms": f_ms,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed
            # `ncu --set full` capture (profiles/traffic.json); None until captured for this kernel/shape
            "traffic": traffic.get(f"gemm_{args.dtype}_{D}"),
            "traffic_unit": "bytes/launch",
            "algorithmic_bytes_per_launch": gemm_algo_bytes(B, D) if f32 else (4 * B * D + 3 * D * D) * 2.0,
        },
        "roofline_kl": roofline_kl,
    }
    if prep_ms is not None:
        pb = prep_algo_bytes(B, D)
        out["roofline_prepass"] = {
            "kernel": "vd_prepare_f16_kernel (fp32 -> row-scaled fp16 operands, |x|^2, exp(log_sigma2), "
                      "row scales, KL sum)", "bound": "hbm",
            "achieved": pb / (prep_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": pb / (prep_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            "ms_per_launch": prep_ms, "algorithmic_bytes_per_launch": pb,
            "traffic": traffic.get(f"prepass_f32_{D}"),
        }
    if res.get("alt") is not None:
        res["alt"]["fwd_frac_of_peak"] = res["alt"]["fwd_tflops"] / peak_tf
        out["alt_bf16"] = res["alt"]
    if extra:
        out["extra"] = extra
    if eager is not None:
        out["torch_eager_gpu"] = eager
        if "fp32" in eager:
            out["vs_torch_eager"] = {"fp32": value / eager["fp32"]["value"],
                                     "tf32": value / eager["tf32"]["value"]}
    if world == 1 and not args.no_cpu:
        unbind_all_threads()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        step_cpu, kind, what = make_reference_step(wl)
        t0 = time.perf_counter()
        step_cpu()
        t = time.perf_counter() - t0
        out["cpu_baseline"] = {
            "value": B / t, "unit": "samples/s", "cores": cores, "kind": kind,
            "sample": f"1 full step ({B} rows forward + KL over all {D}x{D} weights), fp32, {what}, {t:.2f} s",
        }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def conv_config4_leg(dev, iters=10, warm=3):
    """BASELINE.json configs[3] (CplxConv2d 64 -> 64 ch, 3x3, 128x128 input, batch 256): fp32 NCHW planes
    (torch's default layout; per-image scaled fp16 operands) and bf16 channels-last, CUDA events, warm;
    image 0 of the fp32 result against the float64 oracle.  1.199 TFLOP, 4.23 GB (fp32) per call."""
    import torch
    from cplxmodule_b200 import cplx as cx
    from cplxmodule_b200.nn import CplxConv2d
    from oracle import cplx_oracle as orc
    torch.manual_seed(4)
    flops = 8.0 * 256 * 64 * 126 * 126 * 64 * 9
    out = {"workload": "CplxConv2d 64->64 3x3 on 256x64x128x128 (BASELINE.json configs[3])",
           "algorithmic_tflop": flops / 1e12}
    with torch.no_grad():
        conv = CplxConv2d(64, 64, 3).to(dev)
        z = cx.randn(256, 64, 128, 128, device=dev)
        for tag, c, inp in (("fp32_nchw", conv, z), ("bf16_channels_last", None, None)):
            if c is None:
                c = conv.to(torch.bfloat16)
                inp = cx.Cplx(z.real.to(torch.bfloat16).contiguous(memory_format=torch.channels_last),
                              z.imag.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
            for _ in range(warm):
                y = c(inp)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                y = c(inp)
            b.record()
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / iters
            out[tag] = {"ms": ms, "tflops": flops / (ms / 1e3) / 1e12, "images_per_s": 256 / (ms / 1e3)}
            if tag == "fp32_nchw":
                cpu = lambda t: t.detach().double().cpu()
                want = orc.cplx_conv2d(cpu(z.real[:1]), cpu(z.imag[:1]), cpu(conv.weight.real), cpu(conv.weight.imag),
                                       cpu(conv.bias.real), cpu(conv.bias.imag))
                err = max(float((cpu(y.real[:1]) - want[0]).abs().max() / want[0].abs().max()),
                          float((cpu(y.imag[:1]) - want[1]).abs().max() / want[1].abs().max()))
                out[tag]["rel_err_image0_vs_oracle"] = err
                out[tag]["tol"] = 1e-3
            del y
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(WORKLOADS),
                    help="BASELINE.json config: 3 = headline CplxLinearVD 4096^2, 5 = CplxLinearARD 8192^2 shard")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--noise", default="torch", choices=["torch", "fast"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the in-run CPU baseline leg")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra bf16 measurement")
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager-on-GPU reference leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the config-5 line")
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
