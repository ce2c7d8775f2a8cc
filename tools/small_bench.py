"""BASELINE config 1 (real LinearVD 784 -> 256, batch 128; forward + sum(penalties)): the size this
library's users actually train.  Launch-latency bound -- reports the eager step (host path
included), the CUDA-graph replay of the same step (device time only) and a cProfile of the host
side.  One JSON line."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                                   # noqa: E402
from cplxmodule_b200.nn.relevance import LinearVD, penalties   # noqa: E402


def main():
    torch.manual_seed(0)
    dev = "cuda"
    m = LinearVD(784, 256).to(dev).train()
    x = torch.randn(128, 784, device=dev)
    res = {}
    with torch.no_grad():
        step = lambda: (m(x), sum(penalties(m)))
        for name, n in (("eager_us", 2000),):
            for _ in range(50):
                step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            for _ in range(n):
                step()
            b.record()
            host = time.perf_counter() - t0
            torch.cuda.synchronize()
            res[name] = round(1e3 * a.elapsed_time(b) / n, 2)
            res["eager_host_issue_us"] = round(1e6 * host / n, 2)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = step()
        for _ in range(20):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2000):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        res["graph_replay_us"] = round(1e3 * a.elapsed_time(b) / 2000, 2)
        if os.environ.get("PROFILE", "1") == "1":
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(500):
                step()
            pr.disable()
            torch.cuda.synchronize()
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(18)
            sys.stderr.write(buf.getvalue())
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
