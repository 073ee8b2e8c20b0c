"""Correctness + timing of the weight-streaming GEMM (csrc/sq_gemm.cu) vs cuBLASLt (torch.mm) at the 7B layer shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sequoia_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
n = int(os.environ.get("PROBE_N", "128"))
shapes = {"qkv": (12288, 4096), "o": (4096, 4096), "gate_up": (22016, 4096), "down": (4096, 11008), "lm_head": (32000, 4096)}
only = os.environ.get("PROBE_ONLY")
COPIES = 6
peak = 6568.7
err = torch.zeros(4, dtype=torch.int32, device=dev)
for name, (N, K) in shapes.items():
    if only and name not in only.split(","):
        continue
    a = (torch.randn(128, K, device=dev) * 0.5).half()
    ws = [(torch.randn(N, K, device=dev) * 0.05).half() for _ in range(COPIES)]
    c = torch.zeros(128, N, device=dev, dtype=torch.float16)
    plans = [ops.GemmPlan(a, w, c, err, tiled=os.environ.get('PROBE_TILED', '0') == '1') for w in ws]
    plans[0].run(n)
    torch.cuda.synchronize()
    ref = (a[:n].float() @ ws[0].float().t())
    got = c[:n].float()
    cub = torch.mm(a[:n], ws[0].t()).float()
    e_mine = ((got - ref).abs().max() / ref.abs().max()).item()
    e_cub = ((cub - ref).abs().max() / ref.abs().max()).item()
    def timeit(fn):
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for i in range(COPIES): fn(i)
            s.synchronize()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(4 * COPIES): fn(i % COPIES)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): g.replay()
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1) / (12 * COPIES) * 1e3
    t_mine = timeit(lambda i: plans[i].run(n))
    t_cub = timeit(lambda i: torch.mm(a[:n], ws[i].t(), out=c[:n]))
    gb = N * K * 2 / 1e9
    print(f"{name:8s} N={N:6d} K={K:6d} plan(bn,split,stages)={plans[0].info()} relerr mine {e_mine:.2e} cublas {e_cub:.2e} | "
          f"mine {t_mine:7.2f} us ({gb / t_mine * 1e6:6.0f} GB/s, {gb / t_mine * 1e6 / peak * 100:4.1f}%)  cublas {t_cub:7.2f} us "
          f"({gb / t_cub * 1e6:6.0f} GB/s)  watchdog {err.tolist()}", flush=True)
    del ws, plans
