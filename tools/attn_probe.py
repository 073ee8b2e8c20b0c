"""Stand-alone driver of the verify-attention kernel at a named config's shape (ncu captures, quick timing, phase stamps).
PROBE_CFG = c2 (default) | c3 | c4tp8 (one TP-8 rank of 70B) | c4 (unsharded 70B) | c4draft (7B draft, one level)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sequoia_b200 import ops
from sequoia_b200.tree import pack_tree_mask

CFG = {  # H, Hkv, D, M, growmap, P, rows (None = whole tree), n0
    "c2": (32, 32, 128, 384, "A100_growmaps/68m_7b/growmaps/A100-CNN-68m-7b-stochastic.pt", 193, None, 0),
    "c3": (40, 40, 128, 384, "L40_growmaps/8x8-tree.pt", 193, None, 0),
    "c4tp8": (8, 1, 128, 1024, "L40_growmaps/L40-CNN-7b-70b-stochastic.pt", 193, None, 0),
    "c4": (64, 8, 128, 1024, "L40_growmaps/L40-CNN-7b-70b-stochastic.pt", 193, None, 0),
    "c4draft": (32, 32, 128, 1024, "L40_growmaps/L40-CNN-7b-70b-stochastic.pt", 193, 82, 300),
}
name = os.environ.get("PROBE_CFG", "c2")
H, Hkv, D, M, gmp, P, rows, n0 = CFG[name]
L = int(os.environ.get("PROBE_L", "32" if name != "c4" else "8"))
dev = "cuda:0"
gm = torch.load(os.path.join(ROOT, gmp))
S = gm["size"]
n = rows or S
kv_end = (n0 + n) if rows else S
qkv = torch.randn(M, (H + 2 * Hkv) * D, device=dev, dtype=torch.float16)
kc = torch.randn(L, 1, Hkv, M, D, device=dev, dtype=torch.float16)
vc = torch.randn(L, 1, Hkv, M, D, device=dev, dtype=torch.float16)
out = torch.zeros(M, H * D, device=dev, dtype=torch.float16)
plan = ops.AttnPlan(qkv, M, H, Hkv, D, kc, vc, out)
bits = pack_tree_mask(gm["mask"]).to(dev)
state = torch.zeros(16, dtype=torch.int32, device=dev)
state[0] = P
impl = int(os.environ.get("PROBE_IMPL", "0"))


def call(l):
    ops.tree_attn(plan, l % L, n, state=state, n0=n0, kv_end=kv_end, tree_bits=bits, tree_words=bits.shape[1], tree_size=S, impl=impl)


for i in range(8):
    call(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(2 * L):
        call(i)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record(); e1.synchronize()
us = e0.elapsed_time(e1) / (10 * L) * 1e3
kv = P - 1 + kv_end
byt = 2 * D * 2 * (Hkv * kv + H * n)
fl = 4 * H * n * kv * D
print(f"{name}: attention {us:.2f} us/launch  ({byt / us / 1e3:.0f} GB/s algorithmic, {fl / us / 1e6:.1f} TFLOP/s)  H={H} Hkv={Hkv} q={n} kv={kv} "
      f"plan error {plan.error()}")

if os.environ.get("SQ_ATTN_TIMING"):
    import ctypes, numpy as np
    from sequoia_b200 import _lib
    call(0); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 128)()
    rc = _lib.load().sq_attn_plan_debug_times(plan.handle, buf)
    t = np.array(list(buf)).reshape(8, 16)
    nm = {0: "P known", 1: "prologue sync", 2: "S0 ready", 4: "softmax0 done", 5: "O done", 10: "O staged+dealloc", 6: "pushed",
          7: "cluster sync", 9: "weights", 8: "end"}
    for s_ in range(8):
        row = t[s_]
        if row[0] == 0:
            continue
        print("split", s_, " ".join(f"{nm[k]}:+{int(row[k] - row[0])}" for k in (0, 1, 2, 4, 5, 10, 6, 7, 9, 8) if row[k]))
