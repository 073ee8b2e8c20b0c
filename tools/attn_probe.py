"""Stand-alone driver of the verify-attention kernel at config-2 shape (for ncu captures and quick timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sequoia_b200 import ops
from sequoia_b200.tree import pack_tree_mask

H, Hkv, D, M, L = 32, 32, 128, 384, int(os.environ.get("PROBE_L", "32"))
dev = "cuda:0"
gm = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                             "A100_growmaps/68m_7b/growmaps/A100-CNN-68m-7b-stochastic.pt"))
S = gm["size"]
P = 193
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
    ops.tree_attn(plan, l % L, S, state=state, n0=0, kv_end=S, tree_bits=bits, tree_words=bits.shape[1], tree_size=S, impl=impl)

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
print("attention us/launch:", e0.elapsed_time(e1) / (10 * L) * 1e3, "plan error", plan.error())

if os.environ.get("SQ_ATTN_TIMING"):
    import ctypes, numpy as np
    from sequoia_b200 import _lib
    call(0); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 128)()
    rc = _lib.load().sq_attn_plan_debug_times(plan.handle, buf)
    t = np.array(list(buf)).reshape(8, 16)
    names = ["start", "alloc+sync", "TMA q,k landed", "MMA1 done", "softmax done", "MMA2 done", "epilogue+dealloc", "cluster sync 1", "reduction", "cluster sync 2"]
    for s_ in range(3):
        row = t[s_]
        order = [0, 1, 2, 3, 4, 5, 10, 6, 7, 9, 8]
        nm = {0: "start", 1: "alloc+mask+sync", 2: "TMA q,k landed", 3: "MMA1 done", 4: "softmax done", 5: "MMA2 done",
              10: "O staged+dealloc", 6: "pushed", 7: "cluster sync", 9: "weights", 8: "reduced"}
        print("split", s_, " ".join(f"{nm[k]}:+{int(row[k]-row[0])}" for k in order))
