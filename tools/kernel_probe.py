"""Launch every non-attention hot-path kernel at config-2 shapes (68m -> 7B, 128-node tree) for an ncu capture:
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active
    --clock-control none --csv --log-file gpurun_out/kernels.csv python tools/kernel_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sequoia_b200 import ops
from sequoia_b200.tree import _Static

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = "cuda:0"
torch.manual_seed(0)
gm = torch.load(os.path.join(ROOT, "A100_growmaps/68m_7b/growmaps/A100-CNN-68m-7b-stochastic.pt"))
st = _Static(gm, dev)
S, V, M, P = gm["size"], 32000, 384, 193
h, H, Hkv, D, I, L = 4096, 32, 32, 128, 11008, 32
f16 = torch.float16
state = torch.zeros(16, dtype=torch.int32, device=dev); state[0] = P
tokens = torch.randint(3, V, (M,), device=dev)
pos = torch.arange(M, device=dev)
sto = torch.arange(M, device=dev)
draft_logits = (torch.randn(M, V, device=dev) * 2).to(f16)
target_logits = (torch.randn(S, V, device=dev) * 2).to(f16)
rand = torch.empty(S, V, device=dev, dtype=f16).uniform_()
r = torch.rand(M, device=dev).to(f16)
noise = torch.empty(V, device=dev, dtype=f16).exponential_()
acc = torch.zeros(S, dtype=torch.int32, device=dev)
for rep in range(3):
    for lv in st.levels:                                           # sq_sample_level: 5 tree levels
        ops.sample_level(draft_logits, rand, lv["n_parents"], lv["k"], 0.6, 0, parent_rows=lv["parents"],
                         child_first=lv["first"], n_branch=lv["nb"], tokens=tokens, state=state)
    s2 = state.clone(); t2 = tokens.clone(); p2 = pos.clone()
    ops.accept_stochastic(target_logits, draft_logits, r, noise, st.succ_off, st.succ, st.depth, S, 0.6, t2, p2, acc, s2, M)
    ops.softmax_T(target_logits, 0.6)
    ops.argmax_rows(target_logits)
# model-side element-wise kernels at 7B / 128 rows
hid = torch.randn(M, h, device=dev).to(f16); proj = torch.randn(M, h, device=dev).to(f16)
w = torch.ones(h, device=dev, dtype=f16); out = torch.empty_like(hid)
qkv = torch.randn(M, (H + 2 * Hkv) * D, device=dev).to(f16)
kc = torch.zeros(L, 1, Hkv, M, D, device=dev, dtype=f16); vc = torch.zeros_like(kc)
inv = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D)); t = torch.arange(M).float()
emb = torch.cat([torch.outer(t, inv)] * 2, -1)
cos, sin = emb.cos().to(f16).to(dev), emb.sin().to(f16).to(dev)
gu = torch.randn(M, 2 * I, device=dev).to(f16); act = torch.empty(M, I, device=dev, dtype=f16)
table = torch.randn(V, h, device=dev).to(f16)
st5 = torch.zeros(16, dtype=torch.int32, device=dev); st5[3], st5[4] = 5, 150
idx = torch.tensor([160, 170, 180, 190, 200, 0, 0, 0], dtype=torch.int32, device=dev)
for rep in range(3):
    ops.embed_rows(table, tokens, S, hid, state=state, n0=0)
    ops.rmsnorm(hid, w, out, S, 1e-5)
    ops.add_rmsnorm(hid, proj, w, out, S, 1e-5)
    ops.rope_kv_append(qkv, H, Hkv, D, cos, sin, pos, sto, S, kc[rep], vc[rep], M, state=state, n0=0)
    ops.silu_mul(gu, act, S)
    ops.kv_gather(kc, vc, idx, 0, 0, state=st5, max_n=8)
torch.cuda.synchronize()
print("probe done")
