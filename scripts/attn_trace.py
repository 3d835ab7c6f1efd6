#!/usr/bin/env python
"""Pipeline timeline of CTA 0 of the lean-softmax attention kernel (clock64 stamps written through
gyre_b200_debug_attention_trace): python scripts/attn_trace.py [variant]

Slots per tile, softmax groups (actors 0 / 1, thread of row 0): 0 enter tile, 1 scores ready (s_full), 2 scores in
registers (s_free raised), 3 three chunks exponentiated, 4 previous P V retired (pv_done), 5 tile exponentiated,
6 before the P store wait, 7 p_full raised.  Actor 2 (S issuer): 2g issue, 2g+1 committed.  Actor 3 (P V issuer):
2g operands ready (p_full seen), 2g+1 committed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200 import _native as N
dev = torch.device("cuda", 0)
lib = N.load()
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2258   # 2000 + A5_TRACE (256) + A5_POLY25 (2)
N.set_tunable("ATT_VARIANT", variant)
B, heads, Nq, d = 16, 8, 4096, 40
C = heads * d
NT = Nq // 128
qkv = torch.randn(B, Nq, 3 * C, device=dev).half()
buf = torch.zeros(4 * NT * 8, dtype=torch.int64, device=dev)
for _ in range(2):
    N.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads)
torch.cuda.synchronize()
N.check(lib.gyre_b200_debug_attention_trace(buf.data_ptr(), buf.numel()), "trace")
N.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads)
torch.cuda.synchronize()
N.check(lib.gyre_b200_debug_attention_trace(None, 0), "trace")
t = buf.cpu().view(4, NT, 8)
t0 = int(t[t > 0].min())
rel = (t - t0).clamp(min=-1)
print(f"variant {variant}: CTA 0, clocks relative to the first stamp")
for j in list(range(0, 4)) + list(range(14, 20)):
    print(f"tile {j:2d}")
    for g in (0, 1):
        r = rel[g, j].tolist()
        print(f"  softmax g{g}: enter {r[0]:7d}  s_full {r[1]:7d} (+{r[1]-r[0]:4d})  loaded {r[2]:7d} (+{r[2]-r[1]:4d})  3chunks {r[3]:7d} (+{r[3]-r[2]:4d})"
              f"  pv_done {r[4]:7d} (+{r[4]-r[3]:4d})  exp end {r[5]:7d} (+{r[5]-r[4]:4d})  st {r[6]:7d} (+{r[6]-r[5]:4d})  p_full {r[7]:7d} (+{r[7]-r[6]:4d})")
    s = rel[2, j].tolist(); pv = rel[3, j].tolist()
    print(f"  S issue g0 {s[0]:7d} ->{s[1]:7d}   g1 {s[2]:7d} ->{s[3]:7d}      PV g0 ready {pv[0]:7d} ->{pv[1]:7d}   g1 ready {pv[2]:7d} ->{pv[3]:7d}")
per_tile = (rel[0, NT - 2, 7] - rel[0, 2, 7]).item() / (NT - 4)
print(f"steady state: {per_tile:.0f} clocks per KV tile (group 0)")
for g in (0, 1):
    ph = torch.stack([rel[g, 2:NT - 1, k + 1] - rel[g, 2:NT - 1, k] for k in range(7)]).float().mean(dim=1).tolist()
    nxt = (rel[g, 3:NT, 0] - rel[g, 2:NT - 1, 7]).float().mean().item()
    print(f"group {g} mean phase lengths: wait s_full {ph[0]:.0f}, load {ph[1]:.0f}, 3 chunks {ph[2]:.0f}, wait pv_done {ph[3]:.0f}, last chunk {ph[4]:.0f}, vote {ph[5]:.0f}, st wait + arrive {ph[6]:.0f}, loop {nxt:.0f}")
