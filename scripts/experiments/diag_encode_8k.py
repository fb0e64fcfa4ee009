"""Diagnostic: 8K i.i.d. RGB, seed 8 -- where do the GPU encoder and the oracle disagree about equal neighbours?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import rtc_b200
from oracle.oracle import Oracle
orc = Oracle()
ctx = rtc_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
x, y = 7681, 4320
W = x - 1
for seed in (8, 5, 11):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    keys = torch.randint(0, 256, (W * y * 3,), dtype=torch.uint8, device="cuda", generator=g)
    px = keys.view(-1, 3)
    eq = torch.nonzero((px[1:] == px[:-1]).all(1)).view(-1) + 1          # cells equal to their predecessor
    cap = rtc_b200.encode_capacity(x, y, 3)
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.encode(keys.data_ptr(), 0, x, y, 3, out.data_ptr(), cap, total.data_ptr())
    torch.cuda.synchronize()
    n = int(total.item())
    want_n = 20 * W * y - 19 * eq.numel() + y
    got = out[:n].cpu().numpy()
    want = orc.encode_planes(keys.cpu().numpy(), None, x, y, 3)
    print("seed", seed, "equal pairs (torch)", eq.numel(), "cells", eq.tolist()[:12], "mod 1280:", [int(c) % 1280 for c in eq.tolist()[:12]],
          "mod 20:", [int(c) % 20 for c in eq.tolist()[:12]], "col:", [int(c) % W for c in eq.tolist()[:12]])
    print("   gpu n", n, "torch formula", want_n, "oracle", want.size)
    m = min(got.size, want.size)
    d = np.nonzero(got[:m] != want[:m])[0]
    if d.size:
        b = int(d[0])
        print("   first differing byte", b, "~cell", b // 20, "gpu", bytes(got[b - 20:b + 24]), "oracle", bytes(want[b - 20:b + 24]))
