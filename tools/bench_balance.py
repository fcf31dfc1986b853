import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from pldepth_b200.step import FusedPLStep
dev = torch.device('cuda', 0)
B, H, W, K = 32, 448, 448, 5
gt = torch.rand((B, H, W), device=dev); mask = torch.ones((B, H, W), device=dev); pred = torch.randn((B, H, W, 1), device=dev)
for R in (100000, 37*256*10, 37*256*11, 37*256*12):
    st = FusedPLStep(K, R, seed=1)
    for _ in range(5): st.run(gt, mask, pred)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30): st.run(gt, mask, pred)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print("R=%d: %.1f us/step, %.4e lists/s, %.2f ps/list" % (R, ms * 1e3, B * R / (ms * 1e-3), ms * 1e9 / (B * R)))
