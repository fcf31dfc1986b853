import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from pldepth_b200 import ops, synth
dev = torch.device('cuda', 0)
import os
B, H, W, K, R = int(os.environ.get('PS_B', 32)), 448, 448, int(os.environ.get('PS_K', 5)), int(os.environ.get('PS_R', 100000))
strategy = sys.argv[1] if len(sys.argv) > 1 else "thresholded"
emit = not (len(sys.argv) > 2 and sys.argv[2] == "no-emit")
f = {"thresholded": 1.5, "information": 5}[strategy]
base = synth.depth_map(H, W, 7)
gt = torch.from_numpy(np.stack([np.roll(base, 31 * b, axis=1) for b in range(B)])).to(dev)
mask = torch.ones((B, H, W), device=dev); pred = torch.randn((B, H, W, 1), device=dev)
for i in range(4):
    ops.fused_step_scored(mask, gt, pred, K, int(R * f), R, strategy, seed=1, offset=i, want_rankings=emit)
torch.cuda.synchronize()
