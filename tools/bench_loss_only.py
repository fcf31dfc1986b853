import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from pldepth_b200 import ops, synth
dev = torch.device('cuda', 0)
B, H, W, K, R = 32, 448, 448, 5, 100000
gt = torch.rand((B, H, W), device=dev); mask = torch.ones((B, H, W), device=dev); pred = torch.randn((B, H, W, 1), device=dev)
vf, nv = ops.mask_compact(mask, H, W)
rank, _ = ops.sample_lists_philox(gt, vf, nv, K, R, 1, 0, 0)
grad = torch.empty_like(pred)
for _ in range(3): ops.listmle_fwd_bwd(rank, pred, B, K, 1.0/(B*R), grad_out=grad)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): ops.listmle_fwd_bwd(rank, pred, B, K, 1.0/(B*R), grad_out=grad)
b.record(); torch.cuda.synchronize()
print("loss-only (fed rankings) C2 shapes: %.1f us/call, %.3e lists/s" % (a.elapsed_time(b)/20*1e3, B*R/(a.elapsed_time(b)/20*1e-3)))
