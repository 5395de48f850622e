import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
obj, (x, y, _, _) = bench.build_chain(bench.CFG, seed=0, device=dev)
xb, yb = x[:128].to(dev).contiguous(), y[:128].to(dev).contiguous()
for _ in range(3):
    obj.nested_step(xb, yb)
torch.cuda.synchronize()
