import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for n in (1, 4, 8, 10, 12, 14, 16):
    r = bench.bench_replicas(bench.CFG, dev, n, 30, 3)
    print(n, r.get("aggregate_outer_steps_per_s"), r.get("device_span_ms"), r.get("wall_ms"), r.get("error"), flush=True)
