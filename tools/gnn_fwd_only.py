"""Run a few cfg4 training steps (for ncu captures of the GCN kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
w = dict(bench.WORKLOADS["cfg4"]); w["B"] = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
prob = bench.build_problem(w, 0); model = bench.build_model(w, prob, dev); batch = bench.device_batch(prob, dev)
for _ in range(3):
    model.fit_step(batch)
torch.cuda.synchronize()
