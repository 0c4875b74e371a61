"""Developer tool: measured tensor-pipe peaks (tcgen05 kind::i8 issue loop, DMMA loop), burst and sustained, as one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ggp_b200
dev = torch.device("cuda:0"); eng = ggp_b200.Engine.get(dev)
out = {"i8_burst": eng.probe_i8_peak(2000)}
t0 = time.time()
out["i8_sustained"] = eng.probe_i8_peak(60000)      # ~9 launches of ~0.1-0.2 s each back to back
out["i8_sustained_wall_s"] = time.time() - t0
out["dmma"] = eng.probe_dmma_peak(20000)
print(json.dumps(out))
