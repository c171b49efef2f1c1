#!/usr/bin/env python
"""configs[3] per-user shape and configs[4] sweep shape alone (tools/subbench.py legs), one line each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import subbench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
r = subbench.small_entities(dev, 0, 1, 6547.5, E=E)
print(json.dumps({k: r[k] for k in ("entities_per_s", "ms", "plan", "mean_nit")}))
r = subbench.l2_sweep(dev, 0, 1, 6547.5, E=E // 2)
print(json.dumps({k: v for k, v in r.items() if k in ("models_per_s", "ms", "plan")}))
