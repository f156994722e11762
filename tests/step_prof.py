"""two eager training steps of the bench workload (B=256, ViT-B/16 + BERT-base) for ncu launch lists / --set full captures (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easynlp_b200 import _lib as L
from easynlp_b200.engine import ClipEngine
from easynlp_b200.synthetic import random_state_dict, synthetic_batch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import b16_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = b16_config()
eng = ClipEngine(cfg)
eng.params.load_state_dict(random_state_dict(cfg, seed=1234))
px, ids = synthetic_batch(cfg, B, 77, seed=1234)
px, ids = px.cuda(), ids.cuda()
n0 = L.launch_count()
for i in range(2):
    eng.train_step(px, ids, lr=1e-5, use_graph=False)
    torch.cuda.synchronize()
    print("step", i, "launches so far", L.launch_count() - n0, flush=True)
