import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["PACOH_GRAPH"] = "0"
from oracle import pacoh_oracle as orc
from meta_learning_pacoh_b200 import meta_learn as ml
train = orc.sinusoid_tasks(20, 5, seed=26)
a = ml.GPRegressionMetaLearned(train, weight_decay=0.2, lr_decay=0.9, random_seed=30)
b = ml.GPRegressionMetaLearned(train, weight_decay=0.2, lr_decay=0.9, random_seed=30)
print("init diff", float((a._flat - b._flat).abs().max()))
for it in range(3):
    idx = a.rds_numpy.choice(20, size=5); b.rds_numpy.choice(20, size=5)
    la = a.map_step(idx)
    b.optimizer.zero_grad(); lb = b._loss_and_grad(idx); gb = b._gradflat.clone(); b.optimizer.step(); b.lr_scheduler.step()
    print(it, "loss", float(la), float(lb), "param diff", float((a._flat - b._flat).abs().max()), "m diff", float((a._mflat - b._mflat).abs().max()),
          "v diff", float((a._vflat - b._vflat).abs().max()), "state a", a._state.buf[:1].tolist(), a._state.buf.view(torch.float32)[2:5].tolist(),
          "steps", a._state.steps, b._state.steps, "b step tensors", {float(s["step"]) for s in b.optimizer.state.values()},
          "b lrs", [g["lr"] for g in b.optimizer.param_groups], [g["weight_decay"] for g in b.optimizer.param_groups])
    k = int((a._flat - b._flat).abs().argmax())
    print("   worst index", k, [n for n, (lo, hi) in a.arch.entries().items() if lo <= k < hi], float(a._flat[0, k]), float(b._flat[0, k]), "grad", float(gb[k]))
