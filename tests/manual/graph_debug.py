import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pacoh_oracle as orc
from meta_learning_pacoh_b200 import meta_learn as ml

train = orc.sinusoid_tasks(20, 5, seed=26)
def make(decay):
    return ml.GPRegressionMetaLearnedVI(train, svi_batch_size=8, lr=2e-3, lr_decay=decay, random_seed=30)
def params(m):
    return torch.cat([m.posterior.loc.detach(), m.posterior.scale.detach()]).clone()

for decay in (1.0, 0.5):
    os.environ["PACOH_GRAPH"] = "0"
    a = make(decay); ref = []
    for k in range(40):
        a.run_steps(1); ref.append(params(a))
    os.environ["PACOH_GRAPH"] = "1"
    b = make(decay)
    b.run_steps(1)
    for chunk in (10, 10, 10, 3, 3):
        b.run_steps(chunk)
        n = b._state.steps
        print("decay", decay, "steps", n, "diff", float((params(b) - ref[n - 1]).abs().max()), "graph", b._graph is not None)
    c = make(decay)
    c.run_steps(1); c.run_steps(36)
    print("decay", decay, "1+36:", float((params(c) - ref[36]).abs().max()), c._state.steps)
