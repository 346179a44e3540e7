import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import pacoh_oracle as orc
from meta_learning_pacoh_b200 import meta_learn as ml
train, test = orc.sinusoid_tasks(20, 5, seed=26, n_test=50)
m = ml.GPRegressionMetaLearned(train, weight_decay=0.2, num_iter_fit=1, random_seed=30)
m.meta_fit(verbose=False, n_iter=5)
sd = m.state_dict()
m2 = ml.GPRegressionMetaLearned(train, weight_decay=0.2, num_iter_fit=1, random_seed=25)
m2.load_state_dict(sd)
print('pack equal after load', torch.equal(m._pack(), m2._pack()))
for (k1,v1),(k2,v2) in zip(m.optimizer.state_dict()['state'].items(), m2.optimizer.state_dict()['state'].items()):
    for kk in v1:
        if not torch.equal(torch.as_tensor(v1[kk]).cpu(), torch.as_tensor(v2[kk]).cpu()): print('state differs', k1, kk)
print([ (g['lr'], g['weight_decay'], len(g['params'])) for g in m.optimizer.param_groups])
print([ (g['lr'], g['weight_decay'], len(g['params'])) for g in m2.optimizer.param_groups])
m.rds_numpy, m2.rds_numpy = np.random.RandomState(5), np.random.RandomState(5)
for it in range(3):
    for mm in (m, m2):
        mm.optimizer.zero_grad()
    i1 = m.rds_numpy.choice(20, size=5); i2 = m2.rds_numpy.choice(20, size=5)
    l1 = m._loss_and_grad(i1); l2 = m2._loss_and_grad(i2)
    g1 = torch.cat([t.grad.reshape(-1) for _, t in m._named_flat()]); g2 = torch.cat([t.grad.reshape(-1) for _, t in m2._named_flat()])
    print(it, np.array_equal(i1, i2), l1.item(), l2.item(), 'grad equal', torch.equal(g1, g2))
    m.optimizer.step(); m2.optimizer.step()
    print('   pack equal after step', torch.equal(m._pack(), m2._pack()), (m._pack()-m2._pack()).abs().max().item())
