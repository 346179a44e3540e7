"""Host-side cost of one svgd_step (Python + ctypes + launches): run a tiny problem where GPU time is negligible."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meta_learning_pacoh_b200.data_sim import SinusoidDataset
from meta_learning_pacoh_b200.meta_learn import GPRegressionMetaLearnedSVGD
data = SinusoidDataset(random_state=np.random.RandomState(1)).generate_meta_train_data(64, 50)
m = GPRegressionMetaLearnedSVGD(data, num_particles=64, random_seed=3)
for _ in range(20): m.svgd_step(m._sample_task_indices())
torch.cuda.synchronize()
for n in (1, 2):
    t0 = time.perf_counter()
    for _ in range(300): m.svgd_step(m._sample_task_indices())
    t1 = time.perf_counter()          # host issue time only (no sync)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host issue %.1f us/step, with final sync %.1f us/step" % ((t1 - t0) / 300 * 1e6, (t2 - t0) / 300 * 1e6))
