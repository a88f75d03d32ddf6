import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gymnasium_planar_robotics_b200 as gpr
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
noise = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
env = gpr.BenchmarkPlanningVecEnv(256, np.ones((3, 3)), N, device='cuda:0', std_noise=noise, seed=3)
env.reset(seed=3)
torch.cuda.synchronize()
print('reset ok', env.get_state()['pos'][0])
for i in range(3):
    env.step(torch.zeros((256, 2 * N), device='cuda:0'))
torch.cuda.synchronize()
print('steps ok')
