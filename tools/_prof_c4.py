import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests import workloads as wl
w, x0, th = wl.mseir_batch(8192, N=1000)
k = wl.make_kode(w)
x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(th).cuda()
for it in range(2):
    mu, var = k.solve_mv(x0d, theta=thd)
    torch.cuda.synchronize()
