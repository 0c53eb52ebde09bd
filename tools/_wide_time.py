import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import workloads as wl
for name, maker, B in (("mseir", lambda B: wl.mseir_batch(B, N=1000), 8192), ("lorenz", lambda B: wl.lorenz_batch(B, N=5000), 4096)):
    w, x0, th = maker(B)
    k = wl.make_kode(w, force_dense=True)
    k.set_timing(True)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(th).cuda()
    for it in range(3):
        mu, var = k.solve_mv(x0d, theta=thd)
        torch.cuda.synchronize()
    print(name, k.kernel_family, B, "mv kernel ms", k.last_kernel_ms(), flush=True)
    del mu, var
