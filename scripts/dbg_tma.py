import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import accfft_b200 as acc
from oracle import accfft_oracle as orc
os.environ["ACCFFT_B200_TMA"] = "2"
dev = torch.device("cuda:0")
for n, single in [((16, 1024, 30), True), ((1024, 16, 30), True), ((16, 1024, 30), False), ((8, 1024, 254), True), ((1024, 8, 254), True), ((8, 1536, 62), True), ((8, 640, 62), True)]:
    x = orc.synthetic_real(n, 3)
    p = acc.Plan(n, acc.Comm(0, 1, (1, 1)), r2c=True, single=single, flags=acc.ACCFFT_ESTIMATE)
    xd = torch.from_numpy(x.astype(np.float32 if single else np.float64)).to(dev)
    out = torch.zeros(p.alloc_max // (4 if single else 8), dtype=xd.dtype, device=dev)
    p.execute_r2c(xd, out)
    torch.cuda.synchronize()
    n2c = n[2] // 2 + 1
    X = out.cpu().numpy().view(np.complex64 if single else np.complex128)[: n[0] * n[1] * n2c].reshape(n[0], n[1], n2c)
    print(n, single, "err", orc.rel_l2(X, orc.forward_r2c(x)), flush=True)
    p.destroy()
