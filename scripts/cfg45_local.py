"""BASELINE configs 4 and 5 (1024^3 single and 768x768x384 double on a 2x4 pencil grid) on ONE device: the eight
ranks run as host threads with their own streams and exchange through accfftb_plan_attach_local_peers, so the
multi-rank kernels (segmented stores, uneven 49/49/49/46 pencils, barrier) run at full size without eight GPUs.
Checks: forward spectrum vs torch.fft.rfftn of the assembled field (test-only cross-check, on the GPU), round trip.
Usage: python scripts/cfg45_local.py [4|5|all]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import accfft_b200 as acc


def run(n, dims, single):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    nprocs = dims[0] * dims[1]
    tdt = torch.float32 if single else torch.float64
    cdt = torch.complex64 if single else torch.complex128
    es = 4 if single else 8
    gen = torch.Generator(device=dev).manual_seed(20240607 + (4 if single else 5))
    x = torch.rand(n, dtype=tdt, device=dev, generator=gen) * 2 - 1
    comms = [acc.Comm(r, nprocs, list(dims)) for r in range(nprocs)]
    plans = [acc.Plan(n, comms[r], r2c=True, single=single, wire=False) for r in range(nprocs)]
    acc.attach_local_peers(plans)
    streams = [torch.cuda.Stream() for _ in range(nprocs)]
    ins, outs, backs = [], [], []
    for r in range(nprocs):
        p = plans[r]
        i0, i1 = p.istart[0], p.istart[1]
        loc = x[i0:i0 + p.isize[0], i1:i1 + p.isize[1], :].contiguous()
        ins.append(loc if loc.numel() else torch.zeros(2, dtype=tdt, device=dev))
        outs.append(torch.zeros(max(p.alloc_max // es, 2), dtype=tdt, device=dev))
        backs.append(torch.zeros_like(ins[-1]))
        p.set_stream(streams[r].cuda_stream)
    torch.cuda.synchronize()
    errs = []

    def body(r):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(streams[r]):
                for _ in range(2):
                    plans[r].execute_r2c(ins[r], outs[r])
                    plans[r].execute_c2r(outs[r], backs[r])
        except Exception as e:
            errs.append((r, repr(e)))

    t0 = time.time()
    th = [threading.Thread(target=body, args=(r,)) for r in range(nprocs)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    torch.cuda.synchronize()
    assert not errs, errs
    assert not any(t.is_alive() for t in th), "rank thread hung"
    X = torch.fft.rfftn(x)
    num = torch.zeros((), dtype=torch.float64, device=dev)
    rt = torch.zeros((), dtype=torch.float64, device=dev)
    N = n[0] * n[1] * n[2]
    for r in range(nprocs):
        p = plans[r]
        osz, ost = p.osize, p.ostart
        cnt = osz[0] * osz[1] * osz[2]
        if cnt:
            mine = torch.view_as_complex(outs[r][:2 * cnt].view(-1, 2)).view(osz)
            ref = X[ost[0]:ost[0] + osz[0], ost[1]:ost[1] + osz[1], ost[2]:ost[2] + osz[2]]
            num += torch.sum(torch.abs(mine - ref).double() ** 2)
        if ins[r].numel() > 2:
            rt += torch.sum((backs[r] / N - ins[r]).double() ** 2)
    e_f = float(torch.sqrt(num) / torch.linalg.vector_norm(X).double())
    e_r = float(torch.sqrt(rt) / torch.linalg.vector_norm(x).double())
    tol = 1e-5 if single else 1e-12
    print("cfg n=%s dims=%s single=%s: forward rel-L2 vs torch.fft.rfftn %.3e, round trip %.3e (tol %g) %.1fs  kernels %s" %
          (n, dims, single, e_f, e_r, tol, time.time() - t0, plans[0].stage_kernels()), flush=True)
    for p in plans:
        p.destroy()
    # torch's own fft in the same precision is the yardstick for the forward spectrum: allow its error on top
    assert e_r <= tol and e_f <= (3e-6 if single else 5e-15 * 100), (e_f, e_r)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("5", "all"):
        run((768, 768, 384), (2, 4), False)
    if which in ("4", "all"):
        run((1024, 1024, 1024), (2, 4), True)
    print("cfg45_local OK")
