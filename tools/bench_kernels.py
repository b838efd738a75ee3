"""Stand-alone kernel roofline: scp_b200_linearize and scp_b200_reconstruct timed with CUDA events (L2 flushed
before every launch), algorithmic bytes per SURVEY.md 8(d) / DESIGN.md section 4, against MEASURED_PEAKS.json."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))
import torch
from path_planning import _capi

lib = _capi.load()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for r in range(reps):
        flush.fill_(r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

out = []
for name, B, N, K in (("C2 batch", 1024, 25, 50), ("C3 single", 1, 200, 100), ("C4 single", 1, 1000, 100), ("C5 batch 200", 512, 200, 50)):
    P = N * (N - 1) // 2
    pos = torch.rand((B, N, K, 2), dtype=torch.float64, device="cuda") * 100
    eta = torch.empty((B, K, P, 2), dtype=torch.float64, device="cuda")
    bound = torch.empty((B, K, P), dtype=torch.float64, device="cuda")
    minsep = torch.empty(B, dtype=torch.float64, device="cuda")
    first = torch.empty((B, 3), dtype=torch.int32, device="cuda")
    f = lambda: _capi.check(lib.scp_b200_linearize(pos.data_ptr(), B, N, K, 0.8, 0.01, eta.data_ptr(), bound.data_ptr(), minsep.data_ptr(), first.data_ptr(), st))
    med, mn = timeit(f)
    nbytes = 8 * B * (2 * N * K + 3 * P * K)
    out.append(dict(kernel="scp_linearize_kernel", case=name, B=B, N=N, K=K, ms_median=med, ms_min=mn, algorithmic_bytes=nbytes,
                    achieved_gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / peak))
    del eta, bound
    acc = torch.randn((B, N, K, 2), dtype=torch.float64, device="cuda")
    p0 = torch.rand((B, N, 2), dtype=torch.float64, device="cuda"); v0 = torch.zeros_like(p0)
    po = torch.empty_like(acc); ve = torch.empty_like(acc)
    g = lambda: _capi.check(lib.scp_b200_reconstruct(acc.data_ptr(), p0.data_ptr(), v0.data_ptr(), B, N, K, 0.2, po.data_ptr(), ve.data_ptr(), st))
    med, mn = timeit(g)
    nbytes = 48 * B * N * K
    out.append(dict(kernel="scp_reconstruct_kernel", case=name, B=B, N=N, K=K, ms_median=med, ms_min=mn, algorithmic_bytes=nbytes,
                    achieved_gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / peak))
for o in out:
    print(json.dumps(o))
