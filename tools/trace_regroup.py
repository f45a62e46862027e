"""tools/trace_regroup.py N — (diagnosis build, BACON_IVP_LIB=variants/libbacon_ivp_trace.so) time axis of CTA 0's
regroupings for the Lorenz RK45 headline workload with N trajectories."""
import ctypes as C
import sys
import numpy as np
import torch
import bacon_b200 as B
from bacon_b200 import ensembles as E
from bacon_b200._lib import lib

n = int(sys.argv[1])
w = E.LORENZ
y0 = torch.from_numpy(E.lorenz_y0(np.arange(n))).cuda()
p = torch.tensor(w["params"], dtype=torch.float64).cuda()
s = (B.RK45.new(3).with_dt_min(w["dt_min"]).with_dt_max(w["dt_max"]).with_tolerance(w["tol"]).with_start(0.0)
     .with_end(w["t_end"]).with_derivative("lorenz"))
L = lib()
buf = (C.c_ulonglong * (3 * 4096))()
for rep in range(3):
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    out = s.solve_ivp_ensemble_device(y0, p, shared_params=True)
    torch.cuda.synchronize()
    ms = B.last_launch()["kernel_ms"]
    k = L.bacon_debug_trace(buf, 4096)
ev = np.array(buf[:3 * k], dtype=np.uint64).reshape(k, 3)
print(f"n={n} kernel {ms:.3f} ms, {k} regroupings of CTA 0 (grid {B.last_launch()['grid']}x{B.last_launch()['block']})")
t_end = int(ev[-1, 0])
names = {1001: "warp at the meeting", 1003: "warp past the 2nd barrier", 1004: "request raised by a lane of warp"}
for t, wa, run in ev:
    if int(wa) >= 1000:
        print(f"  {-(t_end - int(t)) / 1000:9.1f} us   {names[int(wa)]} {int(run)}")
    else:
        print(f"  {-(t_end - int(t)) / 1000:9.1f} us before the last   warps {int(wa):2d}  live {int(run):4d}   (plan, after the 1st barrier)")
