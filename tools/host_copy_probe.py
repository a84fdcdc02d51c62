"""Measures the host-side costs that decide the cold end-to-end design (run on the GPU box):
pinned allocation, cudaHostRegister of a pageable buffer, pageable vs pinned D2H / H2D."""
import json
import time

import numpy as np
import torch

MB = 400
n = MB * 1024 * 1024 // 4
out = {}
torch.cuda.init()
t0 = time.perf_counter(); torch.zeros(1, device="cuda"); torch.cuda.synchronize(); out["ctx_init_s"] = time.perf_counter() - t0
d = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
torch.cuda.synchronize()
t0 = time.perf_counter(); pin = torch.empty(n, dtype=torch.float32, pin_memory=True); out["pinned_alloc_400MB_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); pg = np.empty(n, dtype=np.float32); pg[::1024] = 0; out["pageable_alloc_touch_s"] = time.perf_counter() - t0
pgt = torch.from_numpy(pg)
for name, dst in (("pinned", pin), ("pageable", pgt)):
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(d); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    out["d2h_%s_gbs" % name] = [round(MB / 1024 / t, 2) for t in ts]
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(dst); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    out["h2d_%s_gbs" % name] = [round(MB / 1024 / t, 2) for t in ts]
rt = torch.cuda.cudart()
pg2 = np.empty(n, dtype=np.float32); pg2[::1024] = 0
t0 = time.perf_counter(); r = rt.cudaHostRegister(pg2.ctypes.data, pg2.nbytes, 0); out["host_register_400MB_s"] = time.perf_counter() - t0
t2 = torch.from_numpy(pg2)
torch.cuda.synchronize(); t0 = time.perf_counter(); t2.copy_(d); torch.cuda.synchronize(); out["d2h_registered_gbs"] = round(MB / 1024 / (time.perf_counter() - t0), 2)
t0 = time.perf_counter(); rt.cudaHostUnregister(pg2.ctypes.data); out["host_unregister_s"] = time.perf_counter() - t0
# chunked D2H through a small pinned bounce buffer + memcpy (what a library can do for a pageable destination)
CH = 16 * 1024 * 1024 // 4
b = [torch.empty(CH, dtype=torch.float32, pin_memory=True) for _ in range(2)]
ev = [torch.cuda.Event() for _ in range(2)]
torch.cuda.synchronize(); t0 = time.perf_counter()
k = 0
pend = []
for o in range(0, n, CH):
    e = min(o + CH, n)
    if len(pend) == 2:
        po, pe, pk = pend.pop(0); ev[pk].synchronize(); pgt[po:pe].copy_(b[pk][:pe - po])
    b[k][:e - o].copy_(d[o:e], non_blocking=True); ev[k].record(); pend.append((o, e, k)); k ^= 1
for po, pe, pk in pend:
    ev[pk].synchronize(); pgt[po:pe].copy_(b[pk][:pe - po])
out["d2h_bounce_16MB_gbs"] = round(MB / 1024 / (time.perf_counter() - t0), 2)
import os
out["cpus"] = os.cpu_count()
print(json.dumps(out))
