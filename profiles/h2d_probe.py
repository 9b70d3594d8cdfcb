"""Pinned host -> device copy rate of the box (what bounds bench.py's e2e leg): 1 GiB in 175 MB pieces, like one batch of 148 frames."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viamd_b200 as vb
n = 1 << 30; piece = 175 << 20
h = vb.host_alloc_pinned(n); d = vb.device_alloc(0, n)
vb.memcpy_h2d(0, d, h, n); vb.device_synchronize(0)
res = {}
for label, step in (("whole", n), ("175MB_pieces", piece)):
    t0 = time.perf_counter()
    for r in range(3):
        o = 0
        while o < n:
            c = min(step, n - o); vb.memcpy_h2d(0, d + o, h + o, c); o += c
    vb.device_synchronize(0)
    res[label + "_GBps"] = 3 * n / (time.perf_counter() - t0) / 1e9
t0 = time.perf_counter()
for r in range(3): vb.memcpy_d2h(0, h, d, n)
vb.device_synchronize(0); res["d2h_GBps"] = 3 * n / (time.perf_counter() - t0) / 1e9
print(json.dumps(res))
