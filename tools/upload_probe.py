"""Probe host->device upload speed through the C ABI: pinned vs pageable, 400 MB."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mellon_b200 as mb
be = mb.get_backend()
x = np.random.default_rng(0).random((1_000_000, 50))
xp = be.pinned_empty(x.shape); xp[...] = x
for name, arr in (("pageable", x), ("pinned", xp)):
    for rep in range(3):
        be.sync(); t0 = time.perf_counter(); d = be.upload(arr); be.sync(); dt = time.perf_counter() - t0
        print(name, rep, f"{dt*1e3:.1f} ms  {arr.nbytes/dt/1e9:.1f} GB/s")
        d.free()
