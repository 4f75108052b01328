"""Sweep GEMM shapes / transposes / kernel variants against NumPy (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mellon_b200 as mb
be = mb.get_backend()
rng = np.random.default_rng(0)
bad = 0
ONLY = os.environ.get("SWEEP_ONLY")
for variant in (0, 2, 4):
    be.set_option("gemm", variant)
    for (m, n, k) in [(128, 128, 16), (128, 128, 32), (128, 128, 64), (130, 257, 77), (130, 256, 77), (130, 256, 96), (130, 258, 64),
                      (256, 256, 100), (64, 64, 33), (300, 66, 1000), (512, 384, 200), (640, 640, 2100), (1300, 300, 4200), (200, 200, 9000)]:
        for ta in (0, 1):
            for tb in (0, 1):
                if ONLY and ONLY != f"{variant}-{m}-{n}-{k}-{ta}-{tb}":
                    continue
                A = rng.standard_normal((k, m) if ta else (m, k))
                B = rng.standard_normal((n, k) if tb else (k, n))
                try:
                    out = be.gemm(A, B, trans_a=bool(ta), trans_b=bool(tb)).numpy()
                    ref = (A.T if ta else A) @ (B.T if tb else B)
                    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
                    nbad = int(np.sum(np.abs(out - ref) > 1e-10 * np.max(np.abs(ref))))
                except Exception as e:  # noqa
                    err, nbad = float("nan"), -1
                    print("EXC", e)
                flag = "" if err < 1e-12 else "  <-- BAD"
                if flag:
                    bad += 1
                    rows, cols = np.nonzero(np.abs(out - ref) > 1e-10 * np.max(np.abs(ref))) if nbad > 0 else ([], [])
                    extra = f" bad rows {sorted(set(rows))[:8]}..{sorted(set(rows))[-3:]} cols {sorted(set(cols))[:8]}..{sorted(set(cols))[-3:]} sample out/ref {out[rows[0], cols[0]]:.6g}/{ref[rows[0], cols[0]]:.6g}" if nbad > 0 else ""
                else:
                    extra = ""
                print(f"variant {variant} m={m} n={n} k={k} ta={ta} tb={tb}: err {err:.2e} nbad {nbad}{flag}{extra}")
print("BAD CASES", bad)
