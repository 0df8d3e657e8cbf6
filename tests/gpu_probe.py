"""Standalone GPU probe (run under `timeout`): checks every conv case on the tcgen05 path and on the SIMT path against
the PyTorch emulator and prints one line per case.  Kept out of pytest so that a device trap in an experimental
kernel cannot take the whole test session's CUDA context with it."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from bs_yolo_b200 import _cabi as A  # noqa: E402
import convcase  # noqa: E402


def main():
    impls = {"simt": A.IMPL_SIMT, "tc": A.IMPL_TCGEN05}
    which = sys.argv[1:] or ["simt", "tc"]
    bad = 0
    for w in which:
        for name in convcase.CASES:
            t = time.time()
            try:
                got, ref, impl = convcase.run_case(name, impls[w])
                err = float((got - ref).norm() / (ref.norm() + 1e-12))
                mx = float((got - ref).abs().max())
                ok = err < 4e-3
            except Exception as e:  # noqa: BLE001
                print(f"{w:5s} {name:24s} EXC {type(e).__name__}: {e}", flush=True)
                bad += 1
                if "CUDA" in str(e) or "cuda" in str(e):
                    print("device error: aborting probe", flush=True)
                    return 2
                continue
            bad += (not ok)
            print(f"{w:5s} {name:24s} impl={impl} rel={err:.2e} maxabs={mx:.3e} {'OK' if ok else 'FAIL'} ({time.time()-t:.2f}s)",
                  flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
