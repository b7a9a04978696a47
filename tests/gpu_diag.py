"""Runs every GPU parity case and prints ALL problems (pytest -x stops at the first).  Usage on the GPU box:
    python tests/gpu_diag.py > gpurun_out/diag.log 2>&1"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_cases  # noqa: E402
from phyngsc_b200 import api  # noqa: E402


def main():
    only = sys.argv[1:] or None
    ctx = api.Context(0, max_batch_bytes=64 << 20, max_subblocks=256)
    bad = 0
    for case in gpu_cases.CASES:
        if only and case[0] not in only:
            continue
        t = time.time()
        try:
            probs = gpu_cases.run_case(ctx, case)
        except Exception as e:  # noqa: BLE001
            probs = [f"EXCEPTION {e!r}"]
        print(f"{case[0]:28s} {'OK' if not probs else 'FAIL'}  ({time.time() - t:.2f}s)")
        for p in probs[:12]:
            print("    ", p)
        if len(probs) > 12:
            print(f"     ... {len(probs) - 12} more")
        bad += bool(probs)
        sys.stdout.flush()
    print("cases failing:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
