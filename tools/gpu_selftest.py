"""Run every kernel-parity group (tests/kernel_checks.py) in its own process on the GPU box.

Usage (under gpurun):  python tools/gpu_selftest.py [group ...] > gpurun_out/selftest.log
A faulting kernel poisons its CUDA context; process isolation keeps the other groups' results."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_one(name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import kernel_checks
    kernel_checks.GROUPS[name]()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        sys.exit(0)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import kernel_checks
    names = sys.argv[1:] or list(kernel_checks.GROUPS)
    summary = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True,
                               text=True, timeout=240)
            ok, out = r.returncode == 0, r.stdout + r.stderr
        except subprocess.TimeoutExpired as e:
            ok, out = False, (e.stdout or b"").decode(errors="replace") + "\nTIMEOUT"
        print(f"===== {n}: {'PASS' if ok else 'FAIL'} ({time.time() - t0:.1f}s)")
        print(out[-6000:] if not ok else out[-2500:], flush=True)
        summary[n] = ok
    print("SUMMARY", summary)
    sys.exit(0 if all(summary.values()) else 1)
