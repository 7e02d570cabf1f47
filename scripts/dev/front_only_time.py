"""Time of the fast mode alone over 2^21 C3 candidates with SM clock samples (for BX_TC_ABLATE ablations): ms, median MHz."""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine as bengine
d, n, M = 20, 4096, 1 << 22
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
Y = np.sin(3 * X.sum(1)).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = bengine.Posterior(X, Y, raw).build()
key = bayex.random.key(1)
fn = lambda: post.sample_front(key, dims, 0, M, "UCB", 0.01, 0.0, 50, engine=_lib.ENGINE_TCGEN05_F16_RAW)
fn(); torch.cuda.synchronize()
clocks = []
stop = False
def poll():
    while not stop:
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5).stdout
            c, p = o.strip().split(",")
            clocks.append((float(c), float(p)))
        except Exception:
            pass
        time.sleep(0.05)
th = threading.Thread(target=poll); th.start()
t0 = time.perf_counter()
reps = 8
for _ in range(reps):
    fn()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps * 1e3
stop = True; th.join()
mhz = np.median([c for c, _ in clocks]) if clocks else float("nan")
pw = np.mean([p for _, p in clocks]) if clocks else float("nan")
print(f"  fast mode alone, 2^22 candidates: {dt:.1f} ms per launch, median SM clock {mhz:.0f} MHz, mean power {pw:.0f} W -> {dt * mhz * 1e3 / 1e6:.1f} M cycles")
