"""End-to-end rotation with one host thread per engine context (debugging aid for bench.py's e2e leg)."""
import sys
import threading
import time

import numpy as np

from crisp_b200.abi import EngineConfig
from crisp_b200.engine import Engine, pin_batch
from crisp_b200.synth import synth_batch

nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = EngineConfig(ploidy=np.full(50, 48), options={})
batch = pin_batch(synth_batch(8192, 50, cfg.ploidy, 200, seed=1, variant_frac=0.15).compact())
engs = [Engine(cfg) for _ in range(nctx)]
for e in engs:
    e.run(batch); e.run(batch)
start = threading.Barrier(nctx + 1)
def worker(e, n):
    start.wait()
    for _ in range(n):
        e.wait(e.submit(batch), copy=False)
per = [steps // nctx + (1 if k < steps % nctx else 0) for k in range(nctx)]
th = [threading.Thread(target=worker, args=(e, n)) for e, n in zip(engs, per)]
for t in th:
    t.start()
start.wait()
t0 = time.perf_counter()
for t in th:
    t.join()
dt = time.perf_counter() - t0
print(f"contexts {nctx} steps {steps}: {dt * 1e3 / steps:.3f} ms/step, {8192 * steps / dt / 1e6:.2f} M sites/s")
