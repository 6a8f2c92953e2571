"""Per-batch stage timeline of the end-to-end rotation (debugging aid): when each batch's copies and kernels ran."""
import ctypes as C
import sys
import time

import numpy as np

from crisp_b200.abi import EngineConfig
from crisp_b200.engine import Engine, load_library, pin_batch
from crisp_b200.synth import synth_batch

nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
use_advance = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
cfg = EngineConfig(ploidy=np.full(50, 48), options={})
batch = pin_batch(synth_batch(8192, 50, cfg.ploidy, 200, seed=1, variant_frac=0.15).compact())
engs = [Engine(cfg) for _ in range(nctx)]
lib = load_library()
lib.crispgpu_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
for e in engs:
    e.run(batch); e.run(batch)
tickets = [None] * nctx
rows = []
host = []
def tl(k):
    out = (C.c_float * 7)()
    lib.crispgpu_debug_timeline(engs[k].h, engs[0].h, out)
    return list(out)
t00 = time.perf_counter()
for step in range(steps):
    k = step % nctx
    h0 = time.perf_counter()
    nxt = (k + 1) % nctx
    if use_advance and nctx > 1 and tickets[nxt] is not None:
        engs[nxt].advance(tickets[nxt])
    h1 = time.perf_counter()
    if tickets[k] is not None:
        engs[k].wait(tickets[k], copy=False); rows.append((k, tl(k)))
    h2 = time.perf_counter()
    tickets[k] = engs[k].submit(batch)
    h3 = time.perf_counter()
    host.append((h1 - h0, h2 - h1, h3 - h2))
for d in range(nctx):
    k = (steps + d) % nctx
    if tickets[k] is not None:
        engs[k].wait(tickets[k], copy=False); rows.append((k, tl(k)))
print("wall per step ms", (time.perf_counter() - t00) * 1e3 / steps)
base = rows[0][1][0]
print("ctx   START     H2D    EVAL    PERM  FILTER    CALL     D2H   (ms since first batch's start)")
for k, r in rows:
    print(k, " ".join(f"{x - base:7.3f}" for x in r))
print("host ms per step (advance, wait, submit):")
for h in host:
    print(" ".join(f"{x * 1e3:6.3f}" for x in h))
