"""Timeline of one k_em launch from CRISPGPU_EM_TRACE: phase costs per task, tail of the launch, what an ordering would change.

Per task eight int64: ns at task start / counts staged / histogram built / planes built / first EM iteration done / EM done /
task end, then SM | task << 16 | (2 iterations + called) << 48.

    CRISPGPU_EM_TRACE=/tmp/em.trace python scripts/em_trace.py --run      # benchmark batch (compact form), 5 launches, trace of the last
    python scripts/em_trace.py /tmp/em.trace
"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    sys.path.insert(0, ROOT)
    import bench
    from crisp_b200.abi import EngineConfig
    from crisp_b200.engine import Engine
    wl = bench.WORKLOADS["exome50x48"]
    b = bench.make_batch(wl, 8192, seed=1).compact()
    eng = Engine(EngineConfig(ploidy=np.full(wl["P"], wl["ploidy"]), options=wl["options"]), device=0)
    for _ in range(5):
        eng.run(b)
    print("em_ms", eng.timing()["em_ms"])
    eng.close()


def sim(d, slots):
    h = [0.0] * slots
    heapq.heapify(h)
    for x in d:
        heapq.heappush(h, heapq.heappop(h) + x)
    return max(h)


def report(path):
    t = np.fromfile(path, dtype=np.int64).reshape(-1, 8)
    t = t[t[:, 6] > 0]
    meta = t[:, 7]
    it, called = (meta >> 49), (meta >> 48) & 1
    dur = (t[:, 6] - t[:, 0]) / 1e3
    span = (t[:, 6].max() - t[:, 0].min()) / 1e3
    ev = np.concatenate([np.stack([t[:, 0], np.ones(len(t))], 1), np.stack([t[:, 6], -np.ones(len(t))], 1)])
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    slots = int(np.cumsum(ev[:, 1]).max())
    print(f"tasks {len(t)}  launch span {span:.1f} us  concurrent tasks {slots}  task mean {dur.mean():.1f} us  p50 {np.median(dur):.1f}  p99 {np.percentile(dur, 99):.1f}  max {dur.max():.1f}")
    print(f"busy task-time / (span x concurrent) = {dur.sum() / (span * slots):.3f};  last task end - last task start = {(t[:, 6].max() - t[:, 0].max()) / 1e3:.1f} us")
    names = ["stage counts", "histogram", "planes (DMMA)", "first iteration", "later iterations", "genotypes + pool statistics + store"]
    print("phase                                   mean us   share   (called tasks / others)")
    for k, nm in enumerate(names):
        ph = (t[:, k + 1] - t[:, k]) / 1e3
        print(f"  {nm:36s} {ph.mean():7.2f}  {ph.sum() / dur.sum():6.1%}   {ph[called == 1].mean():7.2f} / {ph[called == 0].mean():7.2f}")
    later = (t[:, 5] - t[:, 4]) / 1e3
    print(f"  per later iteration: {later.sum() / np.maximum(it - 1, 0).sum():.2f} us;  iterations mean {it.mean():.2f} max {it.max()}")
    A = np.stack([np.ones(len(t)), it, called], 1)
    c = np.linalg.lstsq(A, dur, rcond=None)[0]
    print(f"duration ~ {c[0]:.1f} + {c[1]:.2f} x iterations + {c[2]:.1f} x called  (us)")
    ideal = dur.sum() / slots
    order = np.argsort(t[:, 0])
    print(f"makespan: as run {sim(dur[order], slots):.1f} us, longest first {sim(dur[np.argsort(-dur)], slots):.1f} us, perfect split {ideal:.1f} us")


if __name__ == "__main__":
    if sys.argv[1] == "--run":
        run()
    else:
        report(sys.argv[1])
