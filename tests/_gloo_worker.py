"""Worker of tests/test_host_logic.py::test_two_rank_gloo_gather (launched by torch.distributed.run, gloo)."""
import os

import numpy as np
import torch.distributed as dist

from crisp_b200.abi import EngineConfig
from crisp_b200.shard import gather_results, shard_batch
from crisp_b200.synth import synth_batch
from oracle.loader import RNG_PHILOX, OracleEngine
from util import compare_results


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = EngineConfig(ploidy=np.full(6, 16), options=dict(chisq_permutation=1, max_iter=200))
    batch = synth_batch(60, 6, 16, 50, seed=12, variant_frac=0.5)
    sub, idx = shard_batch(batch, world, rank)
    local = OracleEngine(cfg).run(sub, mode=RNG_PHILOX)
    merged = gather_results(local, idx, batch.n_sites, 6)
    dist.barrier()
    if rank == 0:
        whole = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
        compare_results(merged, whole, tol=0.0, label="gloo")
        print("GLOO_MERGE_OK", merged.n_calls)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
