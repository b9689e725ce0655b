"""Experiments that were written after round 1's GPU budget was spent and have NOT been run on a GPU yet.
They are off by default in the library and SKIPPED here unless DIR_B200_RUN_EXPERIMENTS=1:

    DIR_B200_RUN_EXPERIMENTS=1 python -m pytest tests/test_gpu_zx_experiments.py -m gpu -x -q        (>= 2 GPUs)

  DIR_B200_SHARD_ONEROW=1   one-row (numeric) fields of the sharded layer as replicated parameters
  DIR_B200_IDS=peer         the id exchange over peer memory instead of an NCCL all-to-all
Each flag must leave the multi-rank parity test of tests/test_gpu_sharded.py green.
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DIR_B200_RUN_EXPERIMENTS") != "1",
                                 reason="unvalidated experiments: set DIR_B200_RUN_EXPERIMENTS=1 to run them")]


@pytest.mark.parametrize("flags", [{"DIR_B200_SHARD_ONEROW": "1"}, {"DIR_B200_IDS": "peer"},
                                   {"DIR_B200_SHARD_ONEROW": "1", "DIR_B200_IDS": "peer"}])
@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_parity_under_experiment_flags(pkg, cuda, flags, world, monkeypatch):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from tests import test_gpu_sharded as T
    for k, v in flags.items():
        monkeypatch.setenv(k, v)                 # inherited by the spawned ranks
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = T._free_port()
    procs = [ctx.Process(target=T._rank_main, args=(r, world, port, out, "peer")) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
