"""Ray-sharded trace over 2 GPUs (NCCL) through the public CL_Tracer API: every rank passes the
same sources, traces its contiguous shard, and get_measured_rays() returns on every rank the
merged arrays in the reference order - bit-identical to the single-process oracle."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scene_name, kw, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as td

    from lightpycl_b200 import iterative_tracer as it, scenes

    torch.cuda.set_device(rank)
    td.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                          device_id=torch.device("cuda", rank))
    ls, meshes, params = scenes.CONFIGS[scene_name](**kw)
    tr = it.CL_Tracer(device=torch.device("cuda", rank))
    assert tr.world_size == world and tr._lineage
    tr.iterative_tracer(ls, meshes, **params)
    pos, pwr = tr.get_measured_rays()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), pos=pos, pwr=np.asarray(pwr).reshape(-1),
             live=np.array(tr.stats["live"]), local_rays=np.array(tr.results.ray_counts))
    td.destroy_process_group()


@pytest.mark.parametrize("scene_name,kw", [("c2_directivity_lens", dict(ray_count=3001)),
                                           ("c4_human_eye", dict(ray_count=150, sources=2, trace_iterations=6)),
                                           # 40 generations: past the 31 branch bits of a lineage id (dist.renumber_lineage)
                                           ("glass_cube_cavity", dict(ray_count=301, trace_iterations=40)),
                                           # fewer rays than ranks: rank 0 traces nothing and still takes part in the gather
                                           ("glass_cube_cavity", dict(ray_count=1, trace_iterations=5))])
def test_two_gpu_trace_equals_oracle(tmp_path, scene_name, kw):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from lightpycl_b200 import scenes
    from oracle import oracle

    mp.spawn(_worker, args=(2, _free_port(), scene_name, kw, str(tmp_path)), nprocs=2, join=True)
    ls, meshes, params = scenes.CONFIGS[scene_name](**kw)
    ref = oracle.trace(ls, meshes, power_sum="f64", **params)
    pos, pwr = oracle.measured_rays(ref["results"])
    totals = None
    for r in range(2):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        assert list(z["live"]) == ref["live"]
        assert np.array_equal(z["pos"][:, :3].copy().view(np.uint32), np.ascontiguousarray(pos[:, :3]).view(np.uint32))
        assert np.array_equal(z["pwr"].view(np.uint32), np.ascontiguousarray(np.asarray(pwr).reshape(-1)).view(np.uint32))
        totals = z["local_rays"] if totals is None else totals + z["local_rays"]
    assert list(totals) == [len(x[0]) for x in ref["results"]]  # the shards partition every generation
