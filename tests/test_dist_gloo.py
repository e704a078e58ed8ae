"""N > 1 host logic on CPU (gloo, world_size 2): ray sharding, the per-iteration all-reduce that
implements the reference's global termination tests (it:372-391) and the final lineage-ordered
gather of measured rays (it:395-411).  Each rank runs the ORACLE kernels on its shard (the CUDA
kernels need a GPU; the collective/ordering logic under test is the same lightpycl_b200.dist code
the CUDA tracer calls) and the merged result must equal the single-process oracle bit for bit."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_oracle_trace(rank, world, port, scene_name, kw, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as td

    from lightpycl_b200 import dist, scenes
    from oracle import oracle

    td.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    group = dist.resolve_group(None)
    assert dist.world(group) == (world, rank)
    ls, meshes, params = scenes.CONFIGS[scene_name](**kw)
    origin, direc, power, input_power = oracle.flatten_sources(ls)
    n_all = len(origin)
    lo, hi = dist.shard_bounds(n_all, world, rank)
    origin, direc = origin[lo:hi], direc[lo:hi]
    pw = np.ascontiguousarray(power.reshape(-1)[lo:hi])
    lineage = np.arange(lo, hi, dtype=np.int64)
    prev = np.full(hi - lo, -2, np.int32)
    scene = oracle.Scene(meshes)
    mrl, env = np.float32(params["max_ray_len"]), np.float32(params["ior_env"])
    recs, lins, per_iter, lives = [], [], [], []
    epoch_start = 0
    for it in range(params["trace_iterations"]):
        n = len(origin)
        if it - epoch_start >= dist.PATH_BITS:  # the branch bits are used up: new lineage epoch (a collective)
            lineage = dist.renumber_lineage(torch.from_numpy(lineage), group, "cpu").numpy()
            epoch_start = it
        meas = np.zeros(n, np.int32)
        if n:
            nt, h, ni = oracle.intersect(scene, origin, direc, mrl)
            pp = oracle.postproc(scene, origin, direc, prev, nt, h, ni, mrl)
            ch = oracle.reflect_refract(scene, origin, pp["dest"], direc, pw, meas, pp["n1"], pp["n2"], pp["hit_mesh"],
                                        pp["hit_tri"], env, mrl)
            m = meas == 1
            recs.append(np.concatenate((pp["dest"][m, :3], pw[m, None]), axis=1))
            lins.append(lineage[m])
            per_iter.append(int(m.sum()))
            keep = np.concatenate((ch["refl_meas"], ch["refr_meas"])) == 0
            # lineage rule of the step kernel: refracted children set bit (32 + iteration within the epoch)
            lineage = np.concatenate((lineage, lineage | (np.int64(1) << (32 + it - epoch_start))))[keep]
            origin = np.ascontiguousarray(np.concatenate((ch["refl_origin"], ch["refr_origin"]))[keep])
            direc = np.ascontiguousarray(np.concatenate((ch["refl_dir"], ch["refr_dir"]))[keep])
            pw = np.ascontiguousarray(np.concatenate((ch["refl_power"], ch["refr_power"]))[keep])
            prev = np.ascontiguousarray(np.concatenate((pp["hit_mesh"], pp["hit_mesh"]))[keep])
        else:
            per_iter.append(0)
        live, power_left = dist.all_reduce_pair(len(origin), float(np.sum(pw.astype(np.float64))), group, "cpu")
        lives.append(live)
        if power_left < (1.0 - params["trace_until_dissipated"]) * float(np.asarray(input_power).reshape(-1)[0]):
            break
        if live == 0:
            break
    rec = torch.from_numpy(np.concatenate(recs).astype(np.float32)) if recs else torch.zeros((0, 4))
    lin = torch.from_numpy(np.concatenate(lins)) if lins else torch.zeros(0, dtype=torch.int64)
    merged = dist.gather_measured(rec, lin, per_iter, group, "cpu")
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), merged=merged.numpy(), lives=np.array(lives))
    td.destroy_process_group()


@pytest.mark.parametrize("scene_name,kw", [("c2_directivity_lens", dict(ray_count=301)),
                                           ("c1_nested_cubes", dict(ray_count=400)),
                                           ("c4_human_eye", dict(ray_count=40, sources=2, trace_iterations=5)),
                                           # 40 generations: past the 31 branch bits of a lineage id (renumbering)
                                           ("glass_cube_cavity", dict(ray_count=61, trace_iterations=40)),
                                           # fewer rays than ranks: rank 0 traces nothing and still gathers
                                           ("glass_cube_cavity", dict(ray_count=1, trace_iterations=6))])
def test_two_rank_sharded_trace_equals_single_process(tmp_path, scene_name, kw):
    sys.path.insert(0, ROOT)
    from lightpycl_b200 import scenes
    from oracle import oracle

    mp.spawn(_sharded_oracle_trace, args=(2, _free_port(), scene_name, kw, str(tmp_path)), nprocs=2, join=True)
    ls, meshes, params = scenes.CONFIGS[scene_name](**kw)
    ref = oracle.trace(ls, meshes, power_sum="f64", **params)
    pos, pwr = oracle.measured_rays(ref["results"])
    for r in range(2):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        assert list(z["lives"]) == ref["live"]  # same global termination decisions
        if pwr is None or not np.asarray(pwr).size:
            assert len(z["merged"]) == 0
            continue
        # identical arrays in the reference order on every rank
        assert np.array_equal(z["merged"][:, :3].view(np.uint32), np.ascontiguousarray(pos[:, :3]).view(np.uint32))
        assert np.array_equal(z["merged"][:, 3].copy().view(np.uint32),
                              np.ascontiguousarray(np.asarray(pwr).reshape(-1)).view(np.uint32))


def test_renumber_lineage_single_process_is_dense_order():
    from lightpycl_b200 import dist

    keys = torch.tensor([3, 7 | (1 << 40), 9 | (1 << 62)], dtype=torch.int64)
    assert dist.renumber_lineage(keys, None, "cpu").tolist() == [0, 1, 2]
    assert dist.renumber_lineage(keys[:0], None, "cpu").tolist() == []


def test_merge_measured_orders_by_iteration_then_lineage():
    from lightpycl_b200 import dist

    rec = torch.arange(24, dtype=torch.float32).reshape(6, 4)
    lin = torch.tensor([5, 1 | (1 << 33), 3, 0, 9, 2], dtype=torch.int64)
    itr = torch.tensor([1, 1, 0, 1, 0, 0], dtype=torch.int64)
    m, l, i = dist.merge_measured(rec, lin, itr)
    assert i.tolist() == [0, 0, 0, 1, 1, 1]
    assert l.tolist() == [2, 3, 9, 0, 5, 1 | (1 << 33)]
    assert m[:, 0].tolist() == [20.0, 8.0, 16.0, 12.0, 0.0, 4.0]
