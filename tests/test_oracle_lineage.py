"""The oracle's lineage ids and the ray-independence property the full-size GPU parity test relies
on (tests/test_gpu_fullsize.py): the records of a subset of the source rays inside a full trace
equal a trace of just that subset, matched by lineage and in the same order (cl:247-249: a ray reads
only its own state; it:366-373: pruning keeps the parent order)."""

import numpy as np
import pytest

import util
from oracle import oracle


@pytest.mark.parametrize("scene", ["c1_cubes_300", "mini_lens", "mini_lossy"])
def test_subset_trace_equals_subset_of_full_trace(scene):
    ls, meshes, params = util.mini_scenes()[scene]
    origin, direc, power, _ = oracle.flatten_sources(ls)
    n = len(origin)
    full = oracle.trace(ls, meshes, power_sum="f64", lineage0=np.arange(n, dtype=np.int64),
                        **dict(params, trace_until_dissipated=1.0))
    n_iter = len(full["results"])
    sample = np.sort(np.random.RandomState(5).choice(n, size=n // 3, replace=False)).astype(np.int64)
    sub = util.ArraySource(origin[sample], direc[sample], power.reshape(n, -1)[sample].reshape(
        (len(sample),) + power.shape[1:]))
    part = oracle.trace([sub], meshes, power_sum="f64", lineage0=sample,
                        **dict(params, trace_iterations=n_iter, trace_until_dissipated=1.0))
    picked = np.zeros(n, dtype=bool)
    picked[sample] = True
    compared = 0
    for k in range(len(part["results"])):
        lin = full["lineage"][k]
        assert np.all(np.diff(lin) > 0)  # lineage order == reference order inside a generation
        idx = np.nonzero(picked[lin & 0xFFFFFFFF])[0]
        assert np.array_equal(lin[idx], part["lineage"][k])
        for a, b, what in zip(full["results"][k], part["results"][k], ("origin", "dest", "power", "measured")):
            a = np.asarray(a).reshape(len(lin), -1)[idx]
            b = np.asarray(b).reshape(len(idx), -1)
            assert np.array_equal(util.bits(a), util.bits(b)), (scene, k, what)
        for key in ("hit_mesh", "hit_tri", "n1", "n2"):
            assert np.array_equal(full["debug"][k][key][idx], part["debug"][k][key]), (scene, k, key)
        compared += len(idx)
    assert compared > len(sample)
