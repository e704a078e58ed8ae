"""Generate the golden fixtures in tests/golden/ from the REFERENCE kernels.

Runs only in the authoring container: it drives oracle/_ref/libref_strict.so, i.e. the
unmodified /root/reference/kernel_reflect_refract_intersect.cl compiled by path (strict IEEE),
through oracle.trace() (the restatement of the reference host loop).  Each fixture freezes the
INPUT arrays (rays + flattened scene, float32) together with the reference's outputs per
iteration, so the fixtures do not depend on regenerating inputs bit-identically elsewhere.

    python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle  # noqa: E402
import util  # noqa: E402


def main():
    assert oracle.available("reference"), "needs oracle/_ref (make -C oracle ref) - authoring container only"
    for name, (ls, meshes, params) in util.mini_scenes().items():
        z = util.scene_to_arrays(ls, meshes)
        out = oracle.trace(ls, meshes, backend="reference", **params)
        z["n_iter"] = np.int32(len(out["results"]))
        z["input_power"] = np.float32(np.asarray(out["input_power"]).reshape(-1)[0])
        for k, v in params.items():
            z["param_" + k] = np.asarray(v)
        for k, (o, d, p, m) in enumerate(out["results"]):
            z["it%d_origin" % k], z["it%d_dest" % k], z["it%d_power" % k], z["it%d_meas" % k] = o, d, p, m
            for key in ("hit_mesh", "hit_tri", "n1", "n2", "near_t", "hits", "near_tri"):
                z["it%d_%s" % (k, key)] = out["debug"][k][key]
        pos, pwr = oracle.measured_rays(out["results"])
        z["measured_pos"] = pos if pos is not None else np.zeros((0, 4), np.float32)
        z["measured_pwr"] = np.asarray(pwr).reshape(-1) if pwr is not None else np.zeros(0, np.float32)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **z)
        print("%-14s tris %5d rays/iter %s measured %d -> %s (%.0f KB)" % (
            name, len(z["v0"]), [len(r[0]) for r in out["results"]], len(z["measured_pwr"]), os.path.basename(path),
            os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
