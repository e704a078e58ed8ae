"""Generate tests/golden/refhost_*.npz by RUNNING THE REFERENCE'S OWN PYTHON HOST CODE in this container.

`oracle/ref_host.py` executes /root/reference/{light_source,geo_optical_elements,iterative_tracer}.py (Python-2 source)
by path under a Python-2-semantics shim, with pyopencl replaced by a stand-in whose "kernels" are the reference .cl
compiled by g++ (oracle/_ref).  So every number below comes out of the reference's own statements:

  refhost_sources.npz    light_source: constructor (random_rays), grid_rays, random_collimated_rays, rotate_rays
  refhost_geometry.npz   optical_elements generators + GeoObject transforms/materials: tribuf() digests and samples
  refhost_trace_*.npz    CL_Tracer.iterative_tracer (the host loop it:77-393, partitions included), get_measured_rays,
                         get_binned_data / _stereographic / _angular, replicate_lightsources_and_plot (up to hist_data),
                         get_beam_width_half_power, get_beam_HWHM (their printed numbers)

The scenes are lightpycl_b200.scenes' builders run with the reference's modules swapped in, seeds fixed.
Authoring container only (/root/reference does not travel); the tests read the committed fixtures.

    python tests/golden/make_golden_host.py
"""

import hashlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_host  # noqa: E402
import refhost_cases as cases  # noqa: E402


def digest(a):
    a = np.ascontiguousarray(np.asarray(a))
    return hashlib.sha1(a.tobytes()).hexdigest()


def sources():
    lsm = ref_host.load("light_source")
    z = {}
    with ref_host.captured_stdout():
        for name, make in cases.source_cases(lsm).items():
            s = make()
            z[name + "/origin"] = np.asarray(s.rays_origin)
            z[name + "/dir"] = np.asarray(s.rays_dir)
            z[name + "/power"] = np.asarray(s.rays_power)
            z[name + "/ray_count"] = np.int64(s.ray_count)
    np.savez_compressed(os.path.join(HERE, "refhost_sources.npz"), **z)
    print("refhost_sources.npz: %d cases" % (len(z) // 4))


def geometry():
    goe = ref_host.load("geo_optical_elements")
    z = {}
    with ref_host.captured_stdout():
        for name, make in cases.geometry_cases(goe).items():
            m = make()
            v0, v1, v2 = [np.asarray(b, dtype=np.float32) for b in m.tribuf()]
            z[name + "/count"] = np.int64(len(v0))
            z[name + "/sha1"] = np.array([digest(v0), digest(v1), digest(v2)])
            z[name + "/head"] = np.stack([v0[:4], v1[:4], v2[:4]])
            z[name + "/tail"] = np.stack([v0[-4:], v1[-4:], v2[-4:]])
            mat = m.getMaterialBuf()
            z[name + "/material"] = np.array([mat["type"], mat["IOR"], mat["R"], mat["dissipation"]], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "refhost_geometry.npz"), **z)
    print("refhost_geometry.npz: %d cases" % (len(z) // 5))


_NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?|nan|inf")


def _numbers(text, label):
    for line in text.splitlines():
        if line.startswith(label):
            return np.array([float(x) for x in _NUM.findall(line[len(label):])], dtype=np.float64)
    raise KeyError(label)


def traces():
    for name, case in cases.trace_cases().items():
        with ref_host.reference_scene_modules() as scenes:
            ls, meshes, params = getattr(scenes, case["scene"])(**case["kw"])
        tr = ref_host.reference_tracer()
        with ref_host.captured_stdout():
            res = tr.iterative_tracer(ls, meshes, **params)
        z = {"n_iter": np.int64(len(res)), "rays": np.array([len(r[0]) for r in res], dtype=np.int64),
             "tri_count": np.int64(tr.tri_count)}
        for k, (o, d, p, m) in enumerate(res):
            arrs = dict(origin=o, dest=d, power=p, meas=m)
            for key, a in arrs.items():
                a = np.asarray(a)
                z["it%d_%s_shape" % (k, key)] = np.array(a.shape, dtype=np.int64)
                if case.get("digest_only"):
                    z["it%d_%s_sha1" % (k, key)] = np.array(digest(a))
                else:
                    z["it%d_%s" % (k, key)] = a
        pos, pwr = tr.get_measured_rays()
        n_meas = 0 if pos is None else len(pos)
        z["measured_count"] = np.int64(n_meas)
        if n_meas:
            if case.get("digest_only"):
                z["measured_sha1"] = np.array([digest(pos), digest(pwr)])
            else:
                z["measured_pos"], z["measured_pwr"] = np.asarray(pos), np.asarray(pwr)
        if case.get("post") and n_meas:
            pts, lim_xy, lim_ang = case["points"], case["limits_xy"], ((-1, 1), (-1, 1))
            with ref_host.captured_stdout() as out:
                H, xe, ye = tr.get_binned_data(limits=lim_xy, points=pts)
                z["hist_xy"], z["hist_xy_edges"] = H, np.stack([xe, ye])
                H, xe, ye = tr.get_binned_data_stereographic(limits=lim_ang, points=pts)
                z["hist_stereo"], z["hist_stereo_edges"] = H, np.stack([xe, ye])
                H, xe, ye = tr.get_binned_data_angular(limits=lim_ang, points=pts)
                z["hist_angular"], z["hist_angular_edges"] = H, np.stack([xe, ye])
                tr.hist_data = None
                try:  # the function ends in matplotlib calls (absent here); hist_data is complete before them
                    tr.replicate_lightsources_and_plot(limits=lim_ang, points=pts, axis="z", sources=case["replicas"])
                except Exception:
                    pass
                H, xe, ye = tr.hist_data
                z["hist_replicated"], z["hist_replicated_edges"] = H, np.stack([xe, ye])
                tr.get_beam_width_half_power()
                tr.get_beam_HWHM(points=case["hwhm_points"])
            text = out.getvalue()
            z["metrics_text"] = np.array(text)
            hp, hw = text.index("Throughput results:"), text.index("Throughput results HWHM:")
            z["half_power"] = np.concatenate([_numbers(text[hp:hw], "Total measured power: "),
                                              _numbers(text[hp:hw], "Beam halfpower angle: "),
                                              _numbers(text[hp:hw], "Halfwidth throughput: ")])
            z["hwhm"] = np.concatenate([_numbers(text[hw:], "Total measured power: "),
                                        _numbers(text[hw:], "Beam HWHM angle:      "),
                                        _numbers(text[hw:], "HWHM throughput:      ")])
        path = os.path.join(HERE, "refhost_trace_%s.npz" % name)
        np.savez_compressed(path, **z)
        print("refhost_trace_%s.npz: rays/iter %s measured %d (%.0f KB)" % (name, z["rays"].tolist(), n_meas,
                                                                          os.path.getsize(path) / 1024))


if __name__ == "__main__":
    assert ref_host.available() and ref_host.kernels_available(), "authoring container only (needs /root/reference and oracle/_ref)"
    sources()
    geometry()
    traces()
