"""Shared helpers for the parity tests."""
import numpy as np

SCENES = ["random", "cornell", "cornell_smoke", "final", "mesh"]
# small mesh detail keeps the oracle's BVH build and the fixtures quick
SCENE_KW = {"mesh": {"mesh_detail": 1}}

_host_cache = {}


def host_scene(rt, name):
    if name not in _host_cache:
        _host_cache[name] = rt.HostScene(name, construction_seed=1, **SCENE_KW.get(name, {}))
    return _host_cache[name]


def random_path_ids(n, width, height, spp, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, width, n, dtype=np.uint32), rng.integers(0, height, n, dtype=np.uint32),
            rng.integers(0, spp, n, dtype=np.uint32))


def rel_err(a, b, floor=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def secondary_rays(hits, rays, seed):
    """Rays leaving the hit points of `hits` in random directions (un-normalised, like the
    reference's scattered rays), for first-hit parity beyond camera rays."""
    rng = np.random.default_rng(seed)
    ok = hits["node"] >= 0
    n = int(ok.sum())
    out = np.zeros(n, dtype=rays.dtype)
    out["origin"] = hits["position"][ok]
    d = rng.normal(size=(n, 3))
    # Leave the surface on the side the incoming ray came from (the reference's normal is already
    # face-forwarded), as every scattered ray that carries light does.  Rays travelling INSIDE
    # closed boxes meet coincident faces of neighbouring boxes at exactly equal t; which one the
    # reference returns then depends on its AABB `t_out <= t_in` rejection (aabb.rs:31) and on
    # the last ulp of x/d versus x*(1/d) (DESIGN.md "Ties").
    nrm = hits["normal"][ok]
    flip = (d * nrm).sum(axis=1) < 0
    d[flip] *= -1.0
    d *= rng.uniform(0.2, 3.0, size=(n, 1))
    out["direction"] = d
    out["time"] = rays["time"][ok]
    return out


def compare_hits(dev, ref):
    """Returns dict of mismatch counts / max errors between device and oracle hits."""
    same_id = (dev["node"] == ref["node"]) & (dev["face"] == ref["face"])
    hit = ref["node"] >= 0
    both = same_id & hit
    t_err = rel_err(dev["t"][both], ref["t"][both])
    n_err = np.abs(dev["normal"][both] - ref["normal"][both]).max(axis=1) if both.any() else np.zeros(0)
    p_err = rel_err(dev["position"][both], ref["position"][both], floor=1.0).max(axis=1) if both.any() else np.zeros(0)
    uv_err = np.maximum(np.abs(dev["u"][both] - ref["u"][both]), np.abs(dev["v"][both] - ref["v"][both])) if both.any() else np.zeros(0)
    return {
        "n": int(ref.shape[0]),
        "hits": int(hit.sum()),
        "id_mismatch": int((~same_id).sum()),
        "front_face_mismatch": int((dev["front_face"][both] != ref["front_face"][both]).sum()),
        "material_mismatch": int((dev["material"][both] != ref["material"][both]).sum()),
        "t_max_rel": float(t_err.max()) if t_err.size else 0.0,
        "normal_max_abs": float(n_err.max()) if n_err.size else 0.0,
        "pos_max_rel": float(p_err.max()) if p_err.size else 0.0,
        "uv_max_abs": float(uv_err.max()) if uv_err.size else 0.0,
    }
