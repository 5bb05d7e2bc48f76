"""Shared scene builders for the tests: the same scene as an oracle OScene and as a product Scene."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rings():
    with open(os.path.join(ROOT, "tests", "golden", "example_geojson_rings.json")) as f:
        return json.load(f)["rings"]


def geojson_text():
    feats = [{"type": "Feature", "geometry": {"type": "Polygon", "coordinates": [r]}} for r in rings()]
    return json.dumps({"type": "FeatureCollection", "features": feats})


def oracle_scene_from_product(scene):
    """differt2d_b200.Scene -> oracle.ref_torch.OScene (arrays only)."""
    from oracle import ref_torch as R

    xys, kinds, phis = scene.packed_objects()
    return R.OScene(xys, kinds, phis, {k: v.xy for k, v in scene.transmitters.items()},
                    {k: v.xy for k, v in scene.receivers.items()})


def normalised(scene):
    """Shift/scale a scene to the unit square (well-conditioned variant of the geojson scene)."""
    import differt2d_b200 as d

    bb = scene.bounding_box().astype(np.float64)
    span = float(max(bb[1, 0] - bb[0, 0], bb[1, 1] - bb[0, 1]))

    def f(p):
        return ((np.asarray(p, np.float64) - bb[0]) / span).astype(np.float32)

    objs = []
    for o in scene.objects:
        if isinstance(o, d.Vertex):
            objs.append(d.Vertex(xy=f(o.xy)))
        elif isinstance(o, d.RIS):
            objs.append(d.RIS(xys=f(o.xys), phi=o.phi))
        else:
            objs.append(d.Wall(xys=f(o.xys)))
    return d.Scene({k: d.Point(xy=f(v.xy)) for k, v in scene.transmitters.items()},
                   {k: d.Point(xy=f(v.xy)) for k, v in scene.receivers.items()}, objs)


def jittered_grid(scene, n, m, seed=0, margin=0.02):
    """A grid strictly inside the bounding box with irregular coordinates (avoids the measure-zero
    configurations where the reference's own gradient is NaN: receivers on walls, un == 0, d == 0)."""
    rng = np.random.default_rng(seed)
    bb = scene.bounding_box()
    w = bb[1] - bb[0]
    x = bb[0, 0] + w[0] * (margin + (1 - 2 * margin) * np.sort(rng.random(m)))
    y = bb[0, 1] + w[1] * (margin + (1 - 2 * margin) * np.sort(rng.random(n)))
    X, Y = np.meshgrid(x.astype(np.float32), y.astype(np.float32))
    return X, Y


def generic_position(scene, seed=11, eps=2e-3):
    """Moves every object vertex by a small random offset so that no two walls stay collinear or share
    end-point heights.  Axis-aligned canned scenes contain STRUCTURAL ties (e.g. ta + tb == 1 for a whole
    region of receivers), where min/max sub-gradients w.r.t. the wall vertices depend on last-bit noise in
    the reference itself; parity of d/d(vertices) is only meaningful away from those ties."""
    import differt2d_b200 as d

    rng = np.random.default_rng(seed)
    objs = []
    for o in scene.objects:
        if isinstance(o, d.Vertex):
            objs.append(d.Vertex(xy=o.xy + rng.uniform(-eps, eps, 2).astype(np.float32)))
        elif isinstance(o, d.RIS):
            objs.append(d.RIS(xys=o.xys + rng.uniform(-eps, eps, (2, 2)).astype(np.float32), phi=o.phi))
        else:
            objs.append(d.Wall(xys=o.xys + rng.uniform(-eps, eps, (2, 2)).astype(np.float32)))
    return d.Scene(scene.transmitters, scene.receivers, objs)
