"""
Generates tests/golden/power_fixtures.npz: golden vectors of the hot path on small seeded inputs, produced by the CPU
oracle (oracle/d2d_oracle.c for masks and maps, oracle/ref_torch.py + autograd for the clean VJP, in fp32 and in fp64).  The reference itself
cannot run in this image (no JAX), so these pin the ORACLE: tests/test_oracle_kat.py::test_oracles_reproduce_the_golden_
fixtures checks that both restatements still reproduce them bit for bit, tests/test_gpu_parity.py::test_cuda_path_
against_golden_fixtures checks the CUDA path against the same arrays.
    python tests/golden/make_power_fixtures.py
Inputs: the canned scenes of tests/helpers.py / Appendix B of SURVEY.md, jittered 12 x 10 grids (seed 21), orders 0-2.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "power_fixtures.npz")
SCENE_NAMES = ("obstacle", "basic", "geojson", "geojson_norm")


def scenes():
    import differt2d_b200 as d
    from tests import helpers as H

    gj = d.Scene.from_geojson(H.geojson_text())
    return {"obstacle": d.Scene.square_scene_with_obstacle(), "basic": d.Scene.basic_scene(), "geojson": gj,
            "geojson_norm": H.normalised(gj)}


def inputs(sc):
    from tests import helpers as H

    X, Y = H.jittered_grid(sc, 12, 10, seed=21)
    xys, kinds, phis = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    return X, Y, xys, fixed


def compute(name, sc):
    from oracle import c_oracle as CO
    from oracle import ref_torch as R
    from tests import helpers as H

    X, Y, xys, fixed = inputs(sc)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    out = {f"{name}/X": X, f"{name}/Y": Y, f"{name}/xys": xys, f"{name}/fixed": fixed}
    Z, v = CO.power_map(xys, fixed, grid, max_order=2, mode="hard", want_valid=True)
    out[f"{name}/hard/Z"] = Z
    out[f"{name}/hard/valid_bits"] = np.packbits(v.astype(np.uint8).reshape(-1))
    out[f"{name}/hard/valid_shape"] = np.array(v.shape)
    for alpha in (10.0, 100.0):
        Zs, vs = CO.power_map(xys, fixed, grid, max_order=2, mode="hard_sigmoid", alpha=alpha, want_valid=True)
        out[f"{name}/hard_sigmoid_{alpha:g}/Z"] = Zs
        out[f"{name}/hard_sigmoid_{alpha:g}/valid"] = vs
    # clean VJP (torch autograd over the restated graph), hard_sigmoid alpha = 20, Zbar seeded
    Zbar = np.random.default_rng(22).standard_normal(X.shape).astype(np.float32)
    with R.clean_gradients():
        Zo, g = R.power_map_and_vjp(H.oracle_scene_from_product(sc), X, Y, Zbar, max_order=2, approx=True, alpha=20.0,
                                    function="hard_sigmoid")
    out[f"{name}/vjp/Zbar"] = Zbar
    out[f"{name}/vjp/Z"] = Zo.numpy()
    for k in ("grid", "xys", "fixed", "alpha"):
        out[f"{name}/vjp/{k}_bar"] = g[k].numpy()
    # the same VJP evaluated in binary64 on the same fp32 inputs (ref_torch.precision): where it differs from the
    # fp32 value the quantity is ill-conditioned in fp32 (third leg of the parity triangulation)
    with R.clean_gradients(), R.precision("f64"):
        Z64, g64 = R.power_map_and_vjp(H.oracle_scene_from_product(sc), X, Y, Zbar, max_order=2, approx=True,
                                       alpha=20.0, function="hard_sigmoid")
    out[f"{name}/vjp64/Z"] = Z64.numpy()
    for k in ("grid", "xys", "fixed", "alpha"):
        out[f"{name}/vjp64/{k}_bar"] = g64[k].numpy()
    return out


def main():
    out = {}
    for name, sc in scenes().items():
        out.update(compute(name, sc))
    np.savez_compressed(DST, **out)
    print(DST, os.path.getsize(DST), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
