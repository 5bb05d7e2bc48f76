"""CPU tests of the host mirror (Scene container, builders, keyword resolution) and of the multi-GPU
plumbing with the gloo backend at world_size 2."""
import json
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import differt2d_b200 as d
from differt2d_b200 import distributed as D
from differt2d_b200 import logic
from oracle import ref_torch as R
from tests import helpers as H


def test_builders_match_oracle_scenes():
    for a, b in [(d.Scene.square_scene(), R.square_scene()), (d.Scene.basic_scene(), R.basic_scene()),
                 (d.Scene.square_scene_with_wall(), R.square_scene_with_wall()),
                 (d.Scene.square_scene_with_obstacle(), R.square_scene_with_obstacle())]:
        xys, kinds, phis = a.packed_objects()
        assert np.array_equal(xys, b.xys.numpy()) and kinds.tolist() == b.kinds
        assert np.array_equal(a.transmitters["tx"].xy, b.transmitters["tx"].numpy())
        assert np.array_equal(a.receivers["rx"].xy, b.receivers["rx"].numpy())


def test_from_geojson(geojson_rings):
    """tests/test_scene.py:217-253 of the reference"""
    sc = d.Scene.from_geojson(H.geojson_text())
    assert len(sc.objects) == 28 and all(isinstance(o, d.Wall) for o in sc.objects)
    osc = R.scene_from_geojson_rings(geojson_rings)
    assert np.array_equal(sc.packed_objects()[0], osc.xys.numpy())
    assert np.array_equal(sc.transmitters["tx"].xy, osc.transmitters["tx"].numpy())
    assert np.array_equal(sc.receivers["rx"].xy, osc.receivers["rx"].numpy())
    empty = d.Scene.from_geojson(json.dumps({"features": []}))
    assert len(empty.objects) == 0 and empty.receivers["rx"].xy.tolist() == [1.0, 1.0]


def test_grid_shapes():
    """tests/test_abc.py:10-29 — grid(m, n) has shape (n, m); grid(m) is square"""
    sc = d.Scene.basic_scene()
    X, Y = sc.grid(300, 50)
    assert X.shape == (50, 300) and Y.shape == (50, 300) and X.dtype == np.float32
    X, Y = sc.grid(25)
    assert X.shape == (25, 25)
    assert np.allclose(sc.center(), [0.5, 0.5])
    assert sc.get_location("NW").tolist() == [0.0, 1.0] and sc.get_location("SE").tolist() == [1.0, 0.0]


def test_candidates_api():
    """tests/test_scene.py:372-399 of the reference, through the product API (host enumerator of the .so)"""
    sc = d.Scene.random_uniform_scene(key=1234, n_receivers=10, n_walls=11)
    got = sc.all_path_candidates(min_order=0, max_order=0, device=None)
    assert len(got) == 1 and len(got[0]) == 0
    sc = d.Scene(objects=[d.Wall(), d.Wall(), d.Wall()]).add_objects(d.RIS(), d.Wall(), d.Wall())
    got = sc.all_path_candidates(filter_objects=lambda o: isinstance(o, d.RIS), min_order=0, max_order=2)
    assert [c.tolist() for c in got] == [[], [3]] and all(c.dtype == np.int32 for c in got)


def test_mode_resolution():
    assert logic.resolve_mode(False) == "hard"
    assert logic.resolve_mode(True) == "hard_sigmoid"  # logic.py:266 default activation
    assert logic.resolve_mode(True, logic.sigmoid) == "sigmoid"
    with logic.enable_approx(True):
        assert logic.resolve_mode(None) == "hard_sigmoid"
    with logic.disable_approx():
        assert logic.resolve_mode(None) == "hard"
    with pytest.raises(NotImplementedError):
        logic.resolve_mode(True, lambda x, a: x)


def test_unsupported_requests_raise():
    sc = d.Scene.square_scene()
    X, Y = sc.grid(4)
    # an arbitrary callable selects the generic escape hatch (paths materialised on the GPU, fun on batched arguments);
    # the flag is RETURNED: no per-call state is kept on the Scene
    assert sc._config("receivers", lambda *a: 0.0, (), None, False, d.ImagePath, None, 0, 1, None, None, {})[2] is True
    assert sc._config("receivers", d.received_power, (), None, False, d.ImagePath, None, 0, 1, None, None, {})[2] is False
    assert not hasattr(sc, "_generic")
    with pytest.raises(NotImplementedError):
        sc._config("receivers", "not a function", (), None, False, d.ImagePath, None, 0, 1, None, None, {})
    cfg3, _, _ = sc._config("receivers", d.received_power, (), None, False, d.MinPath, {"many": 3}, 0, 1, None, None, {})
    assert cfg3.many == 3 and sc._x0(cfg3, 7, None).shape == (5, 3, 1)  # optimize.py:142-182 restarts
    with pytest.raises(NotImplementedError):
        sc._config("receivers", d.received_power, (), None, False, d.MinPath, {"optimizer": object()}, 0, 1, None, None, {})
    with pytest.raises(TypeError):
        sc._config("receivers", d.received_power, (), None, False, d.ImagePath, None, 0, 1, None, None, {"bogus": 1})
    with pytest.raises(TypeError):
        sc.add_objects(d.Vertex(xy=[0.3, 0.3]))._config("receivers", d.received_power, (), None, False, d.ImagePath,
                                                        None, 0, 1, None, None, {})
    cfg, alpha, _ = sc._config("receivers", d.received_power, (), {"r_coef": 0.3, "height": 0.0}, True, d.FermatPath,
                            {"steps": 7}, 0, 1, 2, None, {"approx": True, "alpha": 12.0, "tol": 0.5})
    assert (cfg.min_order, cfg.max_order, cfg.steps, cfg.r_coef, cfg.height, cfg.mode, cfg.tol) == \
           (2, 2, 7, 0.3, 0.0, "hard_sigmoid", 0.5) and alpha == 12.0


def test_nan_parity_keyword_selects_the_gradient_mode():
    sc = d.Scene.basic_scene()
    args = ("receivers", d.received_power, (), None, True, d.ImagePath, None, 0, 1, None, None)
    assert sc._config(*args, {})[0].grad_mode == "clean"
    assert sc._config(*args, {"nan_parity": True})[0].grad_mode == "nan_parity"


def test_sanitised_scene_drops_closure_walls_and_normalises():
    from tests import helpers as H

    sc = d.Scene.from_geojson(H.geojson_text())
    assert len(sc.objects) == 28                                   # tests/test_scene.py:217-238 of the reference
    clean, info = sc.sanitised(return_map=True)
    assert len(clean.objects) == 26 and sorted(set(range(28)) - set(info["kept"])) == [0, 7]
    assert all(np.any(o.xys[0] != o.xys[1]) for o in clean.objects)
    assert np.array_equal(clean.objects[0].xys, sc.objects[1].xys) and info["scale"] == 1.0
    unit, info = sc.sanitised(normalise=True, return_map=True)
    bb = unit.bounding_box()
    assert np.allclose(bb[0], 0.0, atol=1e-6) and abs(bb[1].max() - 1.0) < 1e-6 and bb[1].min() > 0.0
    back = np.asarray(unit.objects[3].xys, np.float64) * info["scale"] + info["origin"]
    assert np.allclose(back, np.asarray(sc.objects[info["kept"][3]].xys, np.float64), atol=1e-9 + 1e-6 * info["scale"])
    assert set(unit.transmitters) == set(sc.transmitters) and set(unit.receivers) == set(sc.receivers)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): exactly ONE line on stdout, valid JSON
    with the contract's keys — native libraries' banners go to stderr (bench._quiet_stdout)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "rx_x_candidate_paths_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and line["higher_is_better"] is True


def test_row_blocks_partition_the_grid():
    for n in (1, 7, 10, 1024, 2048):
        for w in (1, 2, 3, 4, 8):
            blocks = [D.row_block(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_cyclic_row_bands_partition_the_grid():
    for n in (1, 7, 10, 1024, 2048, 8192):
        for w in (1, 2, 3, 4, 8):
            parts = [D.row_tiles_cyclic(n, w, r) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))      # disjoint cover
            for r, p in enumerate(parts):
                assert all((int(i) // 8) % w == r for i in p)                    # whole 8-row bands, round robin
            if n % (8 * w) == 0:
                assert len({len(p) for p in parts}) == 1                         # equal shares


@pytest.mark.parametrize("G,slices,shards", [(1000, 1, 2), (124_500_500, 7, 8), (129, 3, 4), (0, 1, 2), (249_500, 37, 3)])
def test_candidate_shards_partition_the_list(G, slices, shards):
    """SURVEY §8e, point-to-point links: the chunks of 128 candidates are dealt to the candidate shards (one per GPU)
    so that every chunk is traced exactly once, and long lists are split evenly (csrc/d2d_driver.cuh)."""
    per = [D.candidate_chunks_of_shard(G, slices, s, shards) for s in range(shards)]
    allq = np.concatenate(per)
    assert sorted(allq.tolist()) == list(range((G + 127) // 128))
    if G > 128 * slices * shards * 8:
        assert max(len(q) for q in per) - min(len(q) for q in per) <= slices


def test_pack_unpack_roundtrip():
    o, ph, f, a = torch.randn(5, 2, 2), torch.randn(5), torch.randn(3, 2), torch.randn(1)
    buf = D.pack_param_grads(o, ph, f, a)
    assert buf.numel() == 5 * 4 + 5 + 3 * 2 + 1  # 4N + N + 2T + 1 floats
    o2, ph2, f2, a2 = D.unpack_param_grads(buf, 5, 3)
    assert torch.equal(o, o2) and torch.equal(ph, ph2) and torch.equal(f, f2) and torch.equal(a, a2)


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    D.init(world, rank, backend="gloo")
    # every rank owns a row block; parameter cotangents are partial sums -> one all-reduce
    n = 9
    r0, r1 = D.row_block(n, world, rank)
    full = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4)
    part = D.pack_param_grads(full[r0:r1].sum(0).reshape(1, 2, 2), torch.tensor([float(rank)]),
                              torch.tensor([[1.0, 2.0]]) * (rank + 1), torch.tensor([0.5]))
    D.allreduce_sum_(part)
    tmax = D.max_over_ranks(10.0 + rank, "cpu")
    D.barrier()
    if rank == 0:
        torch.save({"buf": part, "tmax": tmax}, out)
    D.shutdown()


def test_gloo_world_size_2_allreduce(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    full = torch.arange(9 * 4, dtype=torch.float32).reshape(9, 4)
    o, ph, f, a = D.unpack_param_grads(res["buf"], 1, 1)
    assert torch.equal(o.reshape(-1), full.sum(0))          # sum over both row blocks == full-grid sum
    assert ph.item() == 1.0 and f.tolist() == [[3.0, 6.0]] and a.item() == 1.0
    assert res["tmax"] == 11.0
