"""CPU: host-side logic of the product package that needs no GPU -- mask helpers (class order, bit-exact against the
reference-generated golden vectors), synthetic inputs, flag defaults, and the N>1 sharding arithmetic of bench.py
exercised over a world_size-2 gloo group."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT, golden


@pytest.mark.parametrize("tag", ["a", "b"])
def test_product_mask_helpers_match_reference_golden(tag):
    sem = importlib.import_module(PKG_NAME + ".components.semantic_merge")
    g = golden("masks_%s.npz" % tag)
    d = sem.extract_segmentation_masks(g["seg"])
    assert np.array_equal(np.array(sorted(d), dtype=np.int64), g["keys"])
    m = np.stack([t.numpy() for t in sem.mask_for_tf(d)])
    assert m.dtype == np.float32 and np.array_equal(m, g["masks"])
    assert np.array_equal(sem.reduce_dict(d, np.zeros((1,) + g["seg"].shape)), g["reduced"])
    merged = sem.replace_colors_in_dict(d, {sorted(d)[0]: sorted(d)[1]})
    assert len(merged) == len(d) - 1 and sum(v.sum() for v in merged.values()) == g["seg"].shape[0] * g["seg"].shape[1]
    with pytest.raises(NotImplementedError):
        sem.merge_segments(None, None, 0.5, "li")


def test_synthetic_inputs_are_deterministic(synth):
    a, b = synth.image(8, 9, 3), synth.image(8, 9, 3)
    assert a.dtype == np.float32 and a.shape == (1, 8, 9, 3) and np.array_equal(a, b)
    lab = synth.label_image(64, 96, 8, 1, cell=16)
    assert lab.shape == (64, 96, 3) and lab.dtype == np.uint8
    assert len(np.unique(lab.reshape(-1, 3), axis=0)) == 8
    w = synth.vgg_weights(seed=1)
    assert w["block1_conv1"][0].shape == (3, 3, 3, 64) and w["block5_conv1"][0].shape == (3, 3, 512, 512)
    s = synth.smooth_image(16, 16, 2)
    assert np.allclose(s * 255, np.round(s * 255), atol=1e-4)


def test_flag_defaults_follow_the_reference():
    st = importlib.import_module(PKG_NAME + ".style_transfer")
    a = st.build_parser().parse_args([])
    # style_transfer.py:143-183 of the reference
    assert (a.iter, a.adam_lr, a.matting_window_radius, a.matting_epsilon) == (1000, 0.1, 3, 1e-5)
    assert (a.content_weight, a.style_weight, a.regularization_weight) == (1, 100, 10 ** 4)
    assert st.CONTENT_LAYERS == ["block4_conv2"] and st.STYLE_LAYERS == ["block%d_conv1" % i for i in range(1, 6)]


_WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
# bench.py's N>1 protocol: independent pairs, one per rank (seeds 2r, 2r+1); barrier; MAX of the elapsed times;
# value = world * steps / max_time.  Exercised here with fake timings.
seeds = (2 * rank, 2 * rank + 1)
ms = torch.tensor([100.0 + 50.0 * rank])
dist.barrier()
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
allseeds = [None] * world
dist.all_gather_object(allseeds, seeds)
if rank == 0:
    print(json.dumps({"max_ms": float(ms), "seeds": allseeds, "value": world * 10 / (float(ms) * 1e-3)}))
dist.destroy_process_group()
"""


def test_multi_rank_protocol_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs)
    import json
    d = json.loads(outs[0].strip().splitlines()[-1])
    assert d["max_ms"] == 150.0 and d["seeds"] == [[0, 1], [2, 3]] and abs(d["value"] - 2 * 10 / 0.15) < 1e-9


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_tile_geometry():
    """Spatial tiling (BASELINE configs[3]): strip / halo arithmetic, CPU only."""
    tiled = importlib.import_module(PKG_NAME + ".tiled")
    t = tiled.Tile(3840, 3, 8)
    assert (t.own_lo, t.own_hi, t.ext_lo, t.ext_hi, t.local_w) == (1440, 1920, 1280, 2080, 800)
    assert t.own_cols(800) == (160, 640) and t.own_cols(50) == (10, 40) and t.global_cols(50) == 240
    t0 = tiled.Tile(3840, 0, 8)
    assert (t0.ext_lo, t0.ext_hi) == (0, 640) and t0.own_cols(640) == (0, 480)
    with pytest.raises(ValueError):
        tiled.Tile(1000, 0, 8)
    # every pixel is owned by exactly one rank and every halo is twice the block5_conv1 receptive-field radius
    owned = np.zeros(3840, int)
    for r in range(8):
        tr = tiled.Tile(3840, r, 8)
        owned[tr.own_lo:tr.own_hi] += 1
        assert tr.own_lo - tr.ext_lo in (0, tiled.HALO) and tr.ext_hi - tr.own_hi in (0, tiled.HALO)
    assert (owned == 1).all() and tiled.HALO >= 2 * 78 and tiled.HALO % 16 == 0
