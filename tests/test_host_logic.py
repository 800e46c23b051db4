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
    from pathlib import Path
    st = importlib.import_module(PKG_NAME + ".style_transfer")
    a = vars(st.build_parser().parse_args([]))
    # every flag of style_transfer.py:126-205 of the reference with its default; nima_weight is the one
    # deliberate deviation (reference 1e5; the NIMA term is not built, so 0)
    ref = dict(content_image="blanc.jpg", style_image="bear.jpeg", output_image="result.jpg", dtype="float32", init="content",
               iter=1000, similarity_metric="li", content_weight=1, style_weight=1e2, regularization_weight=1e4,
               adam_lr=1e-1, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-08, matting_epsilon=1e-5,
               matting_window_radius=3, semantic_thresh=0.5, logs_dir=Path("logs"), results_dir=Path("experiments"),
               seg_dir=Path("raw_seg"), gpu="0", experiment_name=None, intermediate_result_interval=20, print_loss_interval=10)
    for k, v in ref.items():
        assert a[k] == v, k
    assert a["nima_weight"] == 0
    assert set(a) - set(ref) == {"nima_weight", "vgg_weights", "matting", "use_masks", "tv_weight"}     # the extensions
    assert st.CONTENT_LAYERS == ["block4_conv2"] and st.STYLE_LAYERS == ["block%d_conv1" % i for i in range(1, 6)]
    with pytest.raises(SystemExit):
        st.main(["--nima_weight", "1e5"])


def test_script_helpers_follow_the_reference(tmp_path):
    """tensor_to_image truncation (style_transfer.py:78), change_filename (:82-94), meta.json keys and order (:97-120),
    load_image = decode + convert_image_dtype (:50-59)."""
    import json
    import cv2
    import torch
    st = importlib.import_module(PKG_NAME + ".style_transfer")
    t = torch.tensor([[[[0.0, 0.999 / 255, 1.0 / 255], [0.5, 254.999 / 255, 1.0]]]], dtype=torch.float32)
    u8 = st.tensor_to_image(t)
    assert u8.dtype == torch.uint8 and tuple(u8.shape) == (1, 2, 3)
    assert u8.flatten().tolist() == [0, 0, 1, 127, 254, 255]                 # truncation, not rounding
    assert st.change_filename('.', 'image.png', '_seg') == './image_seg.png'
    assert st.change_filename('raw_seg', 'blanc.jpg', '_seg', '.png') == 'raw_seg/blanc_seg.png'
    args = st.build_parser().parse_args(["--iter", "7", "-c", "a.png", "-s", "b.png"])
    meta = st.write_metadata(args, True, tmp_path)
    on_disk = json.loads((tmp_path / "meta.json").read_text())
    assert on_disk == meta
    assert list(meta) == ["init", "iter", "content", "style", "content_weight", "style_weight", "regularization_weight",
                          "nima_weight", "semantic_thresh", "similarity_metric", "load_segmentation", "adam"]
    assert list(meta["adam"]) == ["learning_rate", "beta1", "beta2", "epsilon"] and meta["iter"] == 7
    rgb = np.random.default_rng(0).integers(0, 256, (5, 7, 3), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "x.png"), rgb[:, :, ::-1])
    img = st.load_image(tmp_path / "x.png", "float32")
    assert img.shape == (1, 5, 7, 3) and img.dtype == np.float32
    assert np.array_equal(img[0], rgb.astype(np.float32) * np.float32(1.0 / 255.0))
    st.save_image(torch.as_tensor(rgb), tmp_path / "y.png")
    assert np.array_equal(cv2.imread(str(tmp_path / "y.png"))[:, :, ::-1], rgb)
    with pytest.raises(FileNotFoundError):
        st.load_image(tmp_path / "missing.png")


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
    vgg = importlib.import_module(PKG_NAME + ".components.VGG19.model")
    assert tiled.LEVEL_HALO == (4, 4, 8, 4, 2)
    t = tiled.Tile(3840, 3, 8)                                                  # an interior strip of a 3840-px image
    assert (t.own_lo, t.own_hi, t.ext_lo, t.ext_hi, t.local_w) == (1440, 1920, 1436, 1924, 488)
    assert t.widths == (488, 248, 136, 68, 34) and t.pool_xoff == (2, 6, 0, 0)
    assert t.own_cols(488) == (4, 484) and t.own_cols(136) == (8, 128) and t.own_cols(34) == (2, 32)
    assert t.global_cols(34) == 240 and t.halo_cols(34) == 2 and t.halo_cols(248) == 4
    assert (t.mask_lo, t.mask_hi) == (1408, 1952)                               # widest footprint: 8 << 2 = 4 << 3 = 2 << 4 = 32 px
    assert t.mask_cols(488) == (28, 516) and t.mask_cols(136) == (0, 544) and t.mask_cols(248) == (24, 520)
    t0 = tiled.Tile(3840, 0, 8)                                                 # the left edge: no halo on the left
    assert (t0.ext_lo, t0.ext_hi, t0.widths, t0.pool_xoff) == (0, 484, (484, 244, 128, 64, 32), (0, 0, 0, 0))
    assert t0.own_cols(484) == (0, 480) and t0.mask_cols(244) == (0, 488)
    t1 = tiled.Tile(3840, 0, 1)                                                 # a single strip: the plain geometry
    assert t1.geom() is None and t1.widths == (3840, 1920, 960, 480, 240) and t1.own_cols(960) == (0, 960)
    with pytest.raises(ValueError):
        tiled.Tile(1000, 0, 8)
    with pytest.raises(ValueError):
        tiled.Tile(3840, 3, 8, halos=(4, 4, 6, 4, 2))                           # block3 has four convolutions: needs 8
    with pytest.raises(ValueError):
        t.own_cols(100)
    # every pixel is owned by exactly one rank; on every level the halo covers twice the number of convolutions of each segment
    # (forward validity shrinks by one column per convolution, the backward pass needs the ReLU masks that many columns out),
    # and the un-pooled gradient (2 x the pooled level's halo) reaches as far as the segment below needs it
    owned = np.zeros(3840, int)
    for r in range(8):
        tr = tiled.Tile(3840, r, 8)
        owned[tr.own_lo:tr.own_hi] += 1
        for l in range(4):                                                      # the pool of level l fits inside level l + 1
            assert tr.pool_xoff[l] >= 0 and tr.pool_xoff[l] + tr.widths[l] // 2 <= tr.widths[l + 1]
            assert (tr.own_lo >> l) % 2 == 0 and tr.left[l] % 2 == 0           # 2x2 windows line up with the global grid
            lo_next, _ = tr.own_cols(tr.widths[l + 1])
            assert tr.pool_xoff[l] + tr.own_cols(tr.widths[l])[0] // 2 == lo_next
    assert (owned == 1).all()
    for first, last in vgg.SEGMENTS:
        level = sum(1 for p in vgg.POOL_AFTER if p < first)
        n = last - first + 1
        assert tiled.LEVEL_HALO[level] >= 2 * n, (first, last)
        if last in vgg.POOL_AFTER:
            assert 2 * tiled.LEVEL_HALO[level + 1] >= n, (first, last)
    # the tensors exchanged in one step of a 3840x2160 run (interior strip): five on the way up, their gradients on the way
    # down, the image; the mailbox of tiled.PeerHalo is sized from this list
    plan = tiled.exchange_plan(2160, t, 12)
    up = [(1080, 4, 64), (540, 8, 128), (270, 4, 256), (270, 4, 512), (135, 2, 512)]
    assert plan == up + up[::-1] + [(2160, 4, 3)]
    assert all((hl * c) % 4 == 0 for _, hl, c in plan)                         # float4 rows in the push / pull kernels
    assert sum(r * hl * c * 4 for r, hl, c in plan) < 32 << 20


_TILED_WORKER = r"""
import os, sys, json, importlib
sys.path.insert(0, %r)
import torch, torch.distributed as dist
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
tiled = importlib.import_module(%r + ".tiled")
H, W = 6, 16 * 8 * world                       # strips of 128 px
comm = tiled.GlooComm(rank, world)
t = tiled.Tile(W, rank, world)
ok = True
for level, C in ((0, 3), (2, 5), (4, 2)):      # the image, a block-3 input, a block-5 input
    f = 1 << level
    w_l, gw = t.widths[level], W // f
    e_lo, e_hi = t.own_lo // f - t.left[level], t.own_hi // f + t.right[level]      # this level's columns, global coordinates
    glob = torch.arange(H * gw * C, dtype=torch.float32).reshape(1, H, gw, C)
    loc = glob[:, :, e_lo:e_hi].clone()
    lo, hi = t.own_cols(w_l)
    hl = t.halo_cols(w_l)
    loc[0, :, :lo] = -1.0                      # stale halos
    loc[0, :, hi:] = -1.0
    loc[0, :, lo:hi] += 1000.0 * (rank + 1)    # "updated" own columns
    recv = comm.exchange(loc, lo, hi, hl)
    expect = glob[:, :, e_lo:e_hi].clone()
    for r in range(world):
        a, b = max(r * gw // world, e_lo), min((r + 1) * gw // world, e_hi)
        if a < b:
            expect[0, :, a - e_lo:b - e_lo] += 1000.0 * (r + 1)
    ok = ok and bool(torch.equal(loc, expect)) and len(recv) == (rank > 0) + (rank < world - 1)
    ok = ok and hl == tiled.LEVEL_HALO[level] and (hi - lo) == W // world // f and loc.shape[2] == w_l
# the Gram partials: ONE flat float32 buffer + the float64 accumulator, summed over the ranks
flat = torch.full((1000,), float(rank + 1)); acc = torch.tensor([1.0, 0.0, 2.0 * rank, 0.5], dtype=torch.float64)
comm.reduce_sum([flat, acc])
ok_sum = bool((flat == world * (world + 1) / 2).all()) and float(acc[0]) == world and float(acc[2]) == world * (world - 1)
# configs[2] sharding of bench.py: pair p goes to rank p %% world
mine = list(range(rank, 64, world))
allp = [None] * world
dist.all_gather_object(allp, mine)
res = [None] * world
dist.all_gather_object(res, (ok, ok_sum, comm.bytes["exchanges"]))
if rank == 0:
    print(json.dumps({"results": res, "pairs": sorted(sum(allp, []))}))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 3])
def test_tiled_exchange_protocol_gloo(world, tmp_path):
    """The exchange code of tiled.py over a gloo group (CPU tensors): after an exchange every halo column of the image, of a
    block-3 input and of a block-5 input holds the owner's updated value (middle ranks talk to both sides), the flattened
    Gram partials and the float64 accumulator are summed, and the configs[2] round-robin covers all 64 pairs exactly once."""
    import json
    script = tmp_path / "tiled_worker.py"
    script.write_text(_TILED_WORKER % (ROOT, PKG_NAME))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29540 + world), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    d = json.loads(outs[0][0].strip().splitlines()[-1])
    assert d["pairs"] == list(range(64))
    for r, (ok, ok_sum, n) in enumerate(d["results"]):
        assert ok and ok_sum and n == 3, r


def test_bench_deadline_prints_the_line_it_has(tmp_path):
    """bench.py's guard around the multi-GPU tiled record: when the deadline passes, the rank that owns the line prints it
    with an error entry in place of tiled_4k and the process ends with status 0 (the other ranks: status 5, no output)."""
    import json
    code = ("import sys, time; sys.path.insert(0, %r); import bench\n"
            "line = {'metric': 'adam_iters_per_sec_1024x1024', 'value': 1.0} if sys.argv[1] == '0' else None\n"
            "g = bench._Deadline(0.3, line)\n"
            "time.sleep(30)\n") % ROOT
    script = tmp_path / "deadline.py"
    script.write_text(code)
    r0 = subprocess.run([sys.executable, str(script), "0"], capture_output=True, text=True, timeout=120)
    assert r0.returncode == 0, r0.stderr
    d = json.loads(r0.stdout.strip().splitlines()[-1])
    assert d["value"] == 1.0 and "not finished" in d["tiled_4k"]["error"]
    r1 = subprocess.run([sys.executable, str(script), "1"], capture_output=True, text=True, timeout=120)
    assert r1.returncode == 5 and r1.stdout.strip() == ""
    # a cancelled guard does nothing
    code2 = ("import sys, time; sys.path.insert(0, %r); import bench\n"
             "g = bench._Deadline(0.2, {'metric': 'x'}); g.cancel(); time.sleep(0.6); print('alive')\n") % ROOT
    script.write_text(code2)
    r2 = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0 and r2.stdout.strip() == "alive"
