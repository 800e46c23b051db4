"""Regenerate profiles/README.md (round-2 section first, round-1 text kept below) from the artefacts in profiles/."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def jline(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    return json.loads(open(path).read().strip().splitlines()[-1])


def summary(name, pick):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return {}
    d = {}
    for line in open(path):
        f = line.rstrip("\n").split(",")
        if len(f) >= 3 and f[0] in pick and f[0] not in d:
            d[f[0]] = (f[2], f[1])
    return d


def fmt(d, key, nd=1):
    if key not in d:
        return "n/a"
    v, u = d[key]
    try:
        return ("%." + str(nd) + "f %s") % (float(v), u)
    except ValueError:
        return "%s %s" % (v, u)


out = io.StringIO()
w = out.write
w("# Profiles\n\nAll captures: `ncu --clock-control none` under gpurun, after the same command exited 0 without ncu.  Numbers printed by a\n"
  "run under ncu are never quoted as bench values; the bench lines here come from plain runs.  Per-launch times of an ncu launch\n"
  "list are cold-cache and serialised: compare SHARES, not absolutes (they also differ by a few percent between boxes).\n\n")
w("# Round 2\n\n")
b = jline("r2_bench_1gpu.json")
if b:
    w("## 1. Bench line of the final build -- `r2_bench_1gpu.json` (`python bench.py`, one B200)\n\n")
    w("* **%.1f iter/s** device-timed (%.3f ms/step, %d launches per step), **%.1f iter/s end to end** from pinned host buffers;\n"
      "  clocks: %s.\n" % (b["value"], b["ms_per_step"], b["gpu_launches_per_step"], b["e2e"]["value"], json.dumps(b["clocks"])))
    cb = b.get("cpu_baseline")
    if cb:
        w("* CPU port of the reference step: %.3f iter/s on %d host cores (%s).\n" % (cb["value"], cb["cores"], cb["sample"].split(" (")[0]))
    p = b.get("parity")
    if p:
        w("* `parity` (float64 oracle on the benchmark inputs): worst loss scalar %.2e, gradient %.2e of its max-norm (%.2e outside the\n"
          "  receptive fields of the %d ReLU units of %d whose decision differs); tolerance 1e-5; oracle time %.0f s.\n"
          % (p["max_rel_loss_diff"], p["grad_rel_maxnorm"], p["grad_rel_maxnorm_outside_flipped_relu_fields"], p["relu_flips"],
             p["relu_units"], p["oracle_seconds"]))
    r = b["roofline"]
    w("* `roofline` (3x3 convolutions, 24 launches timed alone, L2 flushed): **%.1f TFLOP/s** algorithmic = %.3f of the measured\n"
      "  sustained bf16 peak (%.0f), %.2f of the peak/3 ceiling of the float32-accurate 3xFP16 arithmetic; share of the step %.2f.\n"
      % (r["achieved"], r["frac"], r["peak"], r["fp32_accurate_ceiling"]["frac"], r["share_of_step"]))
    rl = b["roofline_lx"]
    w("* `roofline_lx` (Laplacian mat-vec at 1024^2, 36 B/px algorithmic): **%.0f GB/s** = %.3f of the measured HBM copy bandwidth,\n"
      "  %.1f us per launch (%s).\n" % (rl["achieved"], rl["frac"], rl["ms_per_launch"] * 1e3, rl["l2"]))
    sw = b.get("roofline_lx_sweep")
    if sw:
        w("\n`roofline_lx_sweep` (configs[4]; v2 / v3 operators, build and mat-vec):\n\n| size | v2 mat-vec us | GB/s | of HBM | v2 build us (48 B/px) | v3 mat-vec us | v3 COO build ms |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for row in sw["rows"]:
            a, c = row["v2"], row["v3"]
            w("| %s | %.1f | %.0f | %.3f | %.1f | %.1f | %s |\n" % (row["size"], a["matvec_ms"] * 1e3, a["matvec_GBps"], a["matvec_frac_hbm"],
                                                                 a.get("build_ms", 0) * 1e3, c["matvec_ms"] * 1e3,
                                                                 ("%.2f" % c["build_ms"]) if "build_ms" in c else "-"))
    pr = b.get("pairs_64x512")
    if pr:
        w("\n`pairs_64x512` (configs[2] on one GPU): %.0f iter/s over all 64 pairs incl. set-up; set-up %.1f ms per pair (+%.1f ms host data\n"
          "generation), %.2f ms per iteration; set-up share of a 100-iteration run %.1f %%.\n"
          % (pr["value"], pr["setup_ms_per_pair"], pr["host_datagen_ms_per_pair"], pr["iteration_ms"], 100 * pr["setup_share_at_100_iterations"]))
    t4 = b.get("tiled_4k")
    if t4:
        w("\n`tiled_4k` (configs[3] on one GPU, the whole 3840x2160 image): %.1f iter/s (%.1f ms/step).\n" % (t4["value"], t4["ms_per_step"]))
q = jline("r2_bench_1gpu_quick_final.json")
if q:
    w("\n`r2_bench_1gpu_quick_final.json` (`python bench.py --no-extras --no-parity --no-cpu-baseline`, the LAST build of the round:\n"
      "staged epilogue stores in the forward and style-gradient convolutions only, section 2's launch list is from the same run):\n"
      "**%.1f iter/s** (%.3f ms/step), %.1f end to end, convolutions %.1f TFLOP/s.  The full line above and `r2_conv_layers_1024.txt`\n"
      "were taken one build earlier (the data-gradient kernels of the 64-channel layers staged their stores as well, which this\n"
      "build undoes because it costs time when the ReLU mask is loaded in the same epilogue, as it is in a train step);\n"
      "boxes of the pool differ by +-2.5 %%.\n" % (q["value"], q["ms_per_step"], q["e2e"]["value"], q["roofline"]["achieved"]))
ref = jline("r2_bench_reference_arm.json")
if ref:
    w("\n`r2_bench_reference_arm.json` (`bench.py --impl reference`): %.3f iter/s on %d threads, kind \"%s\" (%s).\n"
      % (ref["value"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["kind"], ref["note"]))

path = os.path.join(P, "r2_launches_train_step_1024.csv")
if os.path.exists(path):
    w("\n## 2. Launch list of one train step (1024x1024, K = 8, TV on) -- `r2_launches_train_step_1024.csv`\n\n"
      "`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python scripts/profile_step.py --steps 1`\n\n")
    w(subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_table.py"), path], capture_output=True, text=True).stdout)
    w("\n`conv3x3_tc_kernel<BN, MODE>`: MODE 0 = forward (bias, ReLU, fused 2x2 max-pool), 1 = data gradient, 2 = style gradient\n"
      "(classes as taps).\n")

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
w("\n## 3. `ncu --set full` captures (summaries: `r2_prof_*_summary.csv`, made by `scripts/ncu_summary.py`)\n\n")
for name, what in (("r2_prof_conv3_2_summary.csv", "`conv3x3_tc_kernel<128,0>`, block3_conv2 forward (256x256, 256 -> 256)"),
                   ("r2_prof_conv1_2_summary.csv", "`conv3x3_tc_kernel<64,0>`, block1_conv2 forward (1024x1024, 64 -> 64)"),
                   ("r2_prof_lap_dia_summary.csv", "`lap_dia_kernel`, 2048x2048"),
                   ("r2_prof_gram64_summary.csv", "`gram_tc_kernel<64>`, block1_conv1 features (1024x1024x64, K = 8)")):
    d = summary(name, KEYS)
    if not d:
        continue
    w("* %s -- `%s`: %s under ncu; tensor pipe %s of active cycles; issue slots %s; DRAM %s read + %s written (%s of peak);\n"
      "  TMA L2->SM %s; %s warp instructions; %s registers; warps active %s; long-scoreboard stalls per issue %s.\n"
      % (what, name, fmt(d, KEYS[0]), fmt(d, KEYS[1]), fmt(d, KEYS[2]), fmt(d, KEYS[3]), fmt(d, KEYS[4]), fmt(d, KEYS[5]), fmt(d, KEYS[6], 2),
         fmt(d, KEYS[7], 0), fmt(d, KEYS[8], 0), fmt(d, KEYS[9]), fmt(d, KEYS[10], 2)))
w("* per-layer convolution rates of all 24 launches: `r2_conv_layers_1024.txt` (`scripts/conv_layers.py 1024`).\n")
w("* `sass_opcodes.txt`: per-kernel counts of UTCHMMA / UTMALDG / LDTM / STTM / UTCBAR / SYNCS ... in the shipped `libadpst.so`\n"
  "  (`scripts/sass_opcodes.py`).\n")

w("\n## 4. Multi-GPU (plain runs of the default `bench.py` line under torch.distributed.run)\n\n"
  "The 2-, 4- and 8-GPU lines were taken one commit before the last change to the epilogue of the 64-channel convolutions (a few\n"
  "percent on two layers); everything multi-GPU is identical.  Different boxes of the pool differ by about +-2.5 % per kernel.\n\n"
  "| GPUs | headline iter/s (one 1024^2 pair per GPU) | e2e | pairs_64x512 iter/s | set-up ms/pair | tiled_4k iter/s | ms/step | halo transport | comm ms (of which all-reduce) | redundant columns, blocks 1 / 3-5 | parity vs 1 GPU |\n"
  "|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|---:|\n")
for n, name in ((1, "r2_bench_1gpu.json"), (2, "r2_bench_2gpu.json"), (4, "r2_bench_4gpu.json"), (8, "r2_bench_8gpu.json")):
    d = jline(name)
    if not d:
        continue
    pr, t4 = d.get("pairs_64x512", {}), d.get("tiled_4k", {})
    bd = t4.get("breakdown_max_over_ranks", {})
    par = t4.get("parity_vs_single_device", {})
    w("| %d | %.1f | %.1f | %.0f | %.1f | %.1f | %.2f | %s | %s | %s | %s |\n"
      % (n, d["value"], d["e2e"]["value"], pr.get("value", 0), pr.get("setup_ms_per_pair", 0), t4.get("value", 0), t4.get("ms_per_step", 0),
         t4.get("halo_transport", "-"),
         ("%.2f (%.2f)" % (bd["communication_ms"], bd.get("allreduce_ms", float("nan")))) if bd else "-",
         ("%.3f / %.3f" % (bd["redundant_column_factor"], bd.get("redundant_column_factor_blocks345", bd["redundant_column_factor"]))) if bd else "-",
         ("%.1e" % par["max_rel_loss_diff"]) if par else "-"))
w("\n`tiled_4k`: one 3840x2160 image in column strips with 4, 4, 8, 4, 2 halo columns on the five resolution levels; per step eleven halo exchanges between the six network\n"
  "segments (activations up, gradients down, the image border) by our push/pull kernels over NVLink peer memory, and NCCL\n"
  "all-reduces of the Gram partials (started as each partial is complete).  The timed steps replay one CUDA graph per rank.\n"
  "`comm ms` comes from a short EAGER run with CUDA events around every exchange / all-reduce call (work enqueued inside an\n"
  "exchange window subtracted), max over ranks: it contains the time a rank waits for a slower neighbour; `redundant columns` =\n"
  "(own + halo) / own.\n")

old = open(os.path.join(P, "README.md")).read()
marker = "# Round 1 profiles"
tail = old[old.index(marker):] if marker in old else old
open(os.path.join(P, "README.md"), "w").write(out.getvalue() + "\n" + tail)
print(out.getvalue()[:1500])
