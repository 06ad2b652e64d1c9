#!/usr/bin/env python
"""Headline benchmark: SFC audio-seconds processed per second, large (24/24) + 24 adapters,
batch 14 x 20 s windows, bf16 operands / fp32 accumulate, synthetic 16 kHz audio, random-init
weights of that architecture (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of 14 windows (280 audio-seconds) per GPU.
Prints ONE JSON line on rank 0 (see the driver contract in DESIGN.md §Measurement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH = 14
WIN_SAMPLES = 320_000
WIN_SEC = 20.0
T_FRAMES = 999          # frames of one 20 s window
CONV_T = [63999, 31999, 15999, 7999, 3999, 1999, 999]
CONV_K = [10, 3, 3, 3, 3, 2, 2]


def work_model(keep=24, adapters=24, D=1024, F=4096, A=512, HF=2048, T=T_FRAMES):
    """algorithmic FLOPs per 20 s window (2 FLOP / MAC; bias, LN, GELU, softmax excluded) —
    BASELINE.md §3 / SURVEY.md §8d. Returns (total, per-kernel-family dict)."""
    fl = {}
    fl["conv0"] = 2 * 1 * 512 * CONV_K[0] * CONV_T[0]
    fl["gemm.conv_k3"] = sum(2 * 512 * 512 * 3 * CONV_T[l] for l in (1, 2, 3, 4))
    fl["gemm.conv_k2"] = sum(2 * 512 * 512 * 2 * CONV_T[l] for l in (5, 6))
    fl["gemm.feat_proj"] = 2 * T * 512 * D
    fl["gemm.pos_conv"] = 2 * T * D * 64 * 128
    fl["gemm.qkv"] = keep * 2 * T * D * 3 * D
    fl["gemm.attn_out"] = keep * 2 * T * D * D
    fl["attention_d64"] = keep * 4 * T * T * D
    fl["gemm.ffn_up"] = keep * 2 * T * D * F + adapters * 2 * T * D * A
    fl["gemm.ffn_down"] = keep * 2 * T * F * D + adapters * 2 * T * A * D
    fl["gemm.head"] = 8 * T * D * D + 4 * T * D * HF
    fl["attention_d128"] = 4 * T * T * D
    fl["head_final"] = 2 * T * D
    return sum(fl.values()), fl


def workload_config(world: int) -> dict:
    """the `config` object of the JSON line: identical for the native and the reference arm"""
    return {"workload": "large (24/24) + 24 FFN adapters SFC inference, batch 14 x 20 s windows per GPU per step",
            "global_batch_windows": world * BATCH, "window_samples": WIN_SAMPLES,
            "frames_per_window": T_FRAMES,
            "parallelism": f"dp{world} (windows sharded, NCCL all_gather of probability rows)" if world > 1 else "single GPU",
            "l2": "inputs rotate over 4 batches; per-step working set ~2.1 GB >> 126 MB L2",
            "weights": "random-init, seed 0"}


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_sustained": d.get("bf16_tflops_sustained"), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "sm_max_mhz": d.get("sm_max_mhz"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for nme, v in zip(names, r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    from wav2vecsegmenter_b200 import _native as nat
    from wav2vecsegmenter_b200 import synth
    from wav2vecsegmenter_b200.engine import SFCEngine

    # NCCL prints its version banner on stdout when the first communicator is created; the
    # contract is ONE JSON line on stdout, so stdout is pointed at stderr until warm-up is over.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    spec = synth.LARGE_ALL
    eng = SFCEngine(spec, dev)
    eng.load_state_dict(synth.random_state_dict(spec, seed=0))
    lib = eng.lib

    # synthetic audio resident in HBM: several distinct batches, rotated, so no step re-reads the
    # previous step's input; the per-step working set (~2.1 GB of activations) is >> 126 MB L2.
    n_rot = 4
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    audio = [torch.randn(BATCH, WIN_SAMPLES, device=dev, generator=g) * 0.1 for _ in range(n_rot)]
    lens = torch.full((BATCH,), WIN_SAMPLES, dtype=torch.int32, device=dev)
    out_len = torch.full((BATCH,), T_FRAMES, dtype=torch.int32, device=dev)
    R = eng.frame_stride(WIN_SAMPLES)
    logits = torch.empty(BATCH, R, device=dev)
    probs = torch.empty(BATCH, R, device=dev)
    # N > 1: the path's only exchange (per-frame probability rows to every rank) runs on its own
    # stream, double-buffered, so the gather of step i overlaps the forward of step i+1 and the ranks
    # are not re-synchronised by a collective between every two steps (the timed region still ends
    # only when the last gather has completed on every rank).
    serial_gather = os.environ.get("W2VSEG_BENCH_SERIAL_GATHER") == "1"   # A/B: gather on the compute stream
    comm = torch.cuda.Stream(dev) if world > 1 else None
    probs2 = [probs, torch.empty_like(probs)]
    gathered2 = [torch.empty(world * BATCH, R, device=dev) for _ in range(2)] if world > 1 else None
    gather_done = [None, None]

    def step(i):
        k = i & 1
        main = torch.cuda.current_stream(dev)
        if world > 1 and gather_done[k] is not None:
            main.wait_event(gather_done[k])          # the gather that read probs2[k] two steps ago
        eng.sfc_forward(audio[i % n_rot], lens, lens, out_len, WIN_SAMPLES, logits, probs2[k])
        if world > 1:
            fwd = torch.cuda.Event()
            fwd.record(main)
            if serial_gather:
                dist.all_gather_into_tensor(gathered2[k], probs2[k])
                return
            with torch.cuda.stream(comm):
                comm.wait_event(fwd)
                dist.all_gather_into_tensor(gathered2[k], probs2[k])
                ev = torch.cuda.Event()
                ev.record(comm)
            gather_done[k] = ev

    def drain():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(comm)

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.w2vseg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    drain()      # the last gathers are inside the timed region
    e1.record()
    # effective SM clock the timed steps ran at (40 us probe enqueued straight after the end event, outside the
    # timed region): nvidia-smi keeps reporting the maximum clock while the power cap throttles the sustained run
    probe_mhz = torch.zeros(148, device=dev)
    nat.check(lib.w2vseg_clock_probe(probe_mhz.data_ptr(), probe_mhz.numel(), 40, nat.current_stream_ptr()),
              "w2vseg_clock_probe")
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.w2vseg_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["effective_sm_mhz"] = round(float(probe_mhz.median()), 1)
        clocks["effective_note"] = ("clock64 per %globaltimer of a 40 us probe kernel enqueued right after the last timed "
                                    "step: the clock the power cap held the sustained run at")
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    audio_sec = world * BATCH * WIN_SEC * args.steps
    value = audio_sec / (ms / 1e3)

    # ---- end to end through the public host API (the call a user makes): 280 s talks held in pinned
    # HOST memory -> TalkRunner.run_stream(): per talk (= one 14-window step) window plan, H2D, fused
    # forward, scatter / NaN fill / tiling average on the device, D2H of the per-frame probabilities.
    # Every step's copies are inside the timed region; run_stream overlaps the H2D of talk k+1 and the
    # D2H of talk k-1 with the forward of talk k (2-deep pipeline). Wall clock between synchronizes.
    from wav2vecsegmenter_b200.pipeline import TalkRunner

    runner = TalkRunner(eng, batch_size=BATCH, segment_sec=WIN_SEC, inference_times=1)
    talks = [(torch.randn(BATCH * WIN_SAMPLES, generator=torch.Generator().manual_seed(7 + i)) * 0.1)
             .pin_memory().numpy() for i in range(2)]

    def e2e_run(n):
        last = None
        for res in runner.run_stream((talks[i % 2] for i in range(n)), depth=2):
            last = res
        return last

    e2e_run(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    last = e2e_run(args.steps)
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = audio_sec / (float(t.item()) / 1e3)
    e2e_d2h = int(last.probs.nbytes + sum(x.nbytes for x in last.per_tiling))

    # ---- BASELINE.json configs[2]: STRONG scaling. A fixed job — 10 h of synthetic talk-length audio
    # (40 talks x 900 s, 1 800 windows), pinned host samples in, per-frame probabilities of every talk on
    # rank 0's host out — through TalkRunner.run(): windows sharded over the ranks, ONE NCCL all_gather of
    # the probability rows, rank 0 assembles the talks. Wall clock, max over ranks. The ideal for N ranks is
    # measured in the same run: every rank processes 1/N of the job on its own (no exchange, no assembly of
    # the other ranks' talks); efficiency = that time / the sharded job's time.
    config3 = None
    if not args.no_strong:
        hours = float(os.environ.get("W2VSEG_BENCH_STRONG_HOURS", "10"))
        n_talks = 40
        talk_n = int(round(hours * 3600 / n_talks * 16000))
        distinct = [(torch.randn(talk_n, generator=torch.Generator().manual_seed(31 + i)) * 0.1).pin_memory().numpy()
                    for i in range(4)]
        job = [distinct[i % 4] for i in range(n_talks)]
        grp = dist.group.WORLD if world > 1 else None
        sharded = TalkRunner(eng, batch_size=BATCH, segment_sec=WIN_SEC, inference_times=1, dist_group=grp)
        alone = TalkRunner(eng, batch_size=BATCH, segment_sec=WIN_SEC, inference_times=1)

        def wall(fn):
            barrier()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()), out

        sharded.run(job[: max(2, world)], results_on=0)                      # warm-up: allocations, NCCL channels
        t_job, res = min((wall(lambda: sharded.run(job, results_on=0)) for _ in range(2)), key=lambda x: x[0])
        share = job[: n_talks // world] if n_talks % world == 0 else job[: n_talks // world + 1]
        t_share = t_one = t_job
        if world > 1:
            alone.run(share[:1])
            t_share, _ = min((wall(lambda: alone.run(share)) for _ in range(2)), key=lambda x: x[0])
            # T_1: the WHOLE job on one GPU (rank 0; the other ranks wait), same run, same box
            t_one, _ = min((wall(lambda: alone.run(job) if rank == 0 else None) for _ in range(2)), key=lambda x: x[0])
        audio_s = n_talks * talk_n / 16000
        config3 = {"workload": f"{hours:g} h of synthetic talks ({n_talks} x {talk_n / 16000:.0f} s, "
                               f"{n_talks * ((talk_n + WIN_SAMPLES - 1) // WIN_SAMPLES)} windows), strong scaling over {world} GPU(s)",
                   "audio_s_per_s": round(audio_s / t_job, 1), "wall_s": round(t_job, 4),
                   "n1_wall_s": round(t_one, 4), "efficiency_vs_n1": round(t_one / (world * t_job), 4),
                   "share_alone_wall_s": round(t_share, 4), "efficiency_vs_share_alone": round(t_share / t_job, 4),
                   "efficiency_definition": "efficiency_vs_n1 = T_1 / (N * T_N): T_1 = the whole job on rank 0's GPU alone, T_N = the sharded job, "
                                            "same run, wall clock, max over ranks; share_alone = one rank processing 1/N of the job with no exchange",
                   "frames_out": int(sum(len(r.probs) for r in res)) if res is not None else None,
                   "api": "wav2vecsegmenter_b200.pipeline.TalkRunner.run(waves, results_on=0)"}

    # ---- per-kernel device timing (CUDA events around every launch, on the launching stream)
    roofline, kernels = None, None
    cpu_baseline = None
    if rank == 0:
        lib.w2vseg_profile_enable(1)
        prof_steps = min(3, args.steps)
        for i in range(prof_steps):
            eng.sfc_forward(audio[i % n_rot], lens, lens, out_len, WIN_SAMPLES, logits, probs)
        buf = nat.C.create_string_buffer(1 << 16)
        lib.w2vseg_profile_collect(buf, len(buf))
        lib.w2vseg_profile_enable(0)
        kernels = {}
        for line in buf.value.decode().splitlines():
            nme, cnt, tot = line.split()
            kernels[nme] = {"launches_per_step": int(cnt) / prof_steps, "ms_per_step": float(tot) / prof_steps}
        total_fl, fl = work_model()
        peaks = read_peaks()
        gemm_names = [k for k in kernels if k.startswith("gemm.") and k != "gemm.pos_conv"]
        gemm_ms = sum(kernels[k]["ms_per_step"] for k in gemm_names)
        gemm_fl = sum(fl[k] for k in gemm_names) * BATCH
        step_ms = sum(k["ms_per_step"] for k in kernels.values())
        # Launch families grouped by the kernel FUNCTION that runs them; the roofline block reports the function
        # with the largest measured share of the step, and lists every tensor-bound family (incl. attention,
        # whose second bound is the XU pipe: 16 ex2 / clk / SM) with its own fraction so that the most
        # flattering family cannot hide the others.
        functions = {
            "gemm_tc2_kernel": gemm_names,
            "attention_tc64_kernel": [k for k in kernels if k == "attention_d64"],
            "attention_tc_kernel<128>": [k for k in kernels if k == "attention_d128"],
            "posconv_tc_kernel": [k for k in kernels if k == "gemm.pos_conv"],
            "layernorm kernels": [k for k in kernels if k in ("layernorm", "ln_gelu.conv")],
            "conv0_tc_kernel": [k for k in kernels if k == "conv0_ln_gelu"],
        }
        fn_ms = {f: sum(kernels[k]["ms_per_step"] for k in ks) for f, ks in functions.items() if ks}
        dom_fn = max(fn_ms, key=fn_ms.get)
        sm_hz = ((clocks or {}).get("effective_sm_mhz") or (clocks or {}).get("sm_mhz")
                 or (peaks.get("sm_max_mhz") or 1965.0))
        families = {}
        for k in sorted((k for k in kernels if k in fl and kernels[k]["ms_per_step"] > 0 and k != "head_final"),
                        key=lambda k: -kernels[k]["ms_per_step"]):
            tf = fl[k] * BATCH / (kernels[k]["ms_per_step"] / 1e3) / 1e12
            families[k] = {"ms_per_step": round(kernels[k]["ms_per_step"], 4), "tflops": round(tf, 1),
                           "frac": round(tf / peaks["tflops_sustained"], 4),
                           "share_of_step": round(kernels[k]["ms_per_step"] / step_ms, 4)}
        if "attention_d64" in families:
            # exponentials per step / (16 per clock per SM x 148 SMs x the SM clock sampled during the timed region)
            n_exp = 24 * BATCH * 16 * T_FRAMES * T_FRAMES
            xu_peak = 16 * 148 * sm_hz * 1e6
            families["attention_d64"]["xu_frac"] = round(n_exp / (kernels["attention_d64"]["ms_per_step"] / 1e3) / xu_peak, 4)
            families["attention_d64"]["xu_note"] = f"MUFU.EX2 issue rate at the effective {sm_hz:.0f} MHz (probe kernel); a 128x128x64 tile costs 1024 XU vs 512 tensor cycles"
        dom_ks = functions[dom_fn]
        dom_ms = fn_ms[dom_fn]
        dom_n = sum(kernels[k]["launches_per_step"] for k in dom_ks)
        dom_fl = sum(fl[k] for k in dom_ks) * BATCH
        achieved = dom_fl / (dom_ms / 1e3) / 1e12
        traffic = None
        tp = ROOT / "profiles" / "gemm_traffic.json"
        if tp.exists() and dom_fn == "gemm_tc2_kernel":
            traffic = json.loads(tp.read_text()).get("kernels", {}).get("gemm.ffn_up", {}).get("dram_bytes_per_launch")
        roofline = {
            "kernel": f"{dom_fn} (all {int(dom_n)} launches per step: " + ", ".join(sorted(dom_ks)) + "; tcgen05 cta_group::2 / TMEM / TMA)",
            "selection": "kernel function with the largest measured time per step (CUDA events around every launch)",
            "bound": "tensor", "achieved": round(achieved, 1), "peak": peaks["tflops_sustained"],
            "unit": "TFLOP/s", "frac": round(achieved / peaks["tflops_sustained"], 4),
            "peak_source": peaks["source"] + ", sustained (kernel timed inside a long step)",
            "traffic": traffic, "traffic_note": "dram bytes of ONE FFN-up launch (ncu --set full, profiles/gemm_traffic.json); algorithmic 167 MB",
            "launches_per_step": dom_n, "avg_launch_ms": round(dom_ms / max(dom_n, 1), 4),
            "algorithmic_gflop_per_launch": round(dom_fl / dom_n / 1e9, 1),
            "share_of_step": round(dom_ms / step_ms, 4),
            "families": families,
            "function_ms_per_step": {f: round(v, 4) for f, v in sorted(fn_ms.items(), key=lambda kv: -kv[1])},
            "whole_path_frac": round(value / world * (total_fl / WIN_SEC) / 1e12 / peaks["tflops_sustained"], 4),
        }
        # memory-bound kernels: ALGORITHMIC bytes per step (each tensor read / written once)
        # over the measured copy bandwidth of MEASURED_PEAKS.json
        M_rows = BATCH * ((WIN_SAMPLES + 319) // 320)          # rows per [B*R, C] activation
        hbm_bytes = {
            "window_stats": 4 * WIN_SAMPLES * BATCH,
            "conv0_ln_gelu": 4 * WIN_SAMPLES * BATCH + 2 * 512 * 64 * M_rows,
            "ln_gelu.conv": (2 + 2) * 512 * 63 * M_rows,        # conv levels 1..6: 32+16+..+1 = 63 x M rows
            "layernorm": 50 * (4 + 2) * 1024 * M_rows + (2 + 2) * 512 * M_rows,
            "cast_to_padded": (4 + 2) * 1024 * M_rows,
            "head_final": 4 * 1024 * M_rows,
        }
        for k, v in kernels.items():
            if k in fl and v["ms_per_step"] > 0:
                v["tflops"] = round(fl[k] * BATCH / (v["ms_per_step"] / 1e3) / 1e12, 1)
            if k in hbm_bytes and v["ms_per_step"] > 0 and peaks.get("hbm_gbs"):
                v["hbm_gbs"] = round(hbm_bytes[k] / (v["ms_per_step"] / 1e3) / 1e9, 1)
                v["hbm_frac"] = round(v["hbm_gbs"] / peaks["hbm_gbs"], 3)
            v["ms_per_step"] = round(v["ms_per_step"], 4)
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline = cpu_reference_timing(n_windows=4, reps=2)   # ~10-15 s of CPU work

    if rank == 0:
        line = {
            "metric": "SFC audio-sec/sec (large 24/24)", "value": round(value, 1), "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": round(e2e_value, 1), "unit": "audio-s/s",
                    "h2d_bytes_per_step": BATCH * WIN_SAMPLES * 4, "d2h_bytes_per_step": e2e_d2h,
                    "api": "wav2vecsegmenter_b200.pipeline.TalkRunner.run_stream (host waves in, per-frame probabilities out; H2D of talk k+1 and D2H of talk k-1 overlap the forward of talk k)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "config3": config3,
            "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_reference_timing(n_windows=1, reps=2):
    """the reference's CPU path for the same model/config, timed on this box's host cores.
    The reference itself (pure Python on HF transformers) cannot travel to the GPU box, so this
    is its restatement (oracle/sfc_oracle.py, torch fp32 CPU kernels == what the reference runs),
    kind "port"."""
    import torch

    from oracle import sfc_oracle
    from wav2vecsegmenter_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = synth.LARGE_ALL
    sd = synth.random_state_dict(spec, seed=0)
    g = torch.Generator().manual_seed(99)
    audio = torch.randn(n_windows, WIN_SAMPLES, generator=g)
    audio = sfc_oracle.normalize_rows(audio, [True] * n_windows)
    out_mask = torch.ones(n_windows, T_FRAMES, dtype=torch.bool)
    times = []
    with torch.no_grad():
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            sfc_oracle.batch_probs(sd, audio, [WIN_SAMPLES] * n_windows, out_mask, spec.keep_layers, spec.head_heads)
            times.append(time.perf_counter() - t0)
    best = min(times[1:])
    return {"value": round(n_windows * WIN_SEC / best, 2), "unit": "audio-s/s", "cores": cores,
            "kind": "port",
            "sample": f"{n_windows} x 20 s window(s) of the same large(24/24)+adapters model, fp32 torch CPU, "
                      f"best of {reps} after 1 warm-up ({best:.2f} s per pass)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import sfc_oracle
    from wav2vecsegmenter_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = synth.LARGE_ALL
    sd = synth.random_state_dict(spec, seed=0)
    g = torch.Generator().manual_seed(99)
    n_win = BATCH   # one step = the full 14-window batch of the native arm (same config)
    audio = sfc_oracle.normalize_rows(torch.randn(n_win, WIN_SAMPLES, generator=g), [True] * n_win)
    out_mask = torch.ones(n_win, T_FRAMES, dtype=torch.bool)
    budget_s = float(os.environ.get("W2VSEG_REF_BUDGET_S", "150"))   # keeps the whole run within a few minutes

    def step():
        with torch.no_grad():
            sfc_oracle.batch_probs(sd, audio, [WIN_SAMPLES] * n_win, out_mask, spec.keep_layers, spec.head_heads)

    warm = 0
    t_w = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        step()
        warm += 1
    per_step = (time.perf_counter() - t_w) / max(warm, 1) if warm else 15.0
    steps = 0
    t0 = time.perf_counter()
    while steps < args.steps and (steps == 0 or (time.perf_counter() - t0) + per_step <= budget_s):
        step()
        steps += 1
    dt = time.perf_counter() - t0
    value = n_win * WIN_SEC * steps / dt
    sample = (f"each step = the full batch of {n_win} x 20 s windows, fp32 torch CPU kernels (what the reference runs), "
              f"{cores} threads; {steps} timed step(s) of --steps {args.steps} (wall budget {budget_s:.0f} s), {warm} warm-up")
    line = {
        "impl": "reference", "metric": "SFC audio-sec/sec (large 24/24)", "value": round(value, 2),
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": round(dt / steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(max(1, int(os.environ.get("WORLD_SIZE", "1")))),
        "cpu_baseline": {"value": round(value, 2), "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the 10 h strong-scaling leg (config3)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
