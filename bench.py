#!/usr/bin/env python
"""bench.py -- HIRL hot path on B200: agent updates/sec (BASELINE.json configs[1] by default).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload soft256|fixed4096|linear4096]

One "step" = one learn() call (reference hirl/agents/HIRL.py:221-334) on one synthetic batch.
  value   updates/s with index/noise streams already resident in HBM: ONE persistent launch runs
          the K timed steps (CUDA events on the launching stream, max over ranks)
  e2e     updates/s through the public drop-in call Agent.learn(): host RNG draws, H2D of the
          step's indices from pinned memory, the launch, D2H of the step's scalars -- every step
  --impl reference   the UNMODIFIED reference hirl.agents.HIRL.Agent.learn (files staged under oracle/_ref by
          `make -C oracle`) on the box's host cores, CUDA hidden; the oracle port only as a loud fallback
Every timed figure is the MEDIAN of several repeats of the K-step region (each repeat bracketed by a barrier +
synchronize on both sides and reduced with MAX over ranks).  The default line also carries sub-records for the other
BASELINE configs (`configs.fixed4096`, `configs.ensemble128`) and the env pass on this rank's shard (`roofline_env`).
N>1: independent agent replicas (different seeds), one per GPU, no collective (weak scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torchrun exports OMP_NUM_THREADS=1; the CPU legs (reference arm, cpu_baseline) run on rank 0 only and
# are meant to use every host core, so undo that before torch / MKL read the environment.
if os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = str(os.cpu_count() or 1)
# stdout carries exactly ONE JSON line: everything else that libraries print (e.g. NCCL's version banner)
# goes to stderr until the result is ready.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) from committed ncu captures, per unit of work so that it
# scales to the launch this run times.  Update kernels (profiles/r2_ncu_dram_steady.csv: the same persistent launch with
# 12 and with 36 calls, so the first touch of the 19 MB working set separates from the steady per-call traffic):
#   tensor-core cluster kernel (round 2 default at batch <= 256): 21.68 MB @ 12 calls, 26.69 MB @ 36 calls
#       -> 19.17 MB once per launch + 0.209 MB per call (gradient partials and operand images; activations stay on the SMs);
#   stage kernel (round 1): 69.32 MB @ 12, 193.67 MB @ 36 -> 7.14 MB once + 5.18 MB per call (scratch draining from L2).
# env kernel (profiles/r1_ncu_full_env_step_kernel_v8_raw.csv): 640.1 MB read + 557.7 MB written for 10 M states -> 119.8 B per
# state (algorithmic 124 B).
NCU_TRAFFIC_SOURCE = "profiles/r2_ncu_dram_steady.csv (update kernels: once-per-launch + per-call bytes) / r1_ncu_full_env_step_kernel_v8_raw.csv, scaled to this launch"
NCU_UPDATE_DRAM_BYTES = {("soft256", "tcc"): (19.17e6, 0.2088e6), ("soft256", "stage"): (7.14e6, 5.181e6)}   # (per launch, per call)
NCU_ENV_DRAM_BYTES_PER_STATE = (640.081664e6 + 557.742080e6) / 10_000_000


WORKLOADS = {
    # BASELINE.json configs[1] / SURVEY 8(d): soft every call, 64 expert rows per batch
    "soft256": dict(batch=256, bcw=100.0, expert_num=64, flop_per_row=3_783_424, desc="HIRL-soft fused update, batch 256"),
    "fixed4096": dict(batch=4096, bcw=0.5, expert_num=1024, flop_per_row=3_511_040, desc="HIRL fixed BC weight, batch 4096"),
    "linear4096": dict(batch=4096, bcw=0.5 - 3 / 5000, expert_num=0, flop_per_row=3_511_040, desc="HIRL linear BC schedule, batch 4096"),
}
WORKLOADS["ensemble128"] = dict(batch=256, bcw=100.0, expert_num=64, flop_per_row=3_783_424, replicas=128,
                                desc="128 independent HIRL-soft agent replicas per GPU (1024 on 8 GPUs), ensemble kernel, batch 256")
N_ONLINE, N_EXPERT = 30_000, 20_000
HP = dict(actorLR=1e-3, criticLR=1e-3, tau=0.005, gamma=0.99, bufferSize=10 ** 5)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return d, "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def cpu_baseline(wl, calls, warm=6, seed=0):
    """The reference's learn() as restated in oracle/ (torch CPU, all host threads)."""
    from oracle import scenarios
    from oracle.hirl_oracle import OracleAgent
    import contextlib, io
    sc = scenarios.make_scenario(N_ONLINE, N_EXPERT)
    with contextlib.redirect_stdout(io.StringIO()):
        ag = scenarios.build_agent(OracleAgent, sc, wl["batch"], seed)
    scenarios.seed_all(1)
    for _ in range(warm):
        ag.learn(wl["bcw"], wl["expert_num"], 0)
    t0 = time.perf_counter()
    for _ in range(calls):
        ag.learn(wl["bcw"], wl["expert_num"], 0)
    dt = time.perf_counter() - t0
    return calls / dt, dt


def reference_cpu(wl, calls, repeats=3, warm=6):
    """The unmodified reference's learn() on the host cores (subprocess, CUDA hidden).  Returns (updates/s, info dict) or
    None when oracle/_ref has not been staged (then the caller falls back to the oracle port, and says so)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref", "hirl", "agents", "HIRL.py")
    if not os.path.exists(ref_dir):
        return None
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS=str(os.cpu_count() or 1), MKL_NUM_THREADS=str(os.cpu_count() or 1))
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), "--batch", str(wl["batch"]), "--bcw", str(wl["bcw"]),
           "--expert", str(wl["expert_num"]), "--calls", str(calls), "--repeats", str(repeats), "--warm", str(warm),
           "--online", str(N_ONLINE), "--experts", str(N_EXPERT)]
    t0 = time.perf_counter()
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    if res.returncode != 0:
        sys.stderr.write("reference arm failed, falling back to the oracle port:\n" + res.stderr[-2000:] + "\n")
        return None
    info = json.loads(res.stdout.strip().splitlines()[-1])
    info["wall_s"] = time.perf_counter() - t0
    return info["updates_per_s"], info


def cpu_arm(wl, calls, repeats=3):
    """(value, cores, kind, sample) of the CPU baseline: the unmodified reference when staged, else the oracle port."""
    got = reference_cpu(wl, calls, repeats)
    if got is not None:
        val, info = got
        return val, info["threads"], "reference", (f"unmodified hirl/agents/HIRL.py learn() (oracle/_ref), CUDA hidden, {info['threads']} torch threads on "
                                                    f"{info['nproc']} host CPUs: median of {repeats} x {calls} consecutive calls (whole 6-call cycles) "
                                                    f"after 6 warm-up calls; repeats {['%.1f' % r for r in info['all']]} updates/s")
    sys.stderr.write("WARNING: oracle/_ref is not staged (run `make -C oracle` where /root/reference exists): timing the oracle PORT instead\n")
    torch.set_num_threads(os.cpu_count() or 1)
    val, dt = cpu_baseline(wl, calls)
    return val, torch.get_num_threads(), "port", f"{calls} learn() calls of the oracle port (oracle/hirl_oracle.py) after 6 warm-up calls ({dt:.1f} s)"


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    # >= 300 calls at batch 256 (>= 30 at 4096), whole 6-call cycles, median of 3 repeats;
    # bounded so that the arm ends within a few minutes whatever --steps says
    calls = 300 if wl["batch"] <= 256 else 36
    val, cores, kind, sample = cpu_arm(wl, calls, repeats=3)
    line = {"impl": "reference", "metric": "agent updates/sec", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, wl),
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "ONE CPU process on rank 0 whatever --gpus says: at N > 1 compare it with the per-GPU rate of the other arm"}
    emit(line)


def config_of(args, wl):
    return {"workload": f"{args.workload}: {wl['desc']}, stateDim 13 / actionDim 4 / hidden 256-512 / LayerNorm, "
                        f"{N_ONLINE} online + {N_EXPERT} expert rows, bc_weight_now={wl['bcw']}, expert_num_now={wl['expert_num']}",
            "batch": wl["batch"], "parallelism": f"replicas x{args.gpus} (independent agents, no collective)",
            "l2": "hot by design: the agent state (8.9 MB) and its replay tables stay L2/HBM-resident across steps of the "
                  "persistent kernel, as in production use; the env pass (roofline_env) streams 1.24 GB >> L2"}


def build_agent(wl, device, seed, max_steps):
    from hirl4ucav_b200 import synth
    from hirl4ucav_b200.agents.HIRL import Agent
    from hirl4ucav_b200.utils.seed import set_seed
    import contextlib, io
    es, ea = synth.expert_table(N_EXPERT)
    set_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ag = Agent(HP["actorLR"], HP["criticLR"], 13, 4, 256, 512, HP["tau"], HP["gamma"], HP["bufferSize"], wl["batch"], True,
                   "Harfang_GYM", es, ea, 0.5, True, device=device, max_fused_steps=max_steps)
    synth.fill_agent(ag, N_ONLINE, expert_states=es, expert_actions=ea)
    ag.buffer.flush(); ag.expert_buffer.flush()
    set_seed(seed + 1)
    return ag


def device_resident_run(ag, wl, n_steps):
    """Pre-draws n_steps calls, puts the streams in HBM and returns a launcher for ONE persistent launch."""
    from hirl4ucav_b200 import _lib
    for k in range(n_steps):
        ag._draw(k, wl["bcw"], wl["expert_num"], 0, None)
    dev = ag.device
    steps_d = torch.from_numpy(ag._st_steps_t.numpy()[: n_steps * 64].copy()).to(dev)
    on_d = ag._st_on_t[:n_steps].to(dev); ex_d = ag._st_ex_t[:n_steps].to(dev); bc_d = ag._st_bc_t[:n_steps].to(dev)
    out_d = torch.zeros((n_steps, 8), dtype=torch.float32, device=dev)
    lib = _lib.load()

    def launch():
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.hirl_learn_dev(ag._h, n_steps, steps_d.data_ptr(), on_d.data_ptr(), ex_d.data_ptr(), bc_d.data_ptr(),
                                      out_d.data_ptr(), stream))
    launch.keep = (steps_d, on_d, ex_d, bc_d, out_d)
    return launch, out_d


def env_roofline(device, peaks, rank, world):
    """The env pass on this rank's contiguous shard of the 10 M synthetic pursuit states (BASELINE config 5)."""
    from hirl4ucav_b200 import synth, parallel
    from hirl4ucav_b200.environments import harfang_math as hm
    n_total = 10_000_000
    a, b = parallel.shard_range(n_total, world, rank)
    n = b - a
    raw = synth.raw_states(np.random.default_rng(5), 1 << 20)
    d_raw = torch.from_numpy(np.tile(raw, (n // (1 << 20) + 1, 1))[:n].copy()).to(device)
    for _ in range(3):
        hm.step_math(d_raw)
    torch.cuda.synchronize()
    reps, inner = 7, 20
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all(world)
        e0.record()
        for _ in range(inner):
            hm.step_math(d_raw)
        e1.record()
        sync_all(world)
        times.append(e0.elapsed_time(e1) / inner)
    ms = median_over_repeats(times, device, world)
    gbs = n * 124 / (ms * 1e-3) / 1e9
    return {"kernel": "env_step_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": NCU_ENV_DRAM_BYTES_PER_STATE * n, "traffic_source": NCU_TRAFFIC_SOURCE, "states_per_gpu": n, "states_total": n_total,
            "ms_per_pass": ms, "evals_per_s": n_total / (ms * 1e-3), "achieved_is": "per GPU (its shard); evals_per_s is the whole job over all GPUs",
            "note": f"{n_total} synthetic pursuit states in contiguous shards over {world} GPU(s), no exchange; 124 algorithmic B/state (64 in + 60 out), "
                    f"{n * 124 / 1e6:.0f} MB per pass per GPU >> L2; median of {reps} x {inner} back-to-back passes, max over ranks; includes torch output allocation"}


def sync_all(world):
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def median_over_repeats(times, dev, world):
    """times: this rank's per-repeat figures -> MAX over ranks per repeat, then the median repeat."""
    t = torch.tensor(times, dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.median().item())


def n_repeats(K, per_step_ms):
    """Enough repeats of the K-step region to make the median robust (>= 0.15 s of timed work), 3..25."""
    return int(max(3, min(25, np.ceil(150.0 / max(K * per_step_ms, 1e-3)))))


def tf32_dense_peak(dev):
    """Dense TF32 matmul throughput (torch.matmul fp32, allow_tf32, 8192^3), same burst protocol as MEASURED_PEAKS.json's bf16 figure:
    the denominator of the 3xTF32 tensor-core paths (divide by 3: every fp32-accurate product costs three TF32 MMAs)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device=dev, dtype=torch.float32)
        b = torch.randn((n, n), device=dev, dtype=torch.float32)
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b)
        e1.record(); torch.cuda.synchronize()
        return 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def time_single_agent(name, K, W, local, rank, world, fp32_peak, tf32_peak, peaks, peak_src, with_fused):
    """Device-resident and end-to-end timing of one single-agent workload; returns the record fields."""
    wl = WORKLOADS[name]
    dev = torch.device("cuda", local)
    ag = build_agent(wl, local, seed=rank, max_steps=max(K, W, 64))
    backend = ag.backend
    B = wl["batch"]
    guess = 0.12 if B <= 256 else 0.5
    # ---- value: device-resident streams, one persistent launch for the K timed steps
    warm, _ = device_resident_run(ag, wl, W)
    warm(); torch.cuda.synchronize()
    timed, out_d = device_resident_run(ag, wl, K)
    reps = n_repeats(K, guess)
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all(world)
        e0.record(); timed(); e1.record()
        sync_all(world)
        times.append(e0.elapsed_time(e1))
    ms_total = median_over_repeats(times, dev, world)
    last = out_d[-1].cpu().numpy()
    assert np.isfinite(last[:4]).all(), f"non-finite losses {last}"
    # ---- e2e: the public per-call API, host buffers, copies inside the timed region
    for _ in range(W):
        ag.learn(wl["bcw"], wl["expert_num"], 0)
    ereps = n_repeats(K, guess * 1.6)
    etimes = []
    for _ in range(ereps):
        sync_all(world)
        t0 = time.perf_counter()
        for _ in range(K):
            ag.learn(wl["bcw"], wl["expert_num"], 0)
        torch.cuda.synchronize()
        etimes.append(time.perf_counter() - t0)
        sync_all(world)
    e2e_s = median_over_repeats(etimes, dev, world)
    rec = {"value": world * K / (ms_total * 1e-3), "unit": "updates/s", "ms_per_step": ms_total / K, "steps": K, "repeats": reps,
           "repeat_spread_ms_per_step": [min(times) / K, max(times) / K],
           "e2e": {"value": world * K / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": 3 * B * 4 + 64, "d2h_bytes_per_step": 32, "repeats": ereps,
                   "api": "Agent.learn() per call (hirl_learn: pinned host index/noise streams -> H2D, one launch, D2H scalars)"}}
    if with_fused:
        # e2e, fused form of the same public API: learn_many in chunks of 60 calls
        chunk = [(wl["bcw"], wl["expert_num"], 0)] * 60
        nchunks = max(1, K // 60)
        ag.learn_many(chunk)
        ftimes = []
        for _ in range(max(3, ereps // 2)):
            sync_all(world)
            t0 = time.perf_counter()
            for _ in range(nchunks):
                ag.learn_many(chunk)
            torch.cuda.synchronize()
            ftimes.append(time.perf_counter() - t0)
            sync_all(world)
        fused_s = median_over_repeats(ftimes, dev, world)
        rec["e2e_fused"] = {"value": world * nchunks * 60 / fused_s, "unit": "updates/s", "api": "Agent.learn_many(60 calls) per launch, host buffers"}
    flops_per_step = wl["flop_per_row"] * B
    achieved = flops_per_step * K / (ms_total * 1e-3) / 1e12          # per GPU
    if B >= 512:    # every 256<->512 contraction on tcgen05, 3xTF32 (csrc/tc_gemm.cuh)
        rec["roofline"] = {"kernel": "hirl_update_kernel (tcgen05 3xTF32 tiles)", "bound": "tensor", "achieved": achieved, "peak": tf32_peak / 3.0,
                           "unit": "TFLOP/s", "frac": achieved / (tf32_peak / 3.0), "traffic": None, "flop_per_step": flops_per_step,
                           "peak_source": f"in-run dense TF32 matmul (torch.matmul fp32, allow_tf32, 8192^3: {tf32_peak:.0f} TFLOP/s) / 3 -- an fp32-accurate "
                                          "product costs three TF32 MMAs; achieved = algorithmic fp32 FLOP/s per GPU",
                           "frac_of_fp32_fma_peak": achieved / fp32_peak, "frac_of_bf16_tensor_burst": achieved / peaks.get("bf16_tflops", 1660.5)}
    else:
        kname = {"tcc": "hirl_tcc_kernel (cluster-resident: first layers FP32 FMA, 256<->512 contractions tcgen05 3xTF32)",
                 "stage": "hirl_update_kernel"}.get(backend, backend)
        rec["roofline"] = {"kernel": kname, "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                           "frac": achieved / fp32_peak if fp32_peak else None,
                           "traffic": (NCU_UPDATE_DRAM_BYTES[(name, backend)][0] + NCU_UPDATE_DRAM_BYTES[(name, backend)][1] * K)
                                      if (name, backend) in NCU_UPDATE_DRAM_BYTES else None,
                           "traffic_per_call_steady": NCU_UPDATE_DRAM_BYTES.get((name, backend), (None, None))[1],
                           "traffic_source": NCU_TRAFFIC_SOURCE,
                           "frac_of_tf32_over_3": achieved / (tf32_peak / 3.0) if tf32_peak else None,
                           "peak_source": "in-run register-resident FFMA loop (hirl_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                           "flop_per_step": flops_per_step, "frac_of_bf16_tensor_burst": achieved / peaks.get("bf16_tflops", 1660.5),
                           "tensor_peak_source": peak_src}
    rec["backend"] = backend
    del ag
    return rec


def time_data_parallel(K, W, local, rank, world):
    """BASELINE config fixed4096 as ONE agent whose global batch of 4096 rows is cut over the N GPUs: strong
    scaling.  Every rank builds the same agent (same seed, same replay rows); per call the shard gradients are all-reduced with
    NCCL at the reference's two optimizer steps (include/hirl_b200.h).  Timed end to end through Agent.learn_many (host RNG
    draws over the global batch, sharding, H2D, 2-5 launches + 2-4 NCCL calls per call, D2H), wall clock with a device sync,
    max over ranks."""
    from hirl4ucav_b200 import synth
    from hirl4ucav_b200.agents.HIRL import Agent
    from hirl4ucav_b200.utils.seed import set_seed
    import contextlib, io
    wl = WORKLOADS["fixed4096"]
    es, ea = synth.expert_table(N_EXPERT)
    set_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        ag = Agent(HP["actorLR"], HP["criticLR"], 13, 4, 256, 512, HP["tau"], HP["gamma"], HP["bufferSize"], wl["batch"], True,
                   "Harfang_GYM", es, ea, 0.5, True, device=local, max_fused_steps=64, data_parallel=(rank, world))
    synth.fill_agent(ag, N_ONLINE, expert_states=es, expert_actions=ea)
    ag.buffer.flush(); ag.expert_buffer.flush()
    set_seed(1)
    chunk = [(wl["bcw"], wl["expert_num"], 0)] * K
    ag.learn_many([(wl["bcw"], wl["expert_num"], 0)] * W)
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        sync_all(world)
        t0 = time.perf_counter()
        out = ag.learn_many(chunk)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        sync_all(world)
    assert np.isfinite([float(v) for v in out[-1][:4]]).all()
    sec = median_over_repeats(times, torch.device("cuda", local), world)
    # device-resident form (fused variant only: a launch covers whole calls): Kd pre-drawn, pre-sharded calls in HBM, ONE launch per
    # repeat on every rank, CUDA events on the launching stream, max over ranks -- what the kernels + in-kernel all-reduces cost
    # without the host's per-call draws over the global batch and the launch skew between the ranks' Python processes
    dev_ms = None
    if getattr(ag, "dp_variant", None) == "fused":
        from hirl4ucav_b200 import _lib
        Kd = min(K, 48)
        for k in range(Kd):
            ag._draw(k, wl["bcw"], wl["expert_num"], 0, None)
        ag._shard_slots(0, Kd)
        dv = torch.device("cuda", local)
        st_d = torch.from_numpy(ag._sh_steps_t.numpy()[: Kd * 64].copy()).to(dv)
        on_d, ex_d, bc_d = ag._sh_on_t[:Kd].to(dv), ag._sh_ex_t[:Kd].to(dv), ag._sh_bc_t[:Kd].to(dv)
        out_d = torch.zeros((Kd, 8), dtype=torch.float32, device=dv)
        lib = _lib.load()

        def launch():
            _lib.check(lib.hirl_learn_dev(ag._h, Kd, st_d.data_ptr(), on_d.data_ptr(), ex_d.data_ptr(), bc_d.data_ptr(), out_d.data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        launch(); torch.cuda.synchronize()
        dts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync_all(world)
            e0.record(); launch(); e1.record()
            sync_all(world)
            dts.append(e0.elapsed_time(e1))
        dev_ms = median_over_repeats(dts, dv, world) / Kd
    variant = getattr(ag, "dp_variant", None)
    rec = {"workload": "dp4096: ONE HIRL agent (fixed BC weight), global batch 4096 cut into %d shard(s) of %d rows, gradients all-reduced over NVLink" % (world, wl["batch"] // world),
           "value": K / sec, "unit": "updates/s", "ms_per_step": sec / K * 1e3, "steps": K, "scaling": "strong", "global_batch": wl["batch"],
           "rows_per_gpu": wl["batch"] // world, "backend": ag.backend, "variant": variant,
           "device_resident": None if dev_ms is None else {"value": 1e3 / dev_ms, "unit": "updates/s", "ms_per_step": dev_ms,
                                                            "how": "pre-drawn, pre-sharded calls in HBM, one launch per repeat, CUDA events, max over ranks"},
           "collectives_per_call": "critic-only call: all-reduce of 276 482 fp32 gradient + loss; actor call: + filter count + all-reduce of 138 756 fp32 + 4 scalars",
           "how": ("fused: the all-reduces run INSIDE the update kernel (every rank sums the ranks' exchange blocks in rank order with P2P loads over NVLink "
                   "peer memory, CUDA IPC); a call is one launch" if variant == "fused" else
                   "nccl: 2 (critic-only) / 5 (actor call) stage-range launches of hirl_update_kernel with ncclAllReduce between them"),
           "timing": "end to end through Agent.learn_many (host draws + sharding + H2D + launches + all-reduces + D2H), median of 5, max over ranks"}
    del ag
    return rec


def time_ensemble(name, K, W, R, local, rank, world, fp32_peak, tf32_peak):
    """BASELINE config 4: R replicas per GPU in one ensemble launch; a step = every replica makes one learn() call."""
    from hirl4ucav_b200 import synth, parallel
    from hirl4ucav_b200.ensemble import AgentEnsemble
    from hirl4ucav_b200.environments import harfang_math as hm
    wl = WORKLOADS[name]
    dev = torch.device("cuda", local)
    es, ea = synth.expert_table(N_EXPERT)
    erows, _ = hm.expert_scan(es, ea, local)
    seeds = parallel.replica_seeds(0, R * world, world, rank)          # replica i has seed i wherever it runs
    ens = AgentEnsemble(seeds, HP["actorLR"], HP["criticLR"], HP["tau"], HP["gamma"], N_ONLINE, wl["batch"], es, ea,
                        device=local, max_fused_steps=max(K, W))
    rng = np.random.default_rng(77 + rank)
    for r in range(R):
        s, a, s2 = synth.env_states(rng, N_ONLINE), synth.env_actions(rng, N_ONLINE), synth.env_states(rng, N_ONLINE)
        rew, succ, done = hm.offline_reward(s, a, s2, local)
        rows = np.empty((N_ONLINE, 32), np.float32)
        rows[:, 0:13], rows[:, 13:17], rows[:, 17:30], rows[:, 30], rows[:, 31] = s, a, s2, rew, done
        ens.store_rows(r, rows)
        ens.store_rows(r, erows, expert=True)
    ens.reseed(1)
    calls = [(wl["bcw"], wl["expert_num"], 0)]
    ens.draw(calls * W)
    warm, _ = ens.device_streams(W)
    warm(); torch.cuda.synchronize()
    ens.draw(calls * K)
    timed, out_d = ens.device_streams(K)
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all(world)
        e0.record(); timed(); e1.record()
        sync_all(world)
        times.append(e0.elapsed_time(e1))
    ms_total = median_over_repeats(times, dev, world)
    assert torch.isfinite(out_d[:, -1, :4]).all()
    etimes = []
    for _ in range(3):
        sync_all(world)
        t0 = time.perf_counter()
        ens.learn_many(calls * K)
        etimes.append(time.perf_counter() - t0)
        sync_all(world)
    e2e_s = median_over_repeats(etimes, dev, world)
    B = wl["batch"]
    achieved = wl["flop_per_row"] * B * R * K / (ms_total * 1e-3) / 1e12     # per GPU
    hbm = 11.1e6 * R * K / (ms_total * 1e-3) / 1e9
    rec = {"value": world * R * K / (ms_total * 1e-3), "unit": "replica-updates/s", "ms_per_step": ms_total / K, "steps": K, "repeats": 3,
           "replicas_per_gpu": R, "replicas_total": R * world, "parallelism": f"{R} replicas/GPU x {world} GPUs, teams of CTAs, no collective",
           "e2e": {"value": world * R * K / e2e_s, "unit": "replica-updates/s", "h2d_bytes_per_step": R * (3 * B * 4 + 64), "d2h_bytes_per_step": R * 32,
                   "api": "AgentEnsemble.learn_many (host RNG streams of every replica + H2D + launch + D2H)"},
           "roofline": {"kernel": "hirl_ensemble_kernel (tcgen05 3xTF32 tiles)", "bound": "tensor", "achieved": achieved, "peak": tf32_peak / 3.0, "unit": "TFLOP/s",
                        "frac": achieved / (tf32_peak / 3.0), "traffic": None, "frac_of_fp32_fma_peak": achieved / fp32_peak,
                        "hbm_stream_gbs_if_weights_streamed": hbm,
                        "peak_source": f"in-run dense TF32 matmul ({tf32_peak:.0f} TFLOP/s) / 3; achieved = algorithmic fp32 FLOP/s per GPU",
                        "note": "weights + Adam state of 128 replicas = 1.1 GB > L2: 11.1 MB algorithmic HBM bytes per replica-update"}}
    del ens
    return rec


def run_ours(args):
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    from hirl4ucav_b200 import _lib
    lib = _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    peaks, peak_src = measured_peaks()
    fp32 = C.c_double(0)
    _lib.check(lib.hirl_measure_fp32_peak(local, C.byref(fp32)))
    tf32 = tf32_dense_peak(dev)
    sampler = ClockSampler(local)
    name = args.workload
    subs = {}
    with sampler:
        if "replicas" in WORKLOADS[name]:
            Ke = max(6, min(K, 60) // 6 * 6)
            head = time_ensemble(name, Ke, 6, args.replicas or WORKLOADS[name]["replicas"], local, rank, world, fp32.value, tf32)
            K, W = Ke, 6
        else:
            head = time_single_agent(name, K, W, local, rank, world, fp32.value, tf32, peaks, peak_src, with_fused=True)
        if not args.no_subrecords:
            # the other BASELINE configs ride along in the default line (the driver runs bench.py without --workload)
            if name != "fixed4096":
                subs["fixed4096"] = dict(time_single_agent("fixed4096", 12, 6, local, rank, world, fp32.value, tf32, peaks, peak_src, with_fused=False),
                                         workload=WORKLOADS["fixed4096"]["desc"])
            if name != "ensemble128":
                subs["ensemble128"] = dict(time_ensemble("ensemble128", 6, 6, WORKLOADS["ensemble128"]["replicas"], local, rank, world, fp32.value, tf32),
                                           workload=WORKLOADS["ensemble128"]["desc"])
            if name != "fixed4096":
                try:
                    subs["dp4096"] = time_data_parallel(24, 6, local, rank, world)
                except Exception as e:      # never lose the headline line to the optional mode (every rank fails or succeeds alike: the
                    subs["dp4096"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}   # verdicts are all-gathered inside the attach)
        roof_env = env_roofline(dev, peaks, rank, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    wl = WORKLOADS[name]
    line = {"metric": "agent updates/sec", "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args, wl), "clocks": sampler.summary(), "e2e": head["e2e"],
            "gpu_launches": head["repeats"],
            "gpu_launches_note": "value: every repeat is ONE persistent update-kernel launch covering the K timed steps (median repeat reported); e2e: K launches per repeat",
            "repeats": head["repeats"], "repeat_spread_ms_per_step": head.get("repeat_spread_ms_per_step"),
            "roofline": head["roofline"], "roofline_env": roof_env}
    if "e2e_fused" in head:
        line["e2e_fused"] = head["e2e_fused"]
    if subs:
        line["configs"] = subs
    B = wl["batch"]
    calls = 300 if B <= 256 else 36
    val, cores, kind, sample = cpu_arm(wl, calls, repeats=3)
    line["cpu_baseline"] = {"value": val, "unit": "updates/s", "cores": cores, "kind": kind, "sample": sample,
                            "note": "one CPU process (rank 0): compare with value / n_gpus at N > 1"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6000)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="soft256", choices=sorted(WORKLOADS))
    ap.add_argument("--replicas", type=int, default=0, help="replicas per GPU for the ensemble workload")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the configs.fixed4096 / configs.ensemble128 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
