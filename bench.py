#!/usr/bin/env python
"""bench.py -- HIRL hot path on B200: agent updates/sec (BASELINE.json configs[1] by default).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload soft256|fixed4096|linear4096]

One "step" = one learn() call (reference hirl/agents/HIRL.py:221-334) on one synthetic batch.
  value   updates/s with index/noise streams already resident in HBM: ONE persistent launch runs
          the K timed steps (CUDA events on the launching stream, max over ranks)
  e2e     updates/s through the public drop-in call Agent.learn(): host RNG draws, H2D of the
          step's indices from pinned memory, the launch, D2H of the step's scalars -- every step
  --impl reference   the CPU torch restatement of the reference's learn() (oracle/, pinned
          bit-for-bit to the unmodified reference) on the box's host cores
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

# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` captures
# (profiles/r1_ncu_full_update_kernel_v10_raw.csv, r1_ncu_full_env_step_kernel_v8_raw.csv), per unit of work so that
# it scales to the launch this run times:
# update kernel, soft256: 10.72 MB read + 53.91 MB written in a 12-call launch -> 5.39 MB per call (the agent state
# and its replay tables are L2-resident; the writes are scratch buffers draining from L2);
# env kernel: 640.1 MB read + 557.7 MB written for 10 M states -> 119.8 B per state (algorithmic 124 B).
NCU_TRAFFIC_SOURCE = "profiles/r1_ncu_full_update_kernel_v10_raw.csv / r1_ncu_full_env_step_kernel_v8_raw.csv, scaled to this launch"
NCU_UPDATE_DRAM_BYTES_PER_STEP = {"soft256": (10.721024e6 + 53.908480e6) / 12}
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


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    torch.set_num_threads(os.cpu_count() or 1)
    per_call_guess = 0.012 if wl["batch"] <= 256 else 0.08
    calls = max(6, min(args.steps, int(150 / per_call_guess)))      # bounded: <= ~2.5 min of CPU work
    calls -= calls % 6 if calls >= 12 else 0                           # whole 6-call cycles
    val, dt = cpu_baseline(wl, calls, warm=max(6, min(args.warmup, 60)))
    sample = f"{calls} consecutive learn() calls (whole 6-call cycles) of the same workload, {dt:.1f} s of CPU time"
    line = {"impl": "reference", "metric": "agent updates/sec", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, wl),
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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


def env_roofline(device, peaks):
    from hirl4ucav_b200 import synth
    from hirl4ucav_b200.environments import harfang_math as hm
    n = 10_000_000
    raw = synth.raw_states(np.random.default_rng(5), 1 << 20)
    d_raw = torch.from_numpy(np.tile(raw, (10, 1))[:n].copy()).to(device)
    for _ in range(3):
        hm.step_math(d_raw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        hm.step_math(d_raw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = n * 124 / (ms * 1e-3) / 1e9
    return {"kernel": "env_step_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": NCU_ENV_DRAM_BYTES_PER_STATE * n, "traffic_source": NCU_TRAFFIC_SOURCE, "states": n, "ms_per_pass": ms, "evals_per_s": n / (ms * 1e-3),
            "note": "10 M synthetic pursuit states, 124 algorithmic B/state (64 in + 60 out), 1.24 GB per pass >> L2, includes torch output allocation"}


def run_ensemble(args):
    """BASELINE config 4: R replicas per GPU in one ensemble launch; a step = every replica makes one learn() call."""
    rank, local, world = dist_env()
    from hirl4ucav_b200 import _lib, synth, parallel
    from hirl4ucav_b200.ensemble import AgentEnsemble
    from hirl4ucav_b200.environments import harfang_math as hm
    wl = WORKLOADS[args.workload]
    R = args.replicas or wl["replicas"]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    K, W = min(args.steps, 60), min(max(args.warmup, 3), 12)
    K -= K % 6 if K >= 6 else 0
    peaks, peak_src = measured_peaks()
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
    sampler = ClockSampler(local)
    with sampler:
        ens.draw(calls * W)
        warm, _ = ens.device_streams(W)
        warm(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ens.draw(calls * K)
        draw_s = time.perf_counter() - t0
        timed, out_d = ens.device_streams(K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        e0.record(); timed(); e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        ms_total = e0.elapsed_time(e1)
        assert torch.isfinite(out_d[:, -1, :4]).all()
        # e2e: the public call with host buffers (host RNG pre-draw + H2D + launch + D2H)
        t0 = time.perf_counter()
        ens.learn_many(calls * K)
        e2e_s = time.perf_counter() - t0
    ms_total = parallel.max_over_ranks(ms_total, dev)
    e2e_s = parallel.max_over_ranks(e2e_s, dev)
    if rank == 0:
        B = wl["batch"]
        value = world * R * K / (ms_total * 1e-3)
        fp32 = C.c_double(0)
        _lib.check(_lib.load().hirl_measure_fp32_peak(local, C.byref(fp32)))
        achieved = wl["flop_per_row"] * B * R * K / (ms_total * 1e-3) / 1e12
        hbm = 11.1e6 * R * K / (ms_total * 1e-3) / 1e9
        line = {"metric": "agent updates/sec", "value": value, "unit": "replica-updates/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(config_of(args, wl), replicas_per_gpu=R, replicas_total=R * world,
                                                      parallelism=f"{R} replicas/GPU x {world} GPUs, teams of CTAs, no collective"),
                "clocks": sampler.summary(),
                "e2e": {"value": world * R * K / e2e_s, "unit": "replica-updates/s", "h2d_bytes_per_step": R * (3 * B * 4 + 64),
                        "d2h_bytes_per_step": R * 32, "api": "AgentEnsemble.learn_many (host RNG streams of every replica + H2D + launch + D2H)",
                        "host_draw_seconds": draw_s},
                "gpu_launches": 1,
                "roofline": {"kernel": "hirl_ensemble_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32.value, "unit": "TFLOP/s",
                             "frac": achieved / fp32.value, "traffic": None, "hbm_stream_gbs_if_weights_streamed": hbm,
                             "note": "weights + Adam state of 128 replicas = 1.1 GB > L2: 11.1 MB algorithmic HBM bytes per replica-update"}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_ours(args):
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    if "replicas" in WORKLOADS[args.workload]:
        return run_ensemble(args)
    from hirl4ucav_b200 import _lib
    lib = _lib.load()
    wl = WORKLOADS[args.workload]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    peaks, peak_src = measured_peaks()
    ag = build_agent(wl, local, seed=rank, max_steps=max(K, W, 64))
    sampler = ClockSampler(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with sampler:
        # ---- value: device-resident streams, one persistent launch for the K timed steps
        warm, _ = device_resident_run(ag, wl, W)
        warm(); torch.cuda.synchronize()
        timed, out_d = device_resident_run(ag, wl, K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(); timed(); e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        last = out_d[-1].cpu().numpy()
        assert np.isfinite(last[:4]).all(), f"non-finite losses {last}"
        # ---- e2e: the public per-call API, host buffers, copies inside the timed region
        for _ in range(W):
            ag.learn(wl["bcw"], wl["expert_num"], 0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            r = ag.learn(wl["bcw"], wl["expert_num"], 0)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        # ---- e2e, fused form of the same public API: learn_many in chunks of 60 calls
        chunk = [(wl["bcw"], wl["expert_num"], 0)] * 60
        nchunks = max(1, K // 60)
        ag.learn_many(chunk)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(nchunks):
            ag.learn_many(chunk)
        torch.cuda.synchronize()
        fused_s = time.perf_counter() - t0
        if rank == 0 and world == 1:
            roof_env = env_roofline(dev, peaks)
    t = torch.tensor([ms_total, e2e_s, fused_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, fused_s = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    B = wl["batch"]
    value = world * K / (ms_total * 1e-3)
    flops_per_step = wl["flop_per_row"] * B
    fp32 = C.c_double(0)
    _lib.check(lib.hirl_measure_fp32_peak(local, C.byref(fp32)))
    achieved = flops_per_step * K / (ms_total * 1e-3) / 1e12
    line = {
        "metric": "agent updates/sec", "value": value, "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_of(args, wl), "clocks": sampler.summary(),
        "e2e": {"value": world * K / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": 3 * B * 4 + 64, "d2h_bytes_per_step": 32,
                "api": "Agent.learn() per call (hirl_learn: pinned host index/noise streams -> H2D, one launch, D2H scalars)"},
        "e2e_fused": {"value": world * nchunks * 60 / fused_s, "unit": "updates/s", "api": "Agent.learn_many(60 calls) per launch, host buffers"},
        "gpu_launches": 1,
        "gpu_launches_note": "value: ONE persistent hirl_update_kernel launch covers the K timed steps; e2e: K launches",
        "roofline": {"kernel": "hirl_update_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32.value, "unit": "TFLOP/s",
                     "frac": achieved / fp32.value if fp32.value else None,
                     "traffic": NCU_UPDATE_DRAM_BYTES_PER_STEP[args.workload] * K if args.workload in NCU_UPDATE_DRAM_BYTES_PER_STEP else None,
                     "traffic_source": NCU_TRAFFIC_SOURCE,
                     "peak_source": "in-run register-resident FFMA loop (hirl_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                     "flop_per_step": flops_per_step,
                     "frac_of_bf16_tensor_burst": achieved / peaks.get("bf16_tflops", 1660.5), "tensor_peak_source": peak_src},
    }
    if world == 1:
        line["roofline_env"] = roof_env
        torch.set_num_threads(os.cpu_count() or 1)
        calls = 300 if B <= 256 else 36
        val, dt = cpu_baseline(wl, calls)
        line["cpu_baseline"] = {"value": val, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{calls} learn() calls of the same workload after 6 warm-up calls ({dt:.1f} s); oracle/hirl_oracle.py, "
                                          "bit-identical to the unmodified reference in the build container"}
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
