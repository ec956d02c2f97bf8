#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline configuration.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (N>1 under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the path on the box's host cores

Workload: Unitree G1-12dof, one 1 000 000-sample synthetic log (BASELINE.json configs[3]); with N GPUs the SAME log is
sharded into contiguous N/G-sample shards (strong scaling), each rank builds its Gram statistics, one NCCL all-reduce of
c^2+c+2 = 23 872 doubles merges them.  A "step" is one pass of regressor -> projector -> Gram over the whole log.
  value  samples/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e    the same samples/s through identify(): pinned HOST arrays -> H2D -> fused kernel -> all-reduce -> LMI solve ->
         D2H of the identified parameters, every step
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "regressor+Gram samples/s (fp64)"
UNIT = "samples/s"
ROBOT = "g1_12dof"
N_SAMPLES = int(os.environ.get("SYSID_BENCH_SAMPLES", 1_000_000))
FLOP_PER_SAMPLE = 18 * 154 * 155 + 2 * 18 * 154          # 435 204: lower-triangle Gram + A^T b
BYTES_PER_SAMPLE = 8 * (19 + 18 + 18 + 12 + 2)            # 552 B of fp64 input per G1-12 sample
FP64_PEAK_FALLBACK_TFLOPS = 35.77                         # profiles/fp64_peak_r01.json: cuBLAS DGEMM 8192^3 on this pool
CPU_SAMPLE = int(os.environ.get("SYSID_BENCH_CPU_SAMPLES", 150_000))      # cpu_baseline leg of OUR arm (bounded, ~10 s); the reference
REF_FULL = os.environ.get("SYSID_BENCH_REF_FULL", "1") == "1"                # arm times the WHOLE log per step (same config as ours)
WORKLOAD = f"{ROBOT} {N_SAMPLES}-sample log (BASELINE configs[3]), regressor+projector+Gram with friction columns, c=154"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    return ap.parse_args()


KERNEL_SOURCES = ("gram_struct.cuh", "gram_tiles_struct.inc", "gram_kernels.cuh", "phases.cuh", "tmem_park.cuh", "gram_tiles.inc")


def kernel_source_stamp():
    """sha256 of the fused kernel's sources as they are in this tree (.git does not travel to the GPU box)."""
    import hashlib
    h = {}
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "system_identification_b200", "csrc", f), "rb") as fh:
            h[f] = hashlib.sha256(fh.read()).hexdigest()[:16]
    return h


def ncu_traffic(n_loc):
    """DRAM bytes of one fused-kernel launch from the committed ncu capture (profiles/gram_fused_traffic.json, written by
    tools/update_traffic.py from an `ncu --set full` report).  Returned only when the capture was taken from the kernel
    sources that are being benchmarked (content stamp) and, at N = 1, from a launch of the same size; otherwise None plus the
    reason -- a stale constant is not a measurement."""
    try:
        with open(os.path.join(ROOT, "profiles", "gram_fused_traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, "no capture under profiles/"
    if t.get("source_stamp") != kernel_source_stamp():
        return None, "the committed capture predates the current kernel sources (stamp mismatch): re-capture with tools/update_traffic.py"
    per_launch = t["dram_bytes_read"] + t["dram_bytes_write"]
    if n_loc != t["samples_per_launch"]:
        return per_launch * n_loc / t["samples_per_launch"], f"scaled linearly from a {t['samples_per_launch']}-sample capture ({t.get('git_sha', '?')})"
    return per_launch, f"ncu --set full capture of the same launch ({t.get('git_sha', '?')})"


def load_flat():
    from system_identification_b200.model import FlatModel
    return FlatModel.load(os.path.join(ROOT, "system_identification_b200", "robots", ROBOT + ".json"))


def host_log(flat, N):
    """The five (channels x N) fp64 arrays of the workload; torques are smooth + noise (any finite values give the
    same arithmetic; bench_identifiable_tau() below replaces them by ground-truth torques on the device)."""
    from system_identification_b200 import synth
    q, dq, ddq, cnt = synth.make_trajectory(flat, N, synth.SEEDS["g1_1m"])
    tau = synth.synth_tau(flat, N, 11, scale=10.0)
    return q, dq, ddq, tau, cnt


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_port_throughput(flat, data, n_samples, threads=0, repeats=1):
    """Oracle C restatement (kind 'port') of regressor -> projector -> Gram on the host cores."""
    from oracle import urdf_tree as ut
    from oracle.cbuild import COracle
    co = COracle(ut.tree_from_flat(flat), flat.ee_names)
    if threads <= 0:
        # every core this process may run on, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
        # otherwise turn the all-cores CPU arm into a single-thread run whenever N > 1
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sub = tuple(a[:, :n_samples] for a in data)
    best = None
    used = 1
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, used = co.gram(*sub, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_samples / best, used, best


def oracle_sdp_seconds(flat, stats, c=154):
    """B3: the oracle's stage-3 solve of the given statistics (log-barrier Newton in numpy, oracle/sdp.py::solve_barrier:
    the oracle's STATED CPU reference solve and the faster of its two methods), one core."""
    import numpy as np
    from oracle import sdp as osdp
    G = stats[:c * c].reshape(c, c); r = stats[c * c:c * c + c]; s_, n_ = float(stats[c * c + c]), float(stats[c * c + c + 1])
    t0 = time.perf_counter()
    prob = osdp.build_problem(G, r, s_, n_, flat.nbodies, flat.phi_prior, flat.robot_mass, flat.ellipsoids, flat.joints_dof)
    x, info = osdp.solve_barrier(prob)
    return time.perf_counter() - t0, x


def reference_shaped_seconds_per_sample(flat, data, n=1500):
    """B1: the path the way the reference executes it -- a Python loop over samples, numpy pinv, per-sample blocks stacked,
    one core (oracle/dynamics.py::stacked_system restates demo/solo_identification.py:36-55,79-84) -- on a bounded sample."""
    from oracle import dynamics as dy
    from oracle import urdf_tree as ut
    t = ut.tree_from_flat(flat)
    sub = tuple(a[:, :n] for a in data)
    t0 = time.perf_counter()
    A, b = dy.stacked_system(t, *sub, flat.ee_names)
    G = A.T @ A
    return (time.perf_counter() - t0) / n, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    flat = load_flat()
    n = N_SAMPLES if REF_FULL else min(CPU_SAMPLE, N_SAMPLES)
    data = host_log(flat, n)
    for _ in range(max(args.warmup, 0)):
        cpu_port_throughput(flat, data, min(n, 20000))
    from oracle import urdf_tree as ut
    from oracle.cbuild import COracle
    co = COracle(ut.tree_from_flat(flat), flat.ee_names)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    times, solve_times = [], []
    used = 1
    for _ in range(max(args.steps, 1)):
        t0 = time.perf_counter()
        stats, used = co.gram(*data, nthreads=threads)
        times.append(time.perf_counter() - t0)
        st, _ = oracle_sdp_seconds(flat, stats)              # B3: the stage-3 solve belongs to the end-to-end figure
        solve_times.append(st)
    T = sum(times)
    value = n * len(times) / T
    e2e_value = n * len(times) / (T + sum(solve_times))
    b1_s, b1_n = reference_shaped_seconds_per_sample(flat, data)
    sample = (f"the whole {n}-sample G1-12dof log per step" if n == N_SAMPLES else f"first {n} samples of the {N_SAMPLES}-sample G1-12dof log per step") + \
             " (regressor+projector+Gram: oracle/sysid_oracle.c, OpenMP; stage 3: oracle/sdp.py solve_barrier, numpy)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "samples_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample,
                         "b1_reference_shaped": {"value": 1.0 / b1_s, "unit": UNIT, "cores": 1, "sample": f"first {b1_n} samples; per-sample Python loop, numpy pinv, stacked A (oracle/dynamics.py::stacked_system)"},
                         "b3_sdp_solve_seconds": sum(solve_times) / len(solve_times),
                         "note": "pinocchio/cvxpy/MOSEK are not installable in this image; oracle/sysid_oracle.c restates the reference's per-sample arithmetic"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "identify_seconds": (T + sum(solve_times)) / len(times),
                "note": "statistics (C/OpenMP, all cores) + LMI solve (numpy oracle, one core) per step: the same work as the GPU arm's identify()"},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # under load = the upper half of the samples (idle samples before/after the region sit at the low end)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from system_identification_b200 import distributed as D
    from system_identification_b200 import ops
    from system_identification_b200.identify import identify
    from system_identification_b200.sys_identification import SystemIdentification

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    flat = load_flat()
    si = SystemIdentification.from_flat_model(flat)
    dm = si.device_model

    lo, hi = D.shard_bounds(N_SAMPLES, rank, world)
    n_loc = hi - lo
    q, dq, ddq, tau, cnt = host_log(flat, N_SAMPLES)
    shard = [np.ascontiguousarray(a[:, lo:hi]) for a in (q, dq, ddq, tau, cnt)]
    dev = [ops.to_device(a) for a in shard]
    # ground-truth torques (identifiable problem for the e2e solve), generated on the device in chunks, off the clock
    tau_dev = identifiable_tau(flat, dm, dev, seed=17 + rank)
    dev[3] = tau_dev
    pinned = [torch.from_numpy(a).pin_memory() for a in shard]
    pinned[3] = tau_dev.cpu().pin_memory()
    torch.cuda.synchronize()

    c = dm.ncols(True)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        stats = dm.gram_accumulate(*dev)
        D.allreduce_stats(stats)
        return stats

    # fp64 roofline denominator, measured live (cuBLAS DGEMM through torch), rank 0's GPU
    peak_tflops, peak_how = measure_fp64_peak(torch)

    if args.warmup < 3:
        print(f"bench.py: --warmup {args.warmup} is below the contract's minimum of 3; running exactly {args.warmup} as asked", file=sys.stderr)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(float(k))                      # evict the log from L2 between timed iterations (not timed)
        barrier()
        ev[k][0].record()
        kev[k][0].record()
        stats = dm.gram_accumulate(*dev)           # fused kernel + 10-us reduction kernel
        kev[k][1].record()
        D.allreduce_stats(stats)
        ev[k][1].record()
    barrier()
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    k_ms = sum(a.elapsed_time(b) for a, b in kev)
    tt = torch.tensor([t_ms, k_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, k_ms = float(tt[0]), float(tt[1])

    # ---- end to end through identify(): pinned host arrays in, parameters out, every step ---------------------------
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_once():
        # pinned host arrays straight into the public call: chunked upload overlapped with the kernel (C-ABI
        # sysid_gram_accumulate_host), all-reduce, LMI solve, parameters back on the host
        return identify(si, *pinned, sharded=True, return_info=True)
    e2e_once()
    barrier()
    t0 = time.perf_counter()
    e2e_ev0, e2e_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ev0.record()
    last = None
    for _ in range(e2e_steps):
        last = e2e_once()
    e2e_ev1.record()
    barrier()
    e2e_ms = e2e_ev0.elapsed_time(e2e_ev1)
    e2e_wall = (time.perf_counter() - t0) * 1e3
    te = torch.tensor([max(e2e_ms, e2e_wall)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])
    clocks = sampler.stop() if rank == 0 else None
    stats_e2e = dm.gram_accumulate(*dev) if world == 1 else None     # the statistics the e2e solve saw (B3 solves the same problem on the CPU)

    if rank == 0:
        value = N_SAMPLES * args.steps / (t_ms * 1e-3)
        kernel_s = (k_ms / args.steps) * 1e-3
        achieved = FLOP_PER_SAMPLE * n_loc / kernel_s * 1e-12
        info = last[3] or {}
        cpu = None
        if world == 1:
            # cpu_baseline leg (rank 0, N = 1 only): B2 = the oracle's C/OpenMP twin on a bounded sample of the same log, B1 = the
            # reference-shaped numpy loop on one core, B3 = the oracle's stage-3 solve of THIS log's statistics
            ncpu = min(CPU_SAMPLE, N_SAMPLES)
            cpu_thr, cpu_cores, cpu_dt = cpu_port_throughput(flat, (q, dq, ddq, tau, cnt), ncpu)
            b1_s, b1_n = reference_shaped_seconds_per_sample(flat, (q, dq, ddq, tau, cnt))
            b3_s, x_cpu = oracle_sdp_seconds(flat, stats_e2e.cpu().numpy())
            x_gpu = __import__("numpy").concatenate([last[0], last[1], last[2]])
            cpu = {"value": cpu_thr, "unit": UNIT, "cores": cpu_cores, "kind": "port",
                   "sample": f"first {ncpu} samples of the same log, oracle/sysid_oracle.c with OpenMP ({cpu_dt:.1f} s)",
                   "b1_reference_shaped": {"value": 1.0 / b1_s, "unit": UNIT, "cores": 1,
                                           "sample": f"first {b1_n} samples; per-sample Python loop, numpy pinv, stacked A (oracle/dynamics.py::stacked_system)"},
                   "b3_sdp_solve_seconds": b3_s,
                   "b3_phi_rel_diff_vs_gpu": float(__import__("numpy").linalg.norm(x_gpu - x_cpu) / __import__("numpy").linalg.norm(x_cpu))}
        traffic, traffic_how = ncu_traffic(n_loc)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "samples_per_rank": n_loc, "sharding": f"contiguous time shards over {world} rank(s), one NCCL all-reduce of {c * c + c + 2} fp64",
                       "l2": "256 MiB buffer rewritten between timed iterations (inputs per rank: %.0f MB)" % (BYTES_PER_SAMPLE * n_loc / 1e6)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": traffic, "traffic_source": traffic_how,
                         "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/gram_fused_traffic.json)",
                         "algorithmic_bytes": BYTES_PER_SAMPLE * n_loc, "peak_source": peak_how,
                         "note": "fp64 DMMA contraction; achieved = 435204 algorithmic FLOP/sample x samples per launch / CUDA-event time of the fused kernel (+ its 10-us reduction kernel)",
                         "hbm_stream_gbs": BYTES_PER_SAMPLE * n_loc / kernel_s * 1e-9},
            "e2e": {"value": N_SAMPLES * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": BYTES_PER_SAMPLE * N_SAMPLES,
                    "d2h_bytes_per_step": 8 * c, "identify_seconds": e2e_ms * 1e-3 / e2e_steps, "steps": e2e_steps,
                    "solver": {k: info.get(k) for k in ("status", "iterations", "refactorizations", "primal_residual", "dual_residual")}},
            "gpu_launches": 2 * args.steps,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_fp64_peak(torch):
    try:
        n = 8192
        a = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        b = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        best = None
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) * 1e-12, "measured in this run: cuBLAS DGEMM 8192^3, best of 4 (MEASURED_PEAKS.json has no fp64 figure)"
    except Exception as e:  # pragma: no cover
        return FP64_PEAK_FALLBACK_TFLOPS, f"fallback profiles/fp64_peak_r01.json ({e})"


def identifiable_tau(flat, dm, dev, seed):
    """Ground-truth torques for the e2e solve (data generation, outside every timed region)."""
    from system_identification_b200.synth import identifiable_tau_device
    return identifiable_tau_device(flat, dm, dev, seed)


# Rank 0 prints exactly ONE line on stdout.  Libraries do not know that (NCCL writes its version banner to stdout whenever
# NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is
# written to the saved descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


if __name__ == "__main__":
    capture_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
