#!/usr/bin/env python
"""bench.py -- BPR triplets/s on B200 (BASELINE.json configs[1]) with roofline + CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (gather -> score -> loss -> sparse gradient -> dedup -> Adagrad)
over one batch of 65 536 synthetic triplets against 1M x 128 user / item tables.

 value : whole-job triplets/s, id batches already resident in HBM, C-ABI called directly.
 e2e   : the same metric through the public API a user calls (openrec.tf2 BPR model + GradientTape +
         optimizer.apply_gradients), ids copied from pinned HOST memory and the loss read back to the
         host every step, all inside the timed region.
 roofline : dominant kernel (k_pair_step) timed live with CUDA events on its launch stream
         (orx_profile_*), algorithmic bytes / duration vs the measured HBM peak.
 cpu_baseline / --impl reference : the CPU restatement of the reference step (oracle/c C+OpenMP port;
         TensorFlow is not installable here) on the box's host cores.
Prints exactly ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

U = I = 1_000_000
D = 128
B = 65_536
N_BATCHES = 16
LR = 0.05
METRIC = "bpr_triplets_per_sec"
UNIT = "triplets/s"
# bytes/triplet = 12 + 4*(3D+2)*(2+2S); Adagrad S=1, D=128 -> 6188
ALG_BYTES_PER_TRIPLET = 12 + 4 * (3 * D + 2) * (2 + 2 * 1)


def workload_name(n_gpus):
    items = I if n_gpus == 1 else 12_500_000 * n_gpus          # BASELINE configs[1] / configs[4] (100M items on 8 GPUs)
    s = (f"BPR {U} users x {items} items, dim {D}, batch {B} per GPU, Adagrad lr {LR} (acc init 0.1), "
         f"ids uniform i.i.d. int32, {N_BATCHES} rotating id batches")
    if n_gpus > 1:
        mode = os.environ.get("ORX_SHARDED", "mailbox")
        how = {"mailbox": "liborx kernels store ids / rows / gradient rows into the peers' IPC-mapped mailboxes over NVLink "
                          "(no collective in the step)",
               "peer": "one-sided peer loads / stores on the mapped shards"}.get(
                   mode, "NCCL all-to-all exchange of counts / ids / rows / gradient rows")
        s += f", user and item tables row-sharded over {n_gpus} GPUs (row r on rank r % N, 12.5M item rows per GPU); {how}"
    return s


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ---------------------------------------------------------------------------------------
# clocks: a separate process polls NVML so the timed Python loop keeps the GIL
# ---------------------------------------------------------------------------------------
_CLOCK_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
out = open(sys.argv[2], "w")
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
while True:
    try:
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    out.write("%f %d %d %d\n" % (time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, r))
    out.flush()
    time.sleep(0.004)
"""
_REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
            0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
            0x100: "display_clock_setting"}


class ClockSampler:
    def __init__(self, index):
        self.path = tempfile.mktemp(prefix="orx_clocks_")
        try:
            self.p = subprocess.Popen([sys.executable, "-c", _CLOCK_SRC, str(index), self.path],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        self.windows = []

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML sampler unavailable"}
        time.sleep(0.02)
        self.p.terminate()
        rows = []
        try:
            for line in open(self.path):
                f = line.split()
                if len(f) == 4:
                    rows.append((float(f[0]), int(f[1]), int(f[2]), int(f[3])))
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if any(a <= r[0] <= b for a, b in self.windows)]
        note = "samples inside the timed regions"
        if not inside:
            inside, note = rows, "timed regions shorter than the 4 ms poll; all samples of this run"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        bits = 0
        for r in inside:
            bits |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(inside[0][2]),
                "reasons": sorted(v for k, v in _REASONS.items() if bits & k), "samples": len(inside), "note": note}


# ---------------------------------------------------------------------------------------
# CPU arm: the restated reference step (C/OpenMP port of the oracle) on the host cores
# ---------------------------------------------------------------------------------------
def cpu_arm(steps, warmup, budget_s, batch=B):
    """Times `steps` steps of `batch` triplets (bounded by budget_s).  Returns (triplets/s, info)."""
    from oracle import c_port
    threads = c_port.num_threads()
    rng = np.random.default_rng(0)

    def tab(rows, cols):
        return rng.random((rows, cols), dtype=np.float32) * np.float32(0.1) - np.float32(0.05)

    user, item, bias = tab(U, D), tab(I, D), tab(I, 1)
    acc = [np.full_like(a, 0.1) for a in (user, item, bias)]
    ids = [tuple(rng.integers(0, n, batch, dtype=np.int32) for n in (U, I, I)) for _ in range(4)]

    def step(i):
        u, p, n = ids[i % len(ids)]
        return c_port.pairwise_step("bpr", user, acc[0], item, acc[1], bias, acc[2], u, p, n, 1, LR, nthreads=threads)

    # "all the host threads it can use": the port is memory-bound and SMT siblings slow it down (r1a: 0.17 M/s on 128
    # threads vs 0.80 M/s on 64), so time one step at all / half / quarter of the visible CPUs and keep the fastest
    try:
        visible = len(os.sched_getaffinity(0))
    except AttributeError:
        visible = os.cpu_count() or threads
    best = None
    for cand in sorted({max(1, visible), max(1, visible // 2), max(1, visible // 4), max(1, threads)}, reverse=True):
        threads = cand
        step(0)                                   # first touch / warm
        t = time.perf_counter()
        step(1)
        t = time.perf_counter() - t
        if best is None or t < best[0]:
            best = (t, cand)
    threads = best[1]

    t0 = time.perf_counter()
    for i in range(max(1, warmup)):
        step(i)
        if time.perf_counter() - t0 > budget_s * 0.25:
            break
    done, t0 = 0, time.perf_counter()
    while done < steps:
        step(done)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    info = {"cores": threads, "kind": "port",
            "sample": f"{done} steps x {batch} triplets of the same workload (same table sizes, Adagrad), "
                      f"C/OpenMP port of the oracle, {threads} threads (fastest of all / half / quarter of the "
                      f"{visible} visible CPUs, one timed step each), {dt:.1f} s"}
    return done * batch / dt, dt / done * 1e3, done, info


def run_reference(args, rank, world):
    if rank != 0:
        return
    budget = float(os.environ.get("ORX_CPU_BUDGET_S", "60"))
    v, ms, done, info = cpu_arm(args.steps, args.warmup, budget)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(1), "note": "CPU restatement of openrec.tf2 (TensorFlow not "
                       "installable): oracle/c/orx_oracle.c, all host threads; always the single-GPU workload (1M x 1M): "
                       "the reference has no multi-device path and the 100M-item tables need ~100 GB of host memory"},
            "cpu_baseline": {"value": v, "unit": UNIT, **info},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from openrec_b200 import native as N

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its "NCCL version ..." banner to stdout whenever NCCL_DEBUG is set; the contract is ONE JSON line
        os.environ.pop("NCCL_DEBUG", None)
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    eng = N.engine(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        from openrec_b200 import sharded
        result = sharded.bench(args, rank, world, eng, barrier)
    else:
        result = bench_single(args, eng, dev, barrier)
    if world > 1:
        t = torch.tensor([result["seconds"], result["e2e_seconds"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max over ranks
        result["seconds"], result["e2e_seconds"] = t[0].item(), t[1].item()
    if rank == 0:
        K = args.steps
        units = K * B * world
        line = {"metric": METRIC, "value": units / result["seconds"], "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": args.warmup, "ms_per_step": result["seconds"] / K * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(world), "optimizer": "Adagrad (Keras sparse semantics)",
                           "l2_flush": ("none needed: tables+accumulators 2.06 GB per GPU and a 406 MB/step random "
                                        "working set >> 126 MB L2") if world == 1 else
                                       ("none needed: 13.3 GB of table + accumulator per GPU, 0.3 GB of mailbox traffic and "
                                        "0.4 GB of random row updates per step >> 126 MB L2"),
                           "parallelism": "single GPU" if world == 1 else f"row-sharded tables x{world}"},
                "clocks": result["clocks"],
                "e2e": {"value": units / result["e2e_seconds"], "unit": UNIT,
                        "h2d_bytes_per_step": 3 * 4 * B * world, "d2h_bytes_per_step": 16 * world,
                        "api": result["e2e_api"]},
                "gpu_launches": result["launches"], "roofline": result["roofline"]}
        if result.get("cpu_baseline"):
            line["cpu_baseline"] = result["cpu_baseline"]
        if result.get("extra"):
            line["extra"] = result["extra"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_single(args, eng, dev, barrier):
    import torch
    from openrec_b200 import native as N
    K, W = args.steps, max(3, args.warmup)
    tu, ti = torch.empty(U, D, device=dev), torch.empty(I, D, device=dev)
    tb = torch.empty(I, 1, device=dev)
    for k, t in enumerate((tu, ti, tb)):
        eng.fill_uniform(t, -0.05, 0.05, 1000 + k)
    acc = [torch.full_like(t, 0.1) for t in (tu, ti, tb)]
    tabs = (N.table(tu, acc[0]), N.table(ti, acc[1]), N.table(tb, acc[2]))
    g = torch.Generator(device="cpu").manual_seed(1)
    host_ids = [tuple(torch.randint(0, n, (B,), generator=g, dtype=torch.int32).pin_memory() for n in (U, I, I))
                for _ in range(N_BATCHES)]
    dev_ids = [tuple(x.to(dev) for x in b) for b in host_ids]
    out4 = torch.zeros(4, device=dev)
    opt = N.opt(N.ORX_OPT_ADAGRAD, LR)
    clocks = ClockSampler(dev.index or 0)

    def step(i):
        u, p, n = dev_ids[i % N_BATCHES]
        eng.pairwise_step(N.ORX_PAIR_BPR, *tabs, u, p, n, opt, out4)

    for i in range(W):
        step(i)
    # ---- value: ids resident in HBM, direct C-ABI; dominant kernel timed live by orx_profile_*
    barrier()
    eng.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    barrier()
    t1 = time.time()
    clocks.window(t0, t1)
    seconds = e0.elapsed_time(e1) * 1e-3
    phase_ms, n_prof = eng.profile_read()
    eng.profile_enable(False)
    loss_check = out4.cpu().numpy().tolist()

    # ---- secondary workloads on the same tables (BASELINE configs[2] UCML; SGD variant), short loops
    def timed(kind, o, n=min(K, 300)):
        for i in range(5):
            u, p, q = dev_ids[i % N_BATCHES]
            eng.pairwise_step(kind, *tabs, u, p, q, o, out4)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            u, p, q = dev_ids[i % N_BATCHES]
            eng.pairwise_step(kind, *tabs, u, p, q, o, out4)
        e1.record()
        torch.cuda.synchronize()
        return n * B / (e0.elapsed_time(e1) * 1e-3)
    secondary = {"ucml_adagrad_triplets_per_sec": timed(N.ORX_PAIR_UCML, opt),
                 "bpr_sgd_triplets_per_sec": timed(N.ORX_PAIR_BPR, N.opt(N.ORX_OPT_SGD, LR))}

    # ---- e2e: public API (openrec.tf2 BPR + GradientTape + Adagrad), host ids in, loss out, every step
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    import tensorflow as tf
    from openrec.tf2.recommenders import BPR
    del tu, ti, tb, acc, tabs
    torch.cuda.empty_cache()
    model = BPR(dim_user_embed=D, dim_item_embed=D, total_users=U, total_items=I)
    optimizer = tf.keras.optimizers.Adagrad(learning_rate=LR)

    def train_step(user_id, p_item_id, n_item_id):
        with tf.GradientTape() as tape:
            loss_value = model(user_id, p_item_id, n_item_id)
        gradients = tape.gradient(loss_value, model.trainable_variables)
        optimizer.apply_gradients(zip(gradients, model.trainable_variables))
        return loss_value

    last = 0.0
    for i in range(W):
        last = float(train_step(*host_ids[i % N_BATCHES])[0])
    barrier()
    t0 = time.time()
    e0.record()
    prev = None
    for i in range(K):
        loss_value = train_step(*host_ids[i % N_BATCHES])   # pinned host ids -> device inside the call
        if prev is not None:
            last = float(prev[0])                            # every step's loss is read on the host, one step
        prev = loss_value                                    # behind so the copy overlaps the next launch
    last = float(prev[0])
    e1.record()
    barrier()
    t1 = time.time()
    clocks.window(t0, t1)
    e2e_seconds = e0.elapsed_time(e1) * 1e-3

    peak, peak_src = measured_peak()
    step_ms = phase_ms[1] / max(n_prof, 1)
    achieved = ALG_BYTES_PER_TRIPLET * B / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k_pair_step_traffic.json")) as f:
            traffic = json.load(f)["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_pair_step<BPR,ADAGRAD,D=128,CH=8,4 CTAs/SM>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_TRIPLET * B, "kernel_ms": step_ms,
                "phase_ms_per_step": {"index_build": phase_ms[0] / max(n_prof, 1), "pair_step": step_ms,
                                      "tail": phase_ms[2] / max(n_prof, 1)},
                "kernel_share_of_step": phase_ms[1] / max(sum(phase_ms), 1e-9),
                "timing": f"CUDA events around the three launches of every 8th step of the timed region ({n_prof} steps), "
                          "on the stream the kernels run on (orx_profile_*)"}
    result = {"seconds": seconds, "e2e_seconds": e2e_seconds, "clocks": clocks.stop(), "launches": 3 * K,
              "roofline": roofline,
              "e2e_api": "openrec.tf2.recommenders.BPR + tf.GradientTape + tf.keras.optimizers.Adagrad (shim); "
                         "pinned host ids in, loss read to host each step",
              "extra": {"last_loss_value_path": loss_check[:2], "last_loss_e2e": last, **secondary}}
    if not args.no_cpu:
        v, ms, done, info = cpu_arm(50, 2, float(os.environ.get("ORX_CPU_BUDGET_S", "20")))
        result["cpu_baseline"] = {"value": v, "unit": UNIT, **info}
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        if rank == 0:
            sys.stderr.write(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run\n")
        sys.exit(2)
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
