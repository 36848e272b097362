#!/usr/bin/env python
"""bench.py - headline benchmark of the Gemini prover hot path on B200.

Metric (BASELINE.json): G1 MSM throughput in scalar-mults/s.  One "step" = one kzg::commit-equivalent
MSM (VariableBaseMSM::msm_unchecked, /root/reference/src/kzg/time.rs:81-83) over n = 2^logn synthetic
BLS12-381 G1 bases resident on the device and n uniformly random Fr scalars.

  own arm       python bench.py --gpus N --steps K --warmup W [--logn 20]
                (N > 1: launched under torchrun, one rank per GPU; each rank owns a contiguous range
                 of n points = weak scaling; partial G1 sums are exchanged with one NCCL all-gather
                 of 144-byte Jacobian points and added on every rank)
  reference arm python bench.py --impl reference ...   the CPU restatement of arkworks' Pippenger
                (oracle/gemini_oracle.c, one thread per window like ark-ec's rayon tasks) on the host
                cores; the reference itself is Rust and cannot be built in this image (no cargo).

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "g1_msm_throughput"
UNIT = "scalar-mults/s"
ALGO_BYTES_PER_TERM = 128  # 32 B Fr scalar + 96 B packed affine base, each read once
# dram__bytes_read.sum + dram__bytes_write.sum of the kernels of the bucket-accumulation phase of ONE MSM (precomputed
# table, uniform scalars), summed from the ncu launch lists profiles/r01_launches_n20_affine2_dram.csv and
# profiles/r01_launches_n24_affine4_dram.csv (same bench.py command under ncu); the phase's serialised ncu time
# (4.68 ms / 55.5 ms) agrees with the live CUDA-event time below
NCU_PHASE_TRAFFIC = {20: 8.77e9, 24: 148.9e9}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU baseline (oracle/gemini_oracle.c) - the ONLY place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------
def load_oracle():
    so = os.path.join(ROOT, "oracle", "libgemini_oracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.go_msm_g1.restype = C.c_int
    lib.go_msm_g1.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    lib.go_generate_bases.argtypes = [C.c_size_t, C.c_uint64, C.c_void_p]
    return lib


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def splitmix_scalars(n: int, seed: int):
    """numpy twin of k_fr_random (gemini_b200/csrc/fr.cu)."""
    import numpy as np

    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    with np.errstate(over="ignore"):
        x = np.arange(4 * n, dtype=np.uint64) + np.uint64(seed)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        w = (x ^ (x >> np.uint64(31))).reshape(n, 4)
    w[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)
    rl = [np.uint64((R >> (64 * j)) & 0xFFFFFFFFFFFFFFFF) for j in range(4)]
    ge = np.zeros(n, dtype=bool)
    eq = np.ones(n, dtype=bool)
    for j in (3, 2, 1, 0):
        ge |= eq & (w[:, j] > rl[j])
        eq &= w[:, j] == rl[j]
    ge |= eq
    for i in np.nonzero(ge)[0]:
        v = sum(int(w[i, j]) << (64 * j) for j in range(4)) - R
        w[i] = [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
    return w


def cpu_msm_timed(lib, bases, scalars, n, threads):
    import numpy as np

    out = np.zeros(12, dtype=np.uint64)
    t0 = time.perf_counter()
    c = lib.go_msm_g1(bases.ctypes.data, scalars.ctypes.data, n, 0, threads, out.ctypes.data)
    return time.perf_counter() - t0, out, c


def pick_sample(lib, bases, scalars, n, threads, target_s):
    """largest power-of-two prefix of the workload whose MSM takes about target_s on this host"""
    probe = min(n, 1 << 13)
    dt, _, _ = cpu_msm_timed(lib, bases, scalars, probe, threads)
    per_term = dt / probe
    m = probe
    while m * 2 <= n and per_term * m * 2 * 0.8 <= target_s:  # larger windows make big instances a bit cheaper per term
        m *= 2
    return m


def run_reference(args):
    """--impl reference: the CPU path on the host cores; rank 0 only."""
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib = load_oracle()
    n = 1 << args.logn
    threads = host_threads()
    cap = min(n, 1 << 20)
    bases = np.zeros((cap, 12), dtype=np.uint64)
    lib.go_generate_bases(cap, 1, bases.ctypes.data)
    scalars = splitmix_scalars(cap, 1000)
    m = pick_sample(lib, bases, scalars, cap, threads, target_s=4.0)
    for _ in range(args.warmup):
        cpu_msm_timed(lib, bases, scalars, min(m, 1 << 12), threads)
    times = []
    c_used = 0
    for _ in range(args.steps):
        dt, _, c_used = cpu_msm_timed(lib, bases, scalars, m, threads)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = m / (ms / 1e3)
    sample = (f"first 2^{m.bit_length() - 1} terms of the 2^{args.logn} workload per step; signed-digit Pippenger c={c_used}, "
              f"one thread per window (<= {(255 + c_used - 1) // c_used} usable threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"kzg::commit n=2^{args.logn} BLS12-381 G1 (CPU restatement of ark-ec msm_unchecked; "
                               "the Rust reference cannot be built: no cargo in the image)",
                   "bases": "P_i=[i+1]G", "scalars": "uniform Fr (splitmix64)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            sm, mx, reasons = [], [], set()
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                sm.append(float(p[1])); mx.append(float(p[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def run_own(args):
    import numpy as np
    import torch

    import gemini_b200 as gm
    from gemini_b200 import field

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - gemini_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = gm.Context(local_rank)
    n = 1 << args.logn

    # workload: rank k owns points [k*n, (k+1)*n) of the global SRS P_i = [i+1]G and its n scalars
    srs = ctx.srs_generate(n, first_multiple=1 + rank * n)
    if not args.no_precompute:
        srs.precompute()  # key setup (untimed, like CommitterKey::new): 2^(c*w) multiples of the SRS in HBM
    pre_c, pre_levels = srs.precompute_info()
    nbuf = 2
    d_scal = [ctx.dev_alloc(n * 32) for _ in range(nbuf)]
    h_scal = []
    for k in range(nbuf):
        ctx.fr_random_dev(d_scal[k], n, 1000 + 7919 * k + 104729 * rank)
        if args.scalars == "equal":  # dummy_r1cs (src/circuit.rs:349-365): every scalar identical
            one = ctx.dev_download(d_scal[k], 32).reshape(1, 4)
            ctx.dev_upload(d_scal[k], np.ascontiguousarray(np.broadcast_to(one, (n, 4))))
        t = torch.empty(n * 4, dtype=torch.int64).pin_memory()
        arr = ctx.dev_download(d_scal[k], n * 32)
        t.copy_(torch.from_numpy(arr.view(np.int64)))
        h_scal.append(t)
    part_dev = torch.zeros(18, dtype=torch.int64, device="cuda")
    gather = [torch.zeros(18, dtype=torch.int64, device="cuda") for _ in range(world)]

    def exchange(partial: np.ndarray) -> np.ndarray:
        """all-reduce of partial G1 accumulators = all-gather of 144-byte points + device adds"""
        if world == 1:
            return partial
        part_dev.copy_(torch.from_numpy(partial.view(np.int64)))
        dist.all_gather(gather, part_dev)
        allp = torch.stack(gather).cpu().numpy().view(np.uint64)
        return ctx.g1_sum(allp)

    def step_resident(k):
        return exchange(ctx.msm_dev(srs, d_scal[k % nbuf], n))

    def step_e2e(k):
        return exchange(ctx.msm(srs, h_scal[k % nbuf]))  # pinned host scalars: H2D inside the call, 144 B D2H

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    results = {}
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~1 s to produce its first line: started before the warm-up steps
    t_warm = time.perf_counter()
    k = 0
    while k < max(args.warmup, 3) or time.perf_counter() - t_warm < 1.5:
        results[("r", k % nbuf)] = step_resident(k)
        k += 1
    # ---- timed: inputs resident in HBM --------------------------------------------------------
    ev_ms, wall_ms, acc_ms, sort_ms, red_ms = [], [], [], [], []
    launches0 = ctx.launch_count
    for k in range(args.steps):
        ctx.l2_flush()  # between timed iterations: 256 MB write > 126 MB L2 (untimed)
        barrier()
        t0 = time.perf_counter()
        ctx.timer_start()
        out = step_resident(k)
        ms = ctx.timer_stop()
        torch.cuda.synchronize()
        wall_ms.append(1e3 * (time.perf_counter() - t0))
        ev_ms.append(ms)
        sort_ms.append(ctx.last_device_ms(1)); acc_ms.append(ctx.last_device_ms(2)); red_ms.append(ctx.last_device_ms(3))
        assert np.array_equal(out, results.setdefault(("r", k % nbuf), out)), "non-deterministic result"
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    # N=1: CUDA events on the library's stream.  N>1: the exchange runs on NCCL's stream, so the step is the
    # synchronised wall time; either way the slowest rank defines the step.
    mine = sum(ev_ms) if world == 1 else sum(wall_ms)
    tot = torch.tensor([mine], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot.item()) / args.steps
    value = world * n / (ms_per_step / 1e3)

    # ---- timed: end to end through the C ABI with host buffers -------------------------------
    for k in range(2):
        results[("e", k % nbuf)] = step_e2e(k)
    e2e_ms = []
    for k in range(args.steps):
        ctx.l2_flush()
        barrier()
        t0 = time.perf_counter()
        out = step_e2e(k)
        torch.cuda.synchronize()
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
        assert np.array_equal(out, results[("e", k % nbuf)])
    tot = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    e2e_value = world * n / (float(tot.item()) / args.steps / 1e3)
    # resident and host paths must agree bit for bit
    assert np.array_equal(results[("e", 0)], results[("r", 0)])

    if rank != 0:
        if dist is not None:
            dist.barrier()
        return
    peak, peak_src = measured_peaks()
    acc = sum(acc_ms) / len(acc_ms)
    achieved = n * ALGO_BYTES_PER_TERM / (acc / 1e3) / 1e9
    plan_c = os.environ.get("GM_MSM_C", "auto")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"kzg::commit n=2^{args.logn} BLS12-381 G1 per GPU (BASELINE.json configs[1] at logn=20)",
                   "bases": "P_i=[i+1]G generated on device, resident in HBM", "scalars": "uniform Fr (splitmix64), 2 alternating sets" if args.scalars == "uniform" else "all scalars equal (dummy_r1cs)",
                   "window_bits": plan_c, "affine_levels": os.environ.get("GM_MSM_AFFINE", "auto"), "srs_precompute": {"window_bits": pre_c, "levels": pre_levels, "hbm_bytes": pre_levels * n * 96},
                   "l2": "256 MB flush between timed iterations",
                   "timing": "CUDA events on the library stream" if world == 1 else "synchronised wall clock incl. NCCL all-gather, max over ranks",
                   "sharding": "contiguous point ranges, all-gather of 144 B partial sums + device adds" if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 144 * world,
                "ms_per_step": float(tot.item()) / args.steps},
        "gpu_launches": launches,
        "phases_ms": {"digits_sort": sum(sort_ms) / len(sort_ms), "bucket_accumulation": acc, "reduce_finish": sum(red_ms) / len(red_ms)},
        "roofline": {"bound": "hbm", "kernel": "bucket accumulation phase: affine levels (k_aff_prepare, k_aff_invert, k_aff_finish) + work list + k_accumulate, "
                                                 "timed as one region by CUDA events on the library stream",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_PHASE_TRAFFIC.get(args.logn) if (args.scalars == "uniform" and not args.no_precompute) else None,
                     "peak_source": peak_src,
                     "note": "MSM is integer-multiplier bound (SURVEY.md 8d): 6-10 Fq products (276 IMAD.WIDE each) per bucket addition, W additions per term; "
                             "ncu: k_aff_finish 78-87 %, k_accumulate 89 % sm throughput (FMA-heavy pipe). The HBM fraction is reported as the contract asks; "
                             "traffic >> algorithmic bytes because every term gathers one 96-B table point PER WINDOW and the affine levels stream their intermediate points"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        lib = load_oracle()
        threads = host_threads()
        cap = min(n, 1 << 20)
        bases = srs.read(0, cap)
        scal = np.ascontiguousarray(ctx.dev_download(d_scal[0], cap * 32).reshape(cap, 4))
        m = pick_sample(lib, bases, scal, cap, threads, target_s=12.0)
        dt, cpu_out, c_used = cpu_msm_timed(lib, bases, scal, m, threads)
        gpu_out = field.jacobian_to_affine(ctx.msm_dev(srs, d_scal[0], m))
        cpu_pt = field.g1_from_limbs(cpu_out.reshape(1, 12))[0]
        line["cpu_baseline"] = {"value": m / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"first 2^{m.bit_length() - 1} terms of the same workload, one MSM, arkworks window c={c_used}, "
                                          f"one thread per window; CPU and GPU results equal: {cpu_pt == gpu_out}"}
        assert cpu_pt == gpu_out, "GPU result differs from the CPU restatement"
    if world == 1 and args.extras:
        # the other BASELINE configs on the same box, as extra keys (never allowed to break the headline line)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_snark
            import bench_sumcheck

            srs.free()
            for p in d_scal:
                ctx.dev_free(p)
            rows = []
            bench_sumcheck.run(ctx, 24, 3, rows.append)
            line["extra"] = {"sumcheck_2^24": rows, "snark_time_prover": bench_snark.run(ctx, args.extras_logn, 2)}
        except Exception as exc:  # pragma: no cover
            line["extra"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--logn", type=int, default=20, help="log2 of the number of MSM terms per GPU")
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--scalars", default="uniform", choices=["uniform", "equal"])
    ap.add_argument("--extras", action="store_true", help="also time config 3 (sumcheck 2^24) and config 4 (snark time prover) into an `extra` key")
    ap.add_argument("--extras-logn", type=int, default=24)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-precompute", action="store_true", help="do not build the table of 2^(c*w) multiples of the SRS")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
