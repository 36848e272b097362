#!/usr/bin/env python
"""bench.py - headline benchmark of the Gemini prover hot path on B200.

Metric (BASELINE.json): G1 MSM throughput in scalar-mults/s.  One "step" = one kzg::commit-equivalent
MSM (VariableBaseMSM::msm_unchecked, /root/reference/src/kzg/time.rs:81-83) over n = 2^logn synthetic
BLS12-381 G1 bases resident on the device and n uniformly random Fr scalars (default logn = 24, the size
the north star is quoted on).

  own arm       python bench.py --gpus N --steps K --warmup W [--logn 24]
                N > 1: launched under torchrun, one rank per GPU.  Each rank owns a contiguous range of n points
                (weak scaling); the exchange of the partial G1 sums is ONE ncclAllGather of 192-byte accumulators
                on the library's own stream (gm_msm_g1_sharded) - torch.distributed only carries the 128-byte
                NCCL id at start-up.  The other BASELINE configs ride along under "extra": configs[1] (2^20),
                config 3 (sumcheck + fold, 2^24 Fr), config 4 (snark time prover, logsize 24), config 5 (streamed
                MSM in 2^20 chunks) and, at N > 1, ONE MSM of fixed total size sharded by point range (strong scaling).
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
# table, uniform scalars), summed from the committed ncu launch lists of the same bench.py command (profiles/):
#   r01: profiles/r01_launches_n20_affine2_dram.csv, profiles/r01_launches_n24_affine4_dram.csv
NCU_PHASE_TRAFFIC = {20: 8.77e9, 24: 148.9e9}
NCU_PHASE_TRAFFIC_SRC = "profiles/r01_launches_n{logn}_affine*_dram.csv"
try:  # refreshed by tools/summarize_launches.py from this round's ncu pass
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as _fh:
        _t = json.load(_fh)
        NCU_PHASE_TRAFFIC.update({int(k): float(v) for k, v in _t.get("phase_traffic_bytes", {}).items()})
        NCU_PHASE_TRAFFIC_SRC = _t.get("source", NCU_PHASE_TRAFFIC_SRC)
except (OSError, ValueError):
    pass


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU baseline (oracle/gemini_oracle.c) - the ONLY place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------
def load_oracle(native: bool = True):
    """The C port, rebuilt for THIS host (-march=native: mulx / adx as arkworks' `asm` feature uses) when possible;
    the portable build (x86-64-v3) that travels with the repo otherwise."""
    odir = os.path.join(ROOT, "oracle")
    so = os.path.join(odir, "libgemini_oracle.so")
    kind = "x86-64-v3"
    if native:
        try:
            subprocess.run(["make", "-C", odir, "native"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
            cand = os.path.join(odir, "_build", "libgemini_oracle_native.so")
            if os.path.exists(cand):
                so, kind = cand, "march=native"
        except Exception:
            pass
    if not os.path.exists(so):
        subprocess.run(["make", "-C", odir], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.go_msm_g1.restype = C.c_int
    lib.go_msm_g1.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    lib.go_generate_bases.argtypes = [C.c_size_t, C.c_uint64, C.c_void_p]
    lib.go_sumcheck_time.restype = C.c_size_t
    lib.go_sumcheck_time.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.build_kind = kind
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
    # values >= r: subtract r limb by limb (borrow chain) on the selected rows
    sel = np.nonzero(ge)[0]
    if sel.size:
        sub = w[sel]
        borrow = np.zeros(sel.size, dtype=np.uint64)
        for j in range(4):
            a = sub[:, j]
            t = a - rl[j]
            b1 = (a < rl[j]).astype(np.uint64)
            t2 = t - borrow
            b2 = (t < borrow).astype(np.uint64)
            sub[:, j] = t2
            borrow = b1 | b2
        w[sel] = sub
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


def workload_name(logn: int) -> str:
    tag = {20: " (BASELINE.json configs[1])", 24: " (the size BASELINE.json's north star is quoted on)"}.get(logn, "")
    return f"kzg::commit n=2^{logn} BLS12-381 G1{tag}"


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
              f"one thread per window (<= {(255 + c_used - 1) // c_used} usable threads); C port built {lib.build_kind}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.logn) + " - CPU restatement of ark-ec msm_unchecked (the Rust reference cannot be built: "
                               "no cargo in the image)",
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
class Job:
    """rank / world / context of this process and the rank-synchronised helpers of the timed loops"""

    def __init__(self):
        import torch

        import gemini_b200 as gm

        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - gemini_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.torch = torch
        self.dist = None
        self.ctx = gm.Context(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist

            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            # the only use of torch.distributed on the data path's behalf: carry the 128-byte NCCL id to every rank
            self.ctx.comm_init_torch(device=f"cuda:{self.local_rank}")

    def barrier(self):
        """all ranks aligned and every queue drained: the library's own collective on its own stream + device sync"""
        self.ctx.comm_barrier()
        self.torch.cuda.synchronize()
        self.ctx.synchronize()

    def max_over_ranks(self, x: float) -> float:
        import numpy as np

        if self.world == 1:
            return x
        rows = self.ctx.comm_allgather(np.array([x], dtype=np.float64).view(np.uint64))
        return float(rows.view(np.float64).max())

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
        self.ctx.close()


def pinned_copy(torch, ctx, d_ptr, n):
    import numpy as np

    t = torch.empty(n * 4, dtype=torch.int64).pin_memory()
    t.copy_(torch.from_numpy(ctx.dev_download(d_ptr, n * 32).view(np.int64)))
    return t


def time_msm(job: Job, srs, d_scal, h_scal, n, steps, warmup, mode):
    """K timed steps of one (sharded) MSM.  mode: "resident" (scalars in HBM, CUDA events on the library stream),
    "pinned" / "pageable" (host scalars through the C ABI, wall clock around the blocking call).
    Returns (sum of step ms - max over ranks, per-phase ms lists, launches, last result)."""
    import numpy as np

    ctx, nbuf = job.ctx, len(d_scal)

    def step(k):
        if mode == "resident":
            return ctx.msm_sharded_dev(srs, d_scal[k % nbuf], n)
        return ctx.msm_sharded(srs, h_scal[k % nbuf], n=n)

    ref = {}
    for k in range(max(warmup, 3)):   # rank-independent count: every rank issues the same collectives
        ref[k % nbuf] = step(k)
    ev_ms, phases = [], ([], [], [])
    launches0 = ctx.launch_count
    for k in range(steps):
        ctx.l2_flush()  # between timed iterations: 256 MB write > 126 MB L2 (untimed)
        job.barrier()
        t0 = time.perf_counter()
        if mode == "resident":
            ctx.timer_start()
            out = step(k)
            ev_ms.append(ctx.timer_stop())
        else:
            out = step(k)
            ev_ms.append(1e3 * (time.perf_counter() - t0))
        for j in range(3):
            phases[j].append(ctx.last_device_ms(j + 1))
        assert np.array_equal(out, ref.setdefault(k % nbuf, out)), "non-deterministic result"
    launches = ctx.launch_count - launches0
    return job.max_over_ranks(sum(ev_ms)), phases, launches, ref


def run_own(args):
    import numpy as np

    from gemini_b200 import field

    job = Job()
    ctx, rank, world, torch = job.ctx, job.rank, job.world, job.torch
    n = 1 << args.logn
    sampler = ClockSampler(job.local_rank)

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
        h_scal.append(pinned_copy(torch, ctx, d_scal[k], n))

    if rank == 0:
        sampler.start()  # nvidia-smi needs ~1 s to produce its first line; no collective depends on it
    # ---- timed: inputs resident in HBM --------------------------------------------------------
    tot_ms, (sort_ms, acc_ms, red_ms), launches, res_r = time_msm(job, srs, d_scal, h_scal, n, args.steps, args.warmup, "resident")
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = tot_ms / args.steps
    value = world * n / (ms_per_step / 1e3)
    # ---- timed: end to end through the C ABI with host buffers (pinned, then plain pageable memory) ----
    e2e_tot, _, _, res_e = time_msm(job, srs, d_scal, h_scal, n, args.steps, 2, "pinned")
    e2e_ms = e2e_tot / args.steps
    assert np.array_equal(res_e[0], res_r[0]), "host and resident paths disagree"
    h_page = [np.array(t.numpy().view(np.uint64).reshape(n, 4)) for t in h_scal]   # ordinary (pageable) numpy memory, like a Rust Vec<Fr>
    page_tot, _, _, res_p = time_msm(job, srs, d_scal, h_page, n, max(2, args.steps // 2), 1, "pageable")
    page_ms = page_tot / max(2, args.steps // 2)
    assert np.array_equal(res_p[0], res_r[0]), "pageable and resident paths disagree"
    del h_page

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        acc = sum(acc_ms) / len(acc_ms)
        achieved = n * ALGO_BYTES_PER_TERM / (acc / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args.logn) + (" per GPU" if world > 1 else ""),
                       "bases": "P_i=[i+1]G generated on device, resident in HBM",
                       "scalars": "uniform Fr (splitmix64), 2 alternating sets" if args.scalars == "uniform" else "all scalars equal (dummy_r1cs)",
                       "window_bits": os.environ.get("GM_MSM_C", "auto"), "affine_levels": os.environ.get("GM_MSM_AFFINE", "auto"),
                       "srs_precompute": {"window_bits": pre_c, "levels": pre_levels, "hbm_bytes": pre_levels * n * 96},
                       "l2": "256 MB flush between timed iterations",
                       "timing": "CUDA events on the library stream (the NCCL all-gather of the partial sums is queued on that stream), max over ranks",
                       "sharding": f"{world} contiguous point ranges, one ncclAllGather of 192-B partial accumulators + {world - 1} device adds per MSM"
                                   if world > 1 else "single GPU"},
            "e2e": {"value": world * n / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 144 * world,
                    "ms_per_step": e2e_ms, "host_memory": "pinned",
                    "pageable": {"value": world * n / (page_ms / 1e3), "ms_per_step": page_ms,
                                 "note": "same call from ordinary pageable host memory (what an arkworks &[Fr] is)"}},
            "gpu_launches": launches,
            "phases_ms": {"digits_sort": sum(sort_ms) / len(sort_ms), "bucket_accumulation": acc, "reduce_finish": sum(red_ms) / len(red_ms)},
            "roofline": {"bound": "hbm", "kernel": "bucket accumulation phase: affine levels (k_aff_prepare, k_aff_invert, k_aff_finish) + work list + k_accumulate, "
                                                     "timed as one region by CUDA events on the library stream",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_PHASE_TRAFFIC.get(args.logn) if (args.scalars == "uniform" and not args.no_precompute) else None,
                         "traffic_source": NCU_PHASE_TRAFFIC_SRC.format(logn=args.logn),
                         "peak_source": peak_src,
                         "note": "MSM is integer-multiplier bound (SURVEY.md 8d): 6-10 Fq products (276 IMAD.WIDE each) per bucket addition, W additions per term. "
                                 "The HBM fraction is reported as the contract asks; traffic >> algorithmic bytes because every term gathers one 96-B table point "
                                 "PER WINDOW and the affine levels stream their intermediate points"},
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
                                          f"one thread per window, C port built {lib.build_kind}; CPU and GPU results equal: {cpu_pt == gpu_out}"}
        assert cpu_pt == gpu_out, "GPU result differs from the CPU restatement"
        if lib.build_kind != "x86-64-v3":   # round 1's portable build next to it, so that the ratio is honest
            old = load_oracle(native=False)
            dt_old, _, _ = cpu_msm_timed(old, bases, scal, m, threads)
            line["cpu_baseline"]["portable_build_value"] = m / dt_old

    srs.free()
    for p in d_scal:
        ctx.dev_free(p)
    del h_scal
    if not args.no_extras:
        extra = {}
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_extras

        for name, fn in bench_extras.plan(job, args):
            try:   # an extra must never break the headline line
                job.barrier()
                extra[name] = fn()
            except Exception as exc:  # pragma: no cover
                extra[name] = {"error": repr(exc)}
                if world > 1:
                    raise          # a rank that skips collectives would hang the others: fail loudly instead
        if rank == 0:
            line["extra"] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    job.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--logn", type=int, default=24, help="log2 of the number of MSM terms per GPU")
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--scalars", default="uniform", choices=["uniform", "equal"])
    ap.add_argument("--no-extras", action="store_true", help="headline line only (skip configs[1], 3, 4, 5 and the strong-scaling MSM)")
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
