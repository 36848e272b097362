"""The BASELINE configs that ride along with the headline line of bench.py (its "extra" key).

  N = 1   msm_2^20 (configs[1]), sumcheck_2^24 (config 3, with its own roofline and single-thread CPU baseline),
          snark_time_prover (config 4 on one GPU), streamed_msm (config 5's per-GPU share), strong_scaling baseline
  N > 1   strong_scaling (ONE MSM of fixed total size sharded by point range), streamed_msm (config 5 shape:
          2^25 scalars per GPU in 2^20 chunks = logsize 28 on 8 GPUs)

Every function returns a JSON-able dict (rank 0's view); all ranks must call them in the same order (collectives)."""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def plan(job, args):
    """[(name, thunk)] in execution order"""
    out = []
    if job.world == 1:
        if args.logn != 20:
            out.append(("msm_2^20_configs1", lambda: msm_line(job, 20, 10)))
        out.append(("sumcheck_2^24_config3", lambda: sumcheck_rows(job, 24)))
        out.append(("snark_time_prover_config4", lambda: snark_line(job, args.extras_logn)))
        out.append(("snark_elastic_prover_1gpu", lambda: elastic_line(job, args.extras_logn)))
    else:
        out.append(("snark_time_prover_config4_sharded", lambda: sharded_snark_line(job, args.extras_logn)))
        out.append(("snark_elastic_prover_config5_sharded", lambda: sharded_elastic_line(job, args.extras_logn)))
    out.append(("streamed_msm_config5", lambda: streamed_line(job, 24 if job.world == 1 else 25, 20)))
    out.append(("strong_scaling_msm_2^24", lambda: strong_line(job, 24, 5)))
    out.append(("strong_scaling_msm_2^26", lambda: strong_line(job, 26, 3)))
    return out


def _scalars(job, n, seed):
    import bench

    ctx = job.ctx
    d = ctx.dev_alloc(n * 32)
    ctx.fr_random_dev(d, n, seed)
    h = bench.pinned_copy(job.torch, ctx, d, n)
    return d, h


def msm_line(job, logn, steps):
    """one more single-GPU MSM size, same procedure as the headline"""
    import bench

    ctx = job.ctx
    n = 1 << logn
    srs = ctx.srs_generate(n, first_multiple=1)
    srs.precompute()
    d0, h0 = _scalars(job, n, 1000)
    d1, h1 = _scalars(job, n, 8919)
    tot, (s_ms, a_ms, r_ms), launches, _ = bench.time_msm(job, srs, [d0, d1], [h0, h1], n, steps, 3, "resident")
    e2e, _, _, _ = bench.time_msm(job, srs, [d0, d1], [h0, h1], n, steps, 2, "pinned")
    peak, _ = bench.measured_peaks()
    acc = sum(a_ms) / len(a_ms)
    res = {"workload": bench.workload_name(logn), "value": n / (tot / steps / 1e3), "unit": bench.UNIT, "ms_per_step": tot / steps,
           "e2e": {"value": n / (e2e / steps / 1e3), "ms_per_step": e2e / steps, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 144},
           "phases_ms": {"digits_sort": sum(s_ms) / len(s_ms), "bucket_accumulation": acc, "reduce_finish": sum(r_ms) / len(r_ms)},
           "roofline_frac_hbm": n * bench.ALGO_BYTES_PER_TERM / (acc / 1e3) / 1e9 / peak, "gpu_launches_per_step": launches // steps,
           "srs_precompute": srs.precompute_info()}
    srs.free()
    ctx.dev_free(d0); ctx.dev_free(d1)
    return res


def strong_line(job, log_total, steps):
    """ONE MSM of 2^log_total terms in total, sharded by contiguous point range over the ranks (strong scaling)."""
    import bench

    ctx, world, rank = job.ctx, job.world, job.rank
    n = (1 << log_total) // world
    srs = ctx.srs_generate(n, first_multiple=1 + rank * n)
    srs.precompute()
    d0, h0 = _scalars(job, n, 31337 + rank)
    tot, (s_ms, a_ms, r_ms), launches, _ = bench.time_msm(job, srs, [d0], [h0], n, steps, 3, "resident")
    res = {"workload": f"one MSM of 2^{log_total} terms sharded by point range over {world} GPU(s): 2^{log_total}/{world} terms per rank",
           "scaling": "strong", "value": world * n / (tot / steps / 1e3), "unit": bench.UNIT, "ms_per_step": tot / steps, "n_gpus": world,
           "phases_ms_rank0": {"digits_sort": sum(s_ms) / len(s_ms), "bucket_accumulation": sum(a_ms) / len(a_ms),
                               "reduce_finish_exchange": sum(r_ms) / len(r_ms)},
           "timing": "CUDA events on the library stream incl. the ncclAllGather, max over ranks", "srs_precompute": srs.precompute_info()}
    srs.free()
    ctx.dev_free(d0)
    return res


def streamed_line(job, logn, chunk_log, reps=3):
    """msm_chunks (src/kzg/space.rs:22-55): every rank streams 2^logn host scalars in chunks of 2^chunk_log against its
    resident SRS range; ONE exchange at finalize.  At 8 ranks and logn 25 this is BASELINE config 5 (logsize 28)."""
    import numpy as np

    from gemini_b200.msm import _DeviceStream

    ctx, world, rank = job.ctx, job.world, job.rank
    n, chunk = 1 << logn, 1 << chunk_log
    srs = ctx.srs_generate(n, first_multiple=1 + rank * n)
    srs.precompute(expected_msm_len=n)
    d, host = _scalars(job, n, 4242 + rank)
    want = ctx.msm_sharded_dev(srs, d, n)   # one-shot sharded MSM over resident scalars: the streamed result must be identical
    ctx.dev_free(d)
    hv = host.view(-1, 4)
    best = None
    for _ in range(reps):
        job.barrier()
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        st = _DeviceStream(ctx, srs, chunk)
        for s0 in range(0, n, chunk):
            st.push_range(s0, hv[s0:s0 + chunk])
        total = st.finalize_sharded_raw()
        dt = job.max_over_ranks(time.perf_counter() - t0)
        st.free()
        assert np.array_equal(total, want), "streamed result differs from the one-shot MSM"
        if best is None or dt < best[0]:
            best = (dt, ctx.launch_count - l0)
    dt, launches = best
    res = {"workload": f"msm_chunks: 2^{logn} scalars per GPU from pinned host memory in 2^{chunk_log}-scalar chunks, SRS range resident"
                       + (" = BASELINE config 5 (logsize 28, 8 GPUs)" if (world == 8 and logn == 25) else ""),
           "value": world * n / dt, "unit": "scalar-mults/s", "n_gpus": world, "wall_s": dt, "chunks_per_gpu": n // chunk,
           "gpu_launches_rank0": launches, "h2d_bytes": n * 32 * world, "srs_precompute": srs.precompute_info(),
           "streamed == one-shot": True, "timing": "wall clock around push* + finalize (H2D inside), max over ranks"}
    srs.free()
    return res


def sumcheck_rows(job, logn):
    """config 3: fold + full TimeProver runs on 2^logn Fr elements, with the HBM roofline of SURVEY.md 8(d)
    (256*n bytes for a full sumcheck, 48 B per input element for one fold) and a single-thread CPU baseline."""
    import numpy as np

    import bench
    import bench_sumcheck

    rows = []
    bench_sumcheck.run(job.ctx, logn, 4, rows.append)
    res = {"workload": f"subprotocols::sumcheck TimeProver + fold_polynomial, |f| = |g| = 2^{logn} Fr, challenges fed back from the host", "rows": rows}
    for r in rows:
        if r.get("prover") == "gemini twist=1":
            peak, src = bench.measured_peaks()
            res["roofline"] = {"bound": "hbm", "kernel": "k_sc_message + k_sc_fold_message (all rounds, CUDA events on the prover's stream)",
                               "achieved": r["GBps_at_256n"], "peak": peak, "unit": "GB/s", "frac": r["frac_of_hbm_peak"],
                               "algorithmic_bytes": 256 << logn, "peak_source": src}
    # CPU: the reference's TimeProver is single-threaded (time_prover.rs:75-123) - bounded sample of 2^20 elements
    try:
        lib = bench.load_oracle()
        m = 1 << 20
        f = bench.splitmix_scalars(m, 1)
        g = bench.splitmix_scalars(m, 2)
        tw = bench.splitmix_scalars(1, 3)
        ch = bench.splitmix_scalars(24, 4)
        msgs = np.zeros((24, 8), dtype=np.uint64)
        fin = np.zeros(8, dtype=np.uint64)
        t0 = time.perf_counter()
        lib.go_sumcheck_time(f.ctypes.data, m, g.ctypes.data, m, tw.ctypes.data, ch.ctypes.data, 24, msgs.ctypes.data, fin.ctypes.data)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 2 * m / dt, "unit": "Fr elements/s", "cores": 1, "kind": "port",
                               "sample": f"full TimeProver (random twist) on |f| = |g| = 2^20, C port built {lib.build_kind}; the reference's prover is single-threaded"}
    except Exception as exc:  # pragma: no cover
        res["cpu_baseline"] = {"error": repr(exc)}
    return res


def snark_line(job, logn):
    import bench_snark

    return bench_snark.run(job.ctx, logn, 2)


def sharded_snark_line(job, logn):
    import bench_snark

    return bench_snark.run_sharded(job, logn, 2)


def sharded_elastic_line(job, logn):
    import bench_snark

    return bench_snark.run_sharded_elastic(job, logn, 2)   # best of two: the first run creates the helper contexts / scratch


def elastic_line(job, logn):
    import bench_snark

    return bench_snark.run_elastic(job.ctx, logn, 2)   # best of two: the first run creates the helper contexts / scratch
