#!/usr/bin/env python
"""bench.py -- headline benchmark of the Genome-on-Diet hot path on B200 (see BASELINE.json).

Workload (config 2 of BASELINE.json, the one the metric is quoted on): batched ksw_extd2 over 1M
synthetic 150x200 bp query/target pairs, band 150, sr scoring, Z-drop on (flag 0: exact max + Z-drop),
CIGARs produced.  A "step" is one pass of the hot path (pack -> DP -> traceback) over the whole batch.

  value      GCUPS with the inputs already resident in HBM (CUDA events on the library's stream)
  e2e        GCUPS through the host-buffer C-ABI call gd_ksw_extd2_batch (pinned host buffers; H2D,
             kernels, D2H of ksw_extz_t records and CIGARs inside the timed region)
  roofline   integer-ALU roofline of the DP kernel from its own launch durations (CUDA events around every
             launch, on the library's stream): SURVEY.md 8d's 51 lane-ops per banded cell against the packed
             16x2 add rate measured live by gd_ubench; plus the backtrack bytes against MEASURED_PEAKS
  cpu_baseline  the unmodified reference (oracle/_ref, ksw_extd2_avx512 when the host has AVX-512) on all
             host threads over a bounded sample of the same pairs
  parity     every pair of the step compared with the reference's result (all ksw_extz_t fields + CIGARs)
  extra      zdrop_subset (30 % edits: Z-drop fires; parity on every pair), sketch (mm_sketch of 200 Mbp: kernel,
             host-buffer end to end, read sketching, CPU baselines), sr_map / lr_map (configs 1 / 3 end to end at N = 1,
             SAM compared with GDiet_avx run in the same process, the batched C host binary), map_strong (10 M reads
             against a 3.1 Gbp index, STRONG-scaled over the ranks: index built on rank 0, NCCL broadcast, read
             shards, SAM text made on the devices and gathered in input order, hash compared with one GPU)

`--impl reference` times only that CPU reference arm.  With torchrun (N>1) every rank maps its own
1M-pair shard for the headline metric (weak scaling, no data-path collective); time is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

QLEN, TLEN, BAND = 150, 200, 150
OPS_PER_CELL = 51          # SSE4.1 lane-ops per cell with backtrack
BYTES_PER_CELL = 1.0       # backtrack byte


def band_prefix(qlen, tlen, w):
    """cells executed after r rows, for r = 0..qlen+tlen-1 (SURVEY.md 8d definition of a banded cell)"""
    pre = [0]
    for r in range(qlen + tlen - 1):
        st0 = max(0, r - qlen + 1, (r - w + 1) >> 1)
        en0 = min(tlen - 1, r, (r + w) >> 1)
        pre.append(pre[-1] + max(0, en0 - st0 + 1))
    return np.array(pre, np.int64)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line).  The values are read through
    NVML in-process -- the library nvidia-smi itself reads -- because spawning an `nvidia-smi` process every 100 ms inside
    a 0.4 s timed region stalled kernel launches by 6-16 ms per step on a fresh box (the kernels' own times did not move);
    `nvidia-smi` is the fallback when the NVML binding is missing."""

    REASONS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            return [str(sm), str(self.max_sm)] + ["Active" if mask & bit else "Not Active" for _, bit in self.REASONS]
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [x.strip() for x in out.split(",")] if out else None

    def run(self):
        while not self.stop_flag:
            try:
                row = self.sample()
                if row:
                    self.rows.append(row)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = [n for n, _ in self.REASONS]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_pairs(n, seed):
    import gdiet_b200  # noqa: F401
    from gdiet_b200 import synth
    return synth.ksw_pairs_fast(n, QLEN, TLEN, 0.05, seed=seed)


def reference_arm(args, rank, world):
    """CPU: the unmodified reference (oracle/_ref) on all host threads; bounded sample per step."""
    if rank != 0:
        return
    from oraclelib import Ref, cpu_has_avx512
    import gdiet_b200  # noqa: F401
    from gdiet_b200 import synth
    variant = "avx" if cpu_has_avx512() else "scalar"
    R = Ref(variant)
    cores = host_threads()
    sc = synth.SCORING["sr"]
    mat = synth.score_matrix(sc["a"], sc["b"])
    pre = band_prefix(QLEN, TLEN, BAND)
    # calibrate the sample so one step is ~3 s of wall time on this host
    probe = make_pairs(4096, seed=3)
    t0 = time.perf_counter()
    R.ksw_extd2_batch(probe["qlen"], probe["qoff"], probe["qbuf"], probe["tlen"], probe["toff"], probe["tbuf"], mat, sc["q"], sc["e"],
                      sc["q2"], sc["e2"], BAND, sc["zdrop"], sc["end_bonus"], args.flag, cores, cigar_stride=QLEN + TLEN)
    dt = time.perf_counter() - t0
    n = int(min(args.pairs, max(4096, 4096 * 3.0 / max(dt, 1e-3))))
    P = make_pairs(n, seed=3)
    times, cells = [], 0
    from oraclelib import EXTZ_DTYPE
    pre_out = (np.zeros(n, EXTZ_DTYPE), np.zeros(n * (QLEN + TLEN), np.uint32))  # allocated (and touched) outside the timed steps
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ez, _ = R.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], mat, sc["q"], sc["e"], sc["q2"],
                                  sc["e2"], BAND, sc["zdrop"], sc["end_bonus"], args.flag, cores, cigar_stride=QLEN + TLEN, out=pre_out)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    # rows executed: the reference does not report them; without Z-drop every pair runs all rows, with
    # Z-drop we count full rows for pairs that did not drop and use the oracle-free bound for the rest
    cells = int(n * pre[-1])
    tot = sum(times)
    value = cells * len(times) / tot / 1e9
    line = {"impl": "reference", "metric": "ksw_extd2 GCUPS", "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": workload_config(args, args.pairs),  # same workload as our arm; each step times a bounded sample of it
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": "reference",
                             "sample": "%d of the %d pairs per step, ksw_extd2_%s, flag %#x, cells counted as full band" % (
                                 n, args.pairs, "avx512" if variant == "avx" else "sse", args.flag)},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": "BASELINE config 2: batched ksw_extd2 microbench, %d synthetic %dx%d bp pairs per GPU, band %d, sr scoring "
                        "(a2 b8 q12 e2 q2 24 e2 1 zdrop100 end_bonus10), flag %#x (%s), 5%% edits, N every 50th query, CIGAR on" % (
                            n, QLEN, TLEN, BAND, args.flag, "exact max + Z-drop" if not (args.flag & 8) else "approx max"),
            "pairs_per_gpu": n, "qlen": QLEN, "tlen": TLEN, "band": BAND, "flag": args.flag,
            "cache": "inputs + backtrack arena per step >> 126 MB L2 (no flush needed)"}


def run_ubench():
    """integer pipe rates measured live on this GPU (gd_ubench); returns {op: lane_ops_per_clk_per_sm}"""
    exe = os.path.join(ROOT, "genome-on-diet_b200", "lib", "gd_ubench")
    out = {}
    try:
        txt = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        for ln in txt.splitlines():
            d = json.loads(ln)
            if "op" in d:
                out[d["op"]] = max(out.get(d["op"], 0.0), d["lane_ops_per_clk_per_sm"])
    except Exception:
        pass
    return out


def ours(args, rank, world, local_rank):
    import torch
    import gdiet_b200 as gd
    from gdiet_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # ---- BASELINE configs 1 and 3 (bounded: 100 k short reads / 500 HiFi reads) end to end, rank 0 at N = 1.  First, with a
    # context of their own that is closed again: the two tools run whole PROGRAMS (the unmodified reference and the batched C
    # host) as child processes, and a child next to a parent that already holds the DP arenas of the headline benchmark would
    # get what is left of the HBM.
    srm = lrm = None
    if not args.no_sketch and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        ctx_tools = gd.Context(local_rank)
        try:
            import sr_map_bench
            srm = sr_map_bench.run(ctx_tools, 5, 100_000, run_ref=not args.no_cpu)
        except Exception as e:
            srm = {"error": str(e)}
        try:
            import lr_map_bench
            lrm = lr_map_bench.run(ctx_tools, "hifi", 50, 500, 0, run_ref=not args.no_cpu)
        except Exception as e:
            lrm = {"error": str(e)}
        ctx_tools.close()
    ctx = gd.Context(local_rank)
    ctx.set_option("ksw_group", args.group)
    ctx.set_option("ksw_blocks_per_sm", args.blocks_per_sm)
    ctx.set_option("time_kernels", 1)  # CUDA events around every DP kernel launch, on the library's own stream
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    n = args.pairs
    sc = synth.SCORING["sr"]
    prm = gd.KswParams(synth.score_matrix(sc["a"], sc["b"]), sc["q"], sc["e"], sc["q2"], sc["e2"], sc["zdrop"], sc["end_bonus"], args.flag)
    P = make_pairs(n, seed=3 + rank)  # every rank maps its own shard
    pre = band_prefix(QLEN, TLEN, BAND)

    # ---- device-resident arm -------------------------------------------------------------------
    dev = torch.device("cuda", local_rank)
    d = {k: torch.from_numpy(P[k]).to(dev) for k in ("qlen", "qoff", "qbuf", "tlen", "toff", "tbuf")}
    stride = QLEN + TLEN
    d_ez = torch.zeros(n * 16, dtype=torch.int32, device=dev)
    d_cig = torch.zeros(n * stride, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step_device():
        ctx.ksw_extd2_batch_device(n, d["qlen"], d["qoff"], d["qbuf"], d["tlen"], d["toff"], d["tbuf"], prm, QLEN, TLEN, BAND, d_ez,
                                   d_cig, stride, w_all=BAND)

    # pre-warm: the GPU idles at 120 MHz while the host generates data; spin it up for ~1.5 s so the W
    # warm-up steps and the timed steps all run at steady clocks
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < float(os.environ.get("GD_BENCH_PREWARM_S", "1.5")):
        step_device()
        stream.synchronize()
    for _ in range(args.warmup):
        step_device()
    stream.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.stat("kernel_launches")
    ctx.stat("ksw_dp_reset")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    e1.synchronize()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.stat("kernel_launches") - l0
    dp_us, dp_n = ctx.stat("ksw_dp_us"), ctx.stat("ksw_dp_launches")  # device time of the DP kernel alone
    ctx.set_option("time_kernels", 0)
    ez = d_ez.cpu().numpy().view(gd.GD_EXTZ_DTYPE)
    cells_step = int(pre[np.clip(ez["rows_done"], 0, len(pre) - 1)].sum())

    # ---- end-to-end arm: host buffers through the C ABI ------------------------------------------
    hp = {k: torch.from_numpy(P[k]).pin_memory() for k in ("qlen", "qoff", "qbuf", "tlen", "toff", "tbuf")}
    out = {"ez": torch.zeros(n * 16, dtype=torch.int32).pin_memory().numpy().view(gd.GD_EXTZ_DTYPE),
           "cigar_off": torch.zeros(n + 1, dtype=torch.int64).pin_memory().numpy(),
           "cigar": torch.zeros(n * 24, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}

    def step_e2e():
        return ctx.ksw_extd2_batch(hp["qlen"].numpy(), hp["qoff"].numpy(), hp["qbuf"].numpy(), hp["tlen"].numpy(), hp["toff"].numpy(),
                                   hp["tbuf"].numpy(), prm, w_all=BAND, out=out)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ez_h, coff_h, _ = step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = int(sum(P[k].nbytes for k in ("qlen", "qoff", "qbuf", "tlen", "toff", "tbuf")))
    d2h = int(n * 64 + (n + 1) * 8 + int(coff_h[n]) * 4)
    assert np.array_equal(ez_h["score"], ez["score"]), "device-resident and host-buffer arms disagree"
    cig_h = out["cigar"]

    # ---- max over ranks ------------------------------------------------------------------------
    # time = MAX over ranks, work = SUM over ranks (genome-on-diet_b200/shard.py; no data-path collective)
    from gdiet_b200 import shard
    (dev_ms, e2e_ms), (tot_cells, launches, h2d, d2h) = shard.reduce_timing([dev_ms, e2e_ms], [cells_step, launches, h2d, d2h],
                                                                              device=dev if dist else "cpu")
    # ---- the mapping path, STRONG-scaled over the ranks (every rank takes part; see tools/map_strong_bench.py) ------------
    ms_out = None
    if not args.no_map_strong:
        del d, d_ez, d_cig
        torch.cuda.empty_cache()
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import map_strong_bench
            ms_out = map_strong_bench.run(ctx, rank, world, dist, dev, args.map_reads, args.map_genome_bp, cores=host_threads())
        except Exception as e:  # (stage failures are agreed on by all ranks inside run(): nobody is left waiting)
            ms_out = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    ms_per_step = dev_ms / args.steps
    value = tot_cells / (ms_per_step * 1e-3) / 1e9
    e2e_val = tot_cells / (e2e_ms / args.steps * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the DP kernel), from its own launch durations ----------
    # achieved = algorithmic lane-ops per launch (51 per banded cell x cells of the launch)
    #            / average launch duration (CUDA events on the library's stream, rank 0's launches);
    # peak     = issue rate of the packed 16x2 integer add the kernel is built from, measured live by
    #            gd_ubench on this GPU (MEASURED_PEAKS.json has no integer entry) x SMs x SM clock under load.
    peaks = measured_peaks()
    ub = run_ubench()
    clk = sampler.summary()
    sms = ctx.stat("device_sms")
    lane_rate = ub.get("VIADD.16x2")
    f_hz = (clk["sm_mhz"] or (peaks or {}).get("sm_max_mhz") or 1965.0) * 1e6
    dp_ms = dp_us / 1e3 / max(dp_n, 1)
    cells_launch = cells_step * args.steps / max(dp_n, 1)  # this rank's cells per DP kernel launch
    dp_cells_s = cells_launch / (dp_ms * 1e-3) if dp_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "dp_traffic.json")))
        traffic = {"bytes_per_launch": tj["dram_bytes_per_pair"][str(args.flag)] * n / max(1, ctx.stat("ksw_chunks")),
                   "source": tj["source"]}
    except Exception:
        pass
    roof = {"bound": "int_alu", "kernel": "gd_ksw_dp_kernel", "unit": "Tlaneop/s",
            "achieved": dp_cells_s * OPS_PER_CELL / 1e12,
            "peak": (lane_rate * sms * f_hz / 1e12) if lane_rate else None,
            "peak_source": ("gd_ubench VIADD.16x2 lane-ops/clk/SM measured live (%.1f) x %d SMs x median SM clock under load "
                            "(%.0f MHz); nominal issue limit is 128" % (lane_rate, sms, f_hz / 1e6)) if lane_rate else "unmeasured",
            "ops_per_cell": OPS_PER_CELL, "kernel_ms_per_launch": dp_ms, "kernel_launches": int(dp_n),
            "kernel_share_of_step": (dp_us / 1e3) / dev_ms if dev_ms > 0 else None,
            "kernel_gcups": dp_cells_s / 1e9,
            "traffic": traffic,
            "hbm": {"achieved_gbs": dp_cells_s * BYTES_PER_CELL / 1e9, "peak_gbs": (peaks or {}).get("hbm_gbs", 6650.0),
                    "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6.65 TB/s"},
            "ubench": ub}
    roof["frac"] = (roof["achieved"] / roof["peak"]) if roof["peak"] else None
    roof["hbm"]["frac"] = roof["hbm"]["achieved_gbs"] / roof["hbm"]["peak_gbs"]

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ------------------------------------------
    cpu = parity = None
    if world == 1 and not args.no_cpu:
        try:
            cpu, parity = cpu_baseline_sample(args, P, ez_h, coff_h, cig_h)
        except Exception as e:  # the checker being absent must not hide the GPU number
            cpu = {"value": None, "unit": "GCUPS", "cores": host_threads(), "kind": "unavailable", "sample": str(e)}

    zd = None
    try:
        zd = zdrop_subset(ctx, args)
    except Exception as e:
        zd = {"error": str(e)}
    sk = None
    if not args.no_sketch:
        try:
            sk = sketch_extra(ctx, stream, dev, peaks)
            if not args.no_cpu:  # (rank 0, at every N)
                try:
                    sk["cpu_baseline"] = sketch_cpu_baseline()
                except Exception as e:
                    sk["cpu_baseline"] = {"kind": "unavailable", "sample": str(e)}
        except Exception as e:
            sk = {"error": str(e)}
    line = {"metric": "ksw_extd2 GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": workload_config(args, n),
            "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "extra": {"ksw_group_lanes": ctx.stat("ksw_group"), "ksw_ring_columns": ctx.stat("ksw_ring"), "chunks_per_step": ctx.stat("ksw_chunks"),
                      "zdropped_frac": float((ez["zdropped"] != 0).mean()), "cells_per_step": tot_cells, "zdrop_subset": zd, "sketch": sk,
                      "sr_map": srm, "lr_map": lrm, "map_strong": ms_out}}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def sketch_cpu_baseline():
    """reference mm_sketch (AVX-512 build of oracle/_ref when the host has it) on all host threads: contigs of 2 Mbp,
    repeated for ~5 s"""
    from oraclelib import Ref, cpu_has_avx512
    import gdiet_b200  # noqa: F401
    from gdiet_b200 import synth
    variant = "avx" if cpu_has_avx512() else "scalar"
    R = Ref(variant)
    cores = host_threads()
    L, nc = 250_000, cores * 32  # the glue hands out 16 sequences per grab: 2 grabs per thread
    g = synth.random_genome(L * nc, seed=11).tobytes()
    off = np.arange(nc, dtype=np.int64) * L
    lens = np.full(nc, L, np.int32)
    R.mm_sketch_batch(off, lens, g, 11, 21, "10", cores)
    tot, passes = 0.0, 0
    while tot < 5.0 and passes < 64:
        t0 = time.perf_counter()
        R.mm_sketch_batch(off, lens, g, 11, 21, "10", cores)
        tot += time.perf_counter() - t0
        passes += 1
    out = {"value": passes * L * nc / tot / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "reference",
           "sample": "%d passes of mm_sketch (%s build) over %d contigs of %d bp on %d threads, %.1f s" % (passes, variant, nc, L, cores, tot)}
    # ... and on 150 bp reads (one mm_sketch call per read, all threads)
    rl = 150
    nr = min(2_000_000, len(g) // rl)
    r_off = np.arange(nr, dtype=np.int64) * rl
    r_len = np.full(nr, rl, np.int32)
    R.mm_sketch_batch(r_off, r_len, g, 11, 21, "10", cores)
    tot, passes = 0.0, 0
    while tot < 3.0 and passes < 64:
        t0 = time.perf_counter()
        R.mm_sketch_batch(r_off, r_len, g, 11, 21, "10", cores)
        tot += time.perf_counter() - t0
        passes += 1
    out["reads"] = {"value": passes * nr * rl / tot / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "reference",
                    "sample": "%d passes of ONE mm_sketch call per 150 bp read over %d reads on %d threads (mm_sketch2 + mm_sketch3 make ~2.1 such passes per read)" % (
                        passes, nr, cores)}
    return out


def sketch_extra(ctx, stream, dev, peaks):
    """Second half of the hot path (BASELINE config 5 shape, reduced): sparsified index sketching (mm_sketch,
    -Z 10 -W 2 -k 21 -w 11) of a 200 Mbp synthetic genome resident in HBM.  Reported in `extra`, with the HBM
    roofline fraction of the sketch kernel (algorithmic bytes = 1 B per input base + 16 B per minimizer)."""
    import torch
    L, nc = 25_000_000, 8
    g = torch.randint(0, 4, (nc * L,), dtype=torch.uint8, device=dev)
    seq = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[g.long()]
    del g
    d_off = torch.arange(nc, dtype=torch.int64, device=dev) * L
    d_len = torch.full((nc,), L, dtype=torch.int32, device=dev)
    d_rid = torch.arange(nc, dtype=torch.int32, device=dev)
    cap = nc * L // 5
    d_out = torch.zeros(cap * 2, dtype=torch.int64, device=dev)
    d_oo = torch.zeros(nc + 1, dtype=torch.int64, device=dev)
    ctx.set_option("time_kernels", 1)
    best_us, best_dt = None, None
    for it in range(5):
        ctx.stat("sketch_reset")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.sketch_ref_batch_device(nc, d_off, d_len, d_rid, seq, nc * L, 11, 21, "10", d_oo, d_out, cap)
        stream.synchronize()
        dt = time.perf_counter() - t0
        us = ctx.stat("sketch_us")
        if it >= 2 and (best_us is None or us < best_us):
            best_us, best_dt = us, dt
    ctx.set_option("time_kernels", 0)
    nmin, bases = int(d_oo[-1].item()), nc * L
    algo = bases + 16 * nmin
    hbm = (peaks or {}).get("hbm_gbs", 6650.0)
    out = {"workload": "mm_sketch of %d x %d bp synthetic contigs, -Z 10 -W 2 -k 21 -w 11, device resident" % (nc, L),
           "gbases_per_s_call": bases / best_dt / 1e9, "gbases_per_s_kernel": bases / (best_us * 1e-6) / 1e9, "minimizers": nmin,
           "roofline": {"bound": "hbm", "achieved": algo / (best_us * 1e-6) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": algo / (best_us * 1e-6) / 1e9 / hbm,
                        "note": "integer work (ncu: ~100 lane-instructions per original base = ~200 per sparsified position with the v3 tile body; ALU pipe 68 % busy, 3.5 of 12.9 stall cycles per issued instruction at block barriers) keeps this kernel ALU / barrier bound, far below the HBM roofline"}}
    # ---- end to end through the host-buffer calls (pinned host ASCII in, minimizers out on the host) ----------------------
    import ctypes as C
    import gdiet_b200 as gd
    lib = ctx.lib
    h_seq = torch.empty(nc * L, dtype=torch.uint8).pin_memory()
    h_seq.copy_(seq)
    del seq, d_out
    h_off = np.arange(nc, dtype=np.int64) * L
    h_len = np.full(nc, L, np.int32)
    h_rid = np.arange(nc, dtype=np.uint32)
    h_oo = np.zeros(nc + 1, np.int64)
    h_out = torch.zeros(cap * 2, dtype=torch.int64).pin_memory().numpy()
    best = None
    for it in range(4):
        t0 = time.perf_counter()
        rc = lib.gd_sketch_ref_batch(ctx.h, nc, gd._ptr(h_off), gd._ptr(h_len), gd._ptr(h_rid), gd._ptr(h_seq.numpy()), 11, 21, b"10", 2,
                                     gd._ptr(h_oo), gd._ptr(h_out), cap)
        dt = time.perf_counter() - t0
        ctx._check(rc, "gd_sketch_ref_batch")
        if it and (best is None or dt < best):
            best = dt
    out["e2e"] = {"value": bases / best / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": int(bases + nc * 16), "d2h_bytes_per_step": int(h_oo[nc]) * 16,
                  "call": "gd_sketch_ref_batch: pinned host ASCII -> mm128_t records on the host", "identical_to_device_arm": int(h_oo[nc]) == nmin}
    del h_seq, h_out
    # ---- read sketching (mm_sketch2 + mm_sketch3 of every shift, the per-read calls of map.c:74-99) for 2 M reads of 150 bp
    nr, rl = 2_000_000, 150
    g = torch.randint(0, 4, (nr * rl,), dtype=torch.uint8, device=dev)
    h_reads = torch.empty(nr * rl, dtype=torch.uint8).pin_memory()
    h_reads.copy_(torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[g.long()])
    del g
    r_off = np.arange(nr, dtype=np.int64) * rl
    r_len = np.full(nr, rl, np.int32)
    W = 2
    s2_counts = torch.zeros(nr * W, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    s3_ret = torch.zeros(nr * W, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    s2_off = torch.zeros(nr + 1, dtype=torch.int64).pin_memory().numpy()
    s3_off = torch.zeros(nr * W + 1, dtype=torch.int64).pin_memory().numpy()
    rcap = nr * 40
    s2 = torch.zeros(rcap * 2, dtype=torch.int64).pin_memory().numpy()
    s3 = torch.zeros(rcap * 2, dtype=torch.int64).pin_memory().numpy()
    ctx.set_option("time_kernels", 1)
    best, best_us = None, None
    for it in range(4):
        ctx.stat("sketch_reset")
        t0 = time.perf_counter()
        rc = lib.gd_sketch_reads_batch(ctx.h, nr, gd._ptr(r_off), gd._ptr(r_len), gd._ptr(h_reads.numpy()), 11, 21, b"10", W, C.c_float(0.1), 800,
                                       gd._ptr(s2_counts), gd._ptr(s2_off), gd._ptr(s2), rcap, gd._ptr(s3_off), gd._ptr(s3_ret), gd._ptr(s3), rcap)
        dt = time.perf_counter() - t0
        ctx._check(rc, "gd_sketch_reads_batch")
        us = ctx.stat("sketch_us")
        if it and (best is None or dt < best):
            best, best_us = dt, us
    ctx.set_option("time_kernels", 0)
    out["reads"] = {"workload": "%d reads x %d bp: mm_sketch2 + mm_sketch3 for both shifts, cap 800" % (nr, rl),
                    "gbases_per_s_kernel": nr * rl / (best_us * 1e-6) / 1e9 if best_us else None,
                    "e2e": {"value": nr * rl / best / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": int(nr * rl + nr * 12),
                            "d2h_bytes_per_step": int((int(s3_off[nr * W]) + int(s2_off[nr])) * 16 + nr * W * 8 + (nr * W + nr + 2) * 8),
                            "call": "gd_sketch_reads_batch: pinned host ASCII -> the lists of every read and shift on the host"}}
    return out


def cpu_baseline_sample(args, P, ez_gpu=None, coff_gpu=None, cig_gpu=None):
    """The unmodified reference on all host threads: ~12 s of passes over the first n pairs of the step (the baseline
    figure), and -- when the GPU's results are handed in -- ONE pass over ALL pairs of the step whose ksw_extz_t fields and
    CIGARs are compared with the GPU's, pair by pair (`parity` in the JSON line)."""
    from oraclelib import Ref, cpu_has_avx512, EXTZ_DTYPE, EXTZ_FIELDS
    import gdiet_b200  # noqa: F401
    from gdiet_b200 import synth
    variant = "avx" if cpu_has_avx512() else "scalar"
    R = Ref(variant)
    cores = host_threads()
    sc = synth.SCORING["sr"]
    mat = synth.score_matrix(sc["a"], sc["b"])
    pre = band_prefix(QLEN, TLEN, BAND)
    stride = QLEN + TLEN
    N = len(P["qlen"])
    buf = (np.zeros(N, EXTZ_DTYPE), np.zeros(N * stride, np.uint32))

    def run(n):
        t0 = time.perf_counter()
        R.ksw_extd2_batch(P["qlen"][:n], P["qoff"][:n], P["qbuf"], P["tlen"][:n], P["toff"][:n], P["tbuf"], mat, sc["q"], sc["e"], sc["q2"],
                          sc["e2"], BAND, sc["zdrop"], sc["end_bonus"], args.flag, cores, cigar_stride=stride, out=buf)
        return time.perf_counter() - t0

    dt = run(4096)
    n = int(min(N, max(4096, 4096 * 4.0 / max(dt, 1e-3))))
    # ~12 s of CPU work: repeat passes over the first n pairs of the step
    passes, tot = 0, 0.0
    while tot < 12.0 and passes < 64:
        tot += run(n)
        passes += 1
    cpu = {"value": passes * n * int(pre[-1]) / tot / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
           "sample": "%d passes over the first %d pairs of the step, ksw_extd2_%s via oracle/_ref on %d threads, %.1f s" % (
               passes, n, "avx512" if variant == "avx" else "sse", cores, tot)}
    parity = None
    if ez_gpu is not None:
        if n < N:
            run(N)  # one pass over every pair of the step
        ez_ref, cig_ref = buf
        bad = np.zeros(N, bool)
        for f in EXTZ_FIELDS:
            bad |= ez_ref[f][:N] != ez_gpu[f][:N]
        ncg = np.clip(ez_ref["n_cigar"][:N], 0, stride).astype(np.int64)
        flat = cig_ref.reshape(N, stride)[np.arange(stride)[None, :] < ncg[:, None]]  # row-major: the pairs' CIGARs back to back
        total = int(coff_gpu[N])
        cig_same = len(flat) == total and bool(np.array_equal(flat, cig_gpu[:total]))
        if not cig_same and len(flat) == total:  # which pairs
            d = np.nonzero(flat != cig_gpu[:total])[0]
            bad[np.unique(np.searchsorted(coff_gpu[1:N + 1], d, side="right"))] = True
        parity = {"checked_pairs": int(N), "mismatching_pairs": int(bad.sum()) if (cig_same or len(flat) == total) else -1,
                  "fields": "max zdropped max_q max_t mqe mqe_t mte mte_q score n_cigar reach_end + every CIGAR entry",
                  "cigar_entries": total, "against": "ksw_extd2_%s of the unmodified reference (oracle/_ref) on the same pairs, flag %#x" % (
                      "avx512" if variant == "avx" else "sse", args.flag)}
    return cpu, parity


def zdrop_subset(ctx, args):
    """a 30 %-edit subset of the config-2 shape so that Z-drop fires; GPU (host-buffer call) against the
    unmodified reference on every pair."""
    import gdiet_b200 as gd
    from gdiet_b200 import synth
    from oraclelib import Ref, cpu_has_avx512, EXTZ_FIELDS
    n = 100_000
    P = synth.ksw_pairs_fast(n, QLEN, TLEN, 0.30, seed=4)
    sc = synth.SCORING["sr"]
    mat = synth.score_matrix(sc["a"], sc["b"])
    prm = gd.KswParams(mat, sc["q"], sc["e"], sc["q2"], sc["e2"], sc["zdrop"], sc["end_bonus"], args.flag)
    ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], prm, w_all=BAND)
    t0 = time.perf_counter()
    ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], prm, w_all=BAND)
    dt = time.perf_counter() - t0
    pre = band_prefix(QLEN, TLEN, BAND)
    cells = int(pre[np.clip(ez["rows_done"], 0, len(pre) - 1)].sum())
    out = {"pairs": n, "edits": 0.30, "flag": args.flag, "zdropped_frac": float((ez["zdropped"] != 0).mean()), "cells": cells,
           "gcups_e2e": cells / dt / 1e9}
    if not args.no_cpu:
        variant = "avx" if cpu_has_avx512() else "scalar"
        R = Ref(variant)
        stride = QLEN + TLEN
        ez_r, cig_r = R.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], mat, sc["q"], sc["e"], sc["q2"],
                                        sc["e2"], BAND, sc["zdrop"], sc["end_bonus"], args.flag, host_threads(), cigar_stride=stride)
        bad = np.zeros(n, bool)
        for f in EXTZ_FIELDS:
            bad |= ez_r[f] != ez[f]
        ncg = np.clip(ez_r["n_cigar"], 0, stride).astype(np.int64)
        flat = cig_r.reshape(n, stride)[np.arange(stride)[None, :] < ncg[:, None]]
        out["parity"] = {"checked_pairs": n, "mismatching_pairs": int(bad.sum()),
                         "cigars_identical": len(flat) == int(coff[n]) and bool(np.array_equal(flat, cig[:int(coff[n])]))}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--flag", type=lambda s: int(s, 0), default=0x00)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sketch", action="store_true")
    ap.add_argument("--no-map-strong", action="store_true", help="skip extra.map_strong (10 M reads vs the 3.1 Gbp index, strong-scaled over the ranks)")
    ap.add_argument("--map-reads", type=int, default=10_000_000)
    ap.add_argument("--map-genome-bp", type=int, default=3_100_000_000)
    ap.add_argument("--group", type=int, default=0, help="lanes per pair (0 = auto)")
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
