#!/usr/bin/env python
"""bench.py -- genome positions profiled / second (pileup + SNV + linkage) on B200, CPU reference beside.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

Workload (BASELINE.json configs[2] / north_star): synthetic metagenome, `--scaffolds` x `--L` bp at `--cov` x coverage,
1 % SNV density, min_cov 5, min_freq 0.05, min_snp 20, window_length 10000, --skip_mm_profiling (M = 1) unless --mm.
Defaults: 100 x 1 Mb x 100x = the full 100 Mb configuration (1e10 aligned bases); it is generated on the device
(instrain_b200/synth.py) because it cannot be produced on, or shipped from, the host in bench time.
--layout cols (default): the data set is resident as COLUMN WORDS (include/instrain_b200.h, isb_cols_batch: the one-hot
nibble words of the reads regrouped per 8-position column, 4 bits per aligned base + a 4-byte pair id per word, ~11 GB);
a "step" = isb_profile_cols (K1c streaming pileup with the SNV call fused into its epilogue -> K3 linkage).
--layout reads: READ-MAJOR aligned segments (~5.6 GB); step = isb_profile_reads (K1r transposing pileup -> K2 -> K3).
--layout events: the same fragments as position-major event columns (10 B per event, 100 GB), step = isb_profile_batch
(K1 -> K2 -> K3).  Either way the inputs are far larger than the 126 MB L2, so no L2 flush is needed between steps.
With --layout reads a short secondary run of the event layout on --also-events scaffolds is reported beside.

Multi-GPU: scaffolds are independent, so every rank profiles its own 100-scaffold shard (weak scaling, no data-path
collective) and the final SNV / linkage tables are gathered to rank 0 over NCCL inside the timed region; the gather of
step i runs on a side stream underneath the kernels of step i + 1 (two alternating sets of row buffers).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genome positions profiled/sec (pileup+SNV+LD)"
UNIT = "positions/s"
SEED = 20260103


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaffolds", type=int, default=100)
    ap.add_argument("--L", type=int, default=1000000)
    ap.add_argument("--cov", type=int, default=100)
    ap.add_argument("--dens", type=float, default=0.01)
    ap.add_argument("--mm", action="store_true", help="keep per-pair mm levels (M ~ 12-15) instead of M = 1")
    ap.add_argument("--e2e-scaffolds", type=int, default=4, help="scaffolds in the bounded host-buffer (e2e) slice")
    ap.add_argument("--cpu-scaffolds", type=int, default=1, help="scaffolds in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="cols", choices=["cols", "reads", "events"],
                    help="resident input layout: column words (default), read-major aligned segments or position-major event columns")
    ap.add_argument("--pipeline", action="store_true",
                    help="ISB_PIPELINE: cut the batch into chunks and run K3 of chunk c underneath the pileup of the later chunks")
    ap.add_argument("--keep-counts", action="store_true",
                    help="ask for the full counts / nmask arrays (column words at M = 1: disables the fused pileup + SNV kernel)")
    ap.add_argument("--seg-words", type=int, default=None, help="words per segment block of the generated read-major batch (21 or 22)")
    ap.add_argument("--also-events", type=int, default=10,
                    help="with --layout reads: also time the position-major path on this many scaffolds (0 = skip)")
    return ap.parse_args()


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_pass(host_batch, lut, dflt, threads):
    """One pass of the reference algorithm (oracle port: oracle/oracle.c) over a host batch.  Chunks of 4 splits are
    farmed to `threads` OpenMP threads inside the C library, the way the reference farms splits to worker processes
    (inStrain/profile/profile_controller.py:243-271).  Returns (#snv rows, #ld rows)."""
    from oracle import restate
    return restate.profile_mt(host_batch, host_batch["ref_codes"], lut, dflt, host_batch["splits"], n_threads=threads)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    if args.impl == "reference" and rank != 0:
        return 0
    if world > 1 and args.impl != "reference":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    from instrain_b200 import _cabi, synth
    from instrain_b200.null_model import load_lut
    lut, dflt = load_lut()
    M_label = "per-pair mm levels" if args.mm else "M=1 (--skip_mm_profiling)"
    workload = "synthetic metagenome %d x %d bp, %dx coverage, %.3g SNV density, %s, full profile (K1+K2+K3)" % (
        args.scaffolds, args.L, args.cov, args.dens, M_label)
    config = {"workload": workload, "scaffolds_per_gpu": args.scaffolds, "scaffold_len": args.L, "coverage": args.cov,
              "snv_density": args.dens, "min_cov": 5, "min_freq": 0.05, "min_snp": 20, "window_length": 10000,
              "sharding": "scaffolds per rank (weak), NCCL gather of SNV/linkage rows to rank 0" if world > 1 else "single GPU",
              "layout": {"cols": "column words (4-bit one-hot code per aligned base, regrouped per 8-position column; pair id per word for linkage)",
                         "reads": "read-major aligned segments (4-bit code per aligned base)",
                         "events": "position-major event columns (10 B per event)"}[args.layout],
              "l2": "inputs (%.1f GB) larger than L2; no flush needed" % (
                  args.scaffolds * args.L * args.cov * {"cols": 1.07, "reads": 0.56, "events": 10}[args.layout] / 1e9)}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        n_sc = max(1, min(args.scaffolds, 2))
        d = synth.generate(local_rank, args.L, n_sc, args.cov, args.dens, SEED, skip_mm=not args.mm)
        hb = synth.to_host_batch(d, 0, n_sc)
        del d
        torch.cuda.empty_cache()
        cores = host_cores()
        for _ in range(args.warmup):
            cpu_oracle_pass(hb, lut, dflt, cores)
        t0 = time.time()
        for _ in range(args.steps):
            rows = cpu_oracle_pass(hb, lut, dflt, cores)
        dt = (time.time() - t0) / max(1, args.steps)
        val = n_sc * args.L / dt
        sample = "%d scaffold(s) x %d bp at %dx (%d events) of the same workload per step" % (n_sc, args.L, args.cov, len(hb["ref_pos"]))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 counts / f64 statistics", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference algorithm = oracle/oracle.c (C restatement pinned on the reference's goldens); the "
                    "reference itself is pure Python + pysam and cannot run on this box", "rows": list(rows)}))
        return 0

    # ------------------------------------------------------------------------------------------------ B200 arm
    from instrain_b200.engine import Engine
    use_cols = args.layout == "cols"
    use_reads = args.layout == "reads"
    t_gen = time.time()
    d = synth.generate(local_rank, args.L, args.scaffolds, args.cov, args.dens, SEED + rank, skip_mm=not args.mm,
                       events=args.layout == "events", reads=use_reads or use_cols, seg_words=args.seg_words)
    torch.cuda.synchronize()
    n, npairs, Ltot = int(d["n_events"]), d["pair_mm"].numel(), args.L * args.scaffolds
    M = int(d["pair_mm"].max().item()) + 1 if npairs else 1
    eng = Engine(local_rank, lut, dflt)
    cd = None
    if use_cols:                                     # lay the generated reads out as column words, drop the read-major copy
        cd = synth.reads_to_cols_device(eng, d)
        cd["n_real_words"] = int((cd["ids"] >= 0).sum().item())
        cd["n_segs"] = int(d["reads"]["n_segs"])
        del d["reads"]
        eng.close()
        eng = Engine(local_rank, lut, dflt)          # releases the conversion's staging buffers
        torch.cuda.empty_cache()
    t_gen = time.time() - t_gen
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    lib, ctx, p = eng.lib, eng.ctx, _cabi.ptr

    counts = torch.empty((Ltot, M, 4), dtype=torch.int32, device=dev)
    nmask = torch.empty(Ltot, dtype=torch.int64, device=dev)
    covT = torch.empty((Ltot, M), dtype=torch.int32, device=dev)
    clonT = torch.empty((Ltot, M), dtype=torch.float32, device=dev)
    flags = torch.empty(Ltot, dtype=torch.uint8, device=dev)
    snv_cap, ld_cap = max(1 << 16, Ltot // 16), max(1 << 18, Ltot // 2)
    res = None
    # world > 1: two sets of row buffers, alternated, so that the NCCL gather of step i (side stream) overlaps the kernels
    # of step i + 1; a set is reused only after its gather has completed (event).
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    sets, gather_done = [], [None, None]

    def alloc_rows():
        nonlocal snv, ld, res
        if side is not None:
            side.synchronize()                           # no gather may still read the buffers being replaced
        sets.clear()
        lean = (use_cols or use_reads) and not args.keep_counts   # counts / nmask not requested: the SNV call runs in the pileup kernel's epilogue at M = 1
        for _ in range(2 if world > 1 else 1):
            s_ = torch.empty(snv_cap * 32, dtype=torch.uint8, device=dev)
            l_ = torch.empty(ld_cap * 48, dtype=torch.uint8, device=dev)
            r_ = _cabi.IsbResult(None if lean else p(counts), None if lean else p(nmask), p(covT), p(clonT), p(flags), p(s_), snv_cap,
                                 p(l_), ld_cap, 0, 0, 0, 0)
            sets.append((s_, l_, r_))
        snv, ld, res = sets[0]
        gather_done[0] = gather_done[1] = None

    snv = ld = None
    alloc_rows()

    def reads_struct(rd, n_pairs, pair_mm, L_, ref, splits, M_):
        return _cabi.IsbReadsBatch(int(rd["n_segs"]), p(rd["seg_start"]), p(rd["seg_len"]), p(rd["seg_pair"]),
                                   p(rd["seg_word"]), int(rd["n_words"]), p(rd["words"]), int(rd["max_seg_len"]), 0,
                                   len(rd["nev_pos"]), p(rd["nev_pos"]), p(rd["nev_pair"]), n_pairs, p(pair_mm), 0, L_,
                                   p(ref), len(splits), p(splits), M_, 0)

    if use_cols:
        batch = _cabi.IsbColsBatch(cd["n_groups"], p(cd["grp_off"]), cd["n_chunks"], p(cd["words"]), p(cd["ids"]), 0, None, None,
                                   npairs, p(d["pair_mm"]), 0, Ltot, p(d["ref_codes"]), len(d["splits"]), p(d["splits"]), M, 0)
        entry = lib.isb_profile_cols
    elif use_reads:
        batch = reads_struct(d["reads"], npairs, d["pair_mm"], Ltot, d["ref_codes"], d["splits"], M)
        entry = lib.isb_profile_reads
    else:
        batch = _cabi.IsbBatch(n, p(d["ref_pos"]), p(d["base"]), p(d["qual"]), p(d["read_id"]), npairs, p(d["pair_mm"]), 0,
                               Ltot, p(d["ref_codes"]), d["splits"].shape[0], p(d["splits"]), M)
        entry = lib.isb_profile_batch
    prm = _cabi.IsbParams(5, 20, 30, _cabi.ISB_PIPELINE if args.pipeline else 0, 0.05)

    def gather_tables(cur):
        """NCCL gather of the final SNV / linkage rows of set `cur` to rank 0 (the only collective of the job), enqueued on
        the side stream: it runs underneath the next step's kernels."""
        if world == 1:
            return
        s_, l_, r_ = sets[cur]
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            mine = torch.tensor([r_.n_snv, r_.n_ld], dtype=torch.int64, device=dev)
            allc = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allc, mine)
            mx = torch.stack(allc).max(0).values.tolist()
            for buf, rowb, m in ((s_, 32, mx[0]), (l_, 48, mx[1])):
                view = buf[:m * rowb]
                dst = [torch.empty_like(view) for _ in range(world)] if rank == 0 else None
                dist.gather(view, dst, dst=0)
            ev = torch.cuda.Event()
            ev.record(side)
            gather_done[cur] = ev

    step_no = 0

    def step():
        nonlocal snv_cap, ld_cap, step_no, res
        cur = step_no % len(sets)
        if gather_done[cur] is not None:
            stream.wait_event(gather_done[cur])          # this set's rows were gathered two steps ago: wait for that only
        r_ = sets[cur][2]
        rc = entry(ctx, C.byref(batch), C.byref(prm), C.byref(r_))
        if rc == _cabi.ISB_ERR_CAPACITY:
            snv_cap, ld_cap = max(snv_cap, int(r_.n_snv) + 1024), max(ld_cap, int(r_.n_ld) + 1024)
            alloc_rows()
            cur, r_ = 0, sets[0][2]
            rc = entry(ctx, C.byref(batch), C.byref(prm), C.byref(r_))
        if rc != 0:
            raise RuntimeError(lib.isb_last_error(ctx).decode())
        res = r_
        gather_tables(cur)
        step_no += 1

    for _ in range(max(3, args.warmup)):
        step()
    eng.enable_timing(True)
    eng.stage_times()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    if side is not None:
        stream.wait_stream(side)                         # the last gathers belong to the timed region
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w1 = time.time()
    ms_total = e0.elapsed_time(e1)
    stage_ms, stage_calls = eng.stage_times()
    eng.enable_timing(False)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(w0, w1) if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * Ltot / (ms_step / 1e3)

    # -------------------------------------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = measured_peak()
    stage_per_step = {"k1_pileup": stage_ms[0] / args.steps, "k2_snv": stage_ms[1] / args.steps,
                      "k3_linkage": stage_ms[2] / args.steps}

    def traffic_of(key, units):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            e = tj[key]
            return (e["dram_bytes_per_unit"] if "dram_bytes_per_unit" in e else e["dram_bytes_per_event"]) * units
        except Exception:
            return None

    def events_roofline(k1_ms_, n_, M_, Ltot_):
        ev_bytes = 10 if M_ > 1 else 6                # at M = 1 K1 does not need (and does not read) read_id
        alg = n_ * ev_bytes + 16 * M_ * Ltot_ + 8 * Ltot_
        return {"kernel": "k1_pileup_tiles_tma<M=1>" if M_ == 1 else "k1_pileup_tiles_tma<M>1>", "bound": "hbm",
                "achieved": alg / k1_ms_ / 1e6, "peak": peak, "unit": "GB/s", "frac": alg / k1_ms_ / 1e6 / peak,
                "traffic": traffic_of("M1" if M_ == 1 else "Mgt1", n_), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg,
                "bytes_def": "%d B/event (ref_pos i32 + base u8 + qual u8%s) + 16*M B/position counts + 8 B/position nmask"
                             % (ev_bytes, " + read_id i32" if M_ > 1 else "; read_id not needed at M=1"),
                "achieved_survey_def": (n_ * 10 + 16 * M_ * Ltot_) / k1_ms_ / 1e6, "launch_ms": k1_ms_}

    k1_ms = stage_ms[0] / max(1, stage_calls[0])
    if use_cols:
        fused = M == 1 and not args.keep_counts and os.environ.get("ISB_K1C_FUSE", "1") != "0"   # the library's own A/B switch
        # algorithmic bytes of K1c: the real (non-padding) nibble words (+ their pair ids when M > 1) + the group offsets in;
        # fused M = 1: ref in, covT + clonT + site_flags out (+ 32 B per SNV row, 16 B of counts per linkage site);
        # otherwise counts + nmask out
        in_bytes = cd["n_real_words"] * (4 if M == 1 else 8) + (cd["n_groups"] + 1) * 8
        out_bytes = (Ltot * (1 + 4 + 4 + 1) + int(res.n_snv) * 32 + int(res.n_sites) * 16) if fused else (16 * M * Ltot + 8 * Ltot)
        alg_bytes = in_bytes + out_bytes
        key = "K1c_fused_M1" if fused else ("K1c_M1" if M == 1 else "K1c_Mgt1")
        roofline = {"kernel": "k1c_pileup_m1<fused SNV call>" if fused else ("k1c_pileup_m1" if M == 1 else "k1c_pileup_mm"),
                    "bound": "hbm", "achieved": alg_bytes / k1_ms / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": alg_bytes / k1_ms / 1e6 / peak, "traffic": traffic_of(key, Ltot), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "bytes_def": ("4 bits per aligned base (one 32-bit word per read and 8-position column, padding words not counted)%s"
                                  " + 8 B per 64 positions of offsets in; %s out") % (
                                      " + 4 B pair id per word" if M > 1 else "",
                                      "ref 1 B in, covT 4 + clonT 4 + site_flags 1 B per position, 32 B per SNV row, 16 B counts per linkage site"
                                      if fused else "16*M B/position counts + 8 B/position nmask"),
                    "padding_words_frac": cd["n_chunks"] * 64 / max(1, cd["n_real_words"]) - 1,
                    "launch_ms": k1_ms,
                    "note": ("the stage is ONE kernel: pileup counts + the per-site SNV call (K2) in its epilogue" if fused else
                             "pileup counts only; K2 runs as its own kernel")}
    elif use_reads:
        rd = d["reads"]
        # algorithmic bytes of K1r: the nibble stream + the segment table (start i32, len u16, word offset i64, + pair
        # id i32 when M > 1) in, counts (+ nmask) out
        alg_bytes = int(rd["n_words"]) * 4 + int(rd["n_segs"]) * (14 + (4 if M > 1 else 0)) + 16 * M * Ltot + 8 * Ltot
        roofline = {"kernel": "k1r_pileup<M=1>" if M == 1 else "k1r_pileup<M>1>", "bound": "hbm",
                    "achieved": alg_bytes / k1_ms / 1e6, "peak": peak, "unit": "GB/s", "frac": alg_bytes / k1_ms / 1e6 / peak,
                    "traffic": traffic_of("K1r_M1" if M == 1 else "K1r_Mgt1", Ltot), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "bytes_def": "4 bits per aligned base (nibble stream incl. separators) + 14-18 B per segment in, 16*M B/position "
                                 "counts + 8 B/position nmask out; the same pileup from event columns would read %d B/event"
                                 % (10 if M > 1 else 6),
                    "equivalent_event_column_rate": (n * (10 if M > 1 else 6) + 16 * M * Ltot + 8 * Ltot) / k1_ms / 1e6,
                    "launch_ms": k1_ms,
                    "note": "K1r is bound by issue slots / shared-memory bandwidth, not HBM: the read-major layout removed "
                            "~90 % of the pileup's DRAM bytes (see position_major_path for the HBM-bound event-column kernel)"}
    else:
        roofline = events_roofline(k1_ms, n, M, Ltot)
    roofline["stage_ms_per_step"] = stage_per_step

    # position-major path on a subset, for comparison (the HBM-bound K1 kernel on 10 B/event columns)
    pos_major = None
    if (use_reads or use_cols) and rank == 0 and args.also_events > 0:
        n_sc = min(args.also_events, args.scaffolds)
        de = synth.generate(local_rank, args.L, n_sc, args.cov, args.dens, SEED + rank, skip_mm=not args.mm)
        Le = n_sc * args.L
        ne, npe = de["ref_pos"].numel(), de["pair_mm"].numel()
        be = _cabi.IsbBatch(ne, p(de["ref_pos"]), p(de["base"]), p(de["qual"]), p(de["read_id"]), npe, p(de["pair_mm"]), 0,
                            Le, p(de["ref_codes"]), de["splits"].shape[0], p(de["splits"]), M)
        re_ = _cabi.IsbResult(p(counts), p(nmask), p(covT), p(clonT), p(flags), p(snv), snv_cap, p(ld), ld_cap, 0, 0, 0, 0)
        for it in range(3 + 3):
            if it == 3:
                eng.enable_timing(True)
                eng.stage_times()
                torch.cuda.synchronize()
                e0.record(stream)
            if lib.isb_profile_batch(ctx, C.byref(be), C.byref(prm), C.byref(re_)) != 0:
                raise RuntimeError(lib.isb_last_error(ctx).decode())
        e1.record(stream)
        torch.cuda.synchronize()
        sm, sc = eng.stage_times()
        eng.enable_timing(False)
        ms_e = e0.elapsed_time(e1) / 3
        pos_major = {"value": Le / (ms_e / 1e3), "unit": UNIT, "ms_per_step": ms_e, "scaffolds": n_sc,
                     "roofline": events_roofline(sm[0] / max(1, sc[0]), ne, M, Le),
                     "stage_ms_per_step": {"k1_pileup": sm[0] / 3, "k2_snv": sm[1] / 3, "k3_linkage": sm[2] / 3},
                     "rows_equal": None}
        del de, be
        torch.cuda.empty_cache()

    # -------------------------------------------------------------------------------- e2e: host buffers through the C-ABI
    # Public call a user makes: pinned HOST buffers in (the host packer's packed transfer format, ~1 B/event),
    # isb_profile_batch_packed (H2D + K0 expand + K1 + K2 + K3 + D2H of every result table), pinned HOST tables out.
    # Also measured with the 10 B/event columnar host buffers (isb_profile_batch) for comparison.
    e2e = None
    if rank == 0:
        from instrain_b200.packed import encode_packed
        n_sc = max(1, min(args.e2e_scaffolds, args.scaffolds))
        # the first n_sc scaffolds of the data set, regenerated in both layouts (the generator is deterministic per seed)
        ds = synth.generate(local_rank, args.L, n_sc, args.cov, args.dens, SEED + rank, skip_mm=not args.mm, events=True,
                            reads=True, seg_words=args.seg_words)
        hb = synth.to_host_batch(ds, 0, n_sc)
        hr = synth.reads_to_host(ds, 0, n_sc)["reads"]
        del ds
        torch.cuda.empty_cache()
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        Ls = n_sc * args.L
        Ms = int(hb["pair_mm"].max()) + 1 if len(hb["pair_mm"]) else 1
        pk = encode_packed(hb, 0, Ls, 30)
        h = {k: pin(hb[k]) for k in ("ref_pos", "base", "qual", "read_id", "pair_mm", "ref_codes", "splits")}
        hp = {k: pin(pk[k]) for k in ("pos_off", "id_base", "bqd", "esc_evt", "esc_id")}
        o = dict(covT=torch.empty((Ls, Ms), dtype=torch.int32).pin_memory(),
                 clonT=torch.empty((Ls, Ms), dtype=torch.float32).pin_memory(),
                 flags=torch.empty(Ls, dtype=torch.uint8).pin_memory(),
                 snv=torch.empty(max(1 << 16, (Ls // 16) * (1 if Ms == 1 else 16)) * 32, dtype=torch.uint8).pin_memory(),
                 ld=torch.empty(max(1 << 18, (Ls // 2) * (1 if Ms == 1 else 8)) * 48, dtype=torch.uint8).pin_memory())
        hbatch = _cabi.IsbBatch(len(hb["ref_pos"]), p(h["ref_pos"]), p(h["base"]), p(h["qual"]), p(h["read_id"]),
                                len(hb["pair_mm"]), p(h["pair_mm"]), 0, Ls, p(h["ref_codes"]), len(hb["splits"]),
                                p(h["splits"]), Ms)
        pbatch = _cabi.IsbPackedBatch(pk["n_events"], p(hp["pos_off"]), p(hp["id_base"]), p(hp["bqd"]), len(pk["esc_evt"]),
                                      p(hp["esc_evt"]), p(hp["esc_id"]), len(hb["pair_mm"]), p(h["pair_mm"]), 0, Ls,
                                      p(h["ref_codes"]), len(hb["splits"]), p(h["splits"]), Ms, 30)
        hres = _cabi.IsbResult(None, None, p(o["covT"]), p(o["clonT"]), p(o["flags"]), p(o["snv"]),
                               o["snv"].numel() // 32, p(o["ld"]), o["ld"].numel() // 48, 0, 0, 0, 0)
        hrp = {k: pin(hr[k].view(np.int16) if k == "seg_len" else (hr[k].view(np.int32) if k == "words" else hr[k]))
               for k in ("seg_start", "seg_len", "seg_pair", "seg_word", "words")}
        hrp.update(n_segs=hr["n_segs"], n_words=hr["n_words"], max_seg_len=hr["max_seg_len"], nev_pos=hr["nev_pos"],
                   nev_pair=hr["nev_pair"])
        rbatch = reads_struct(hrp, len(hb["pair_mm"]), h["pair_mm"], Ls, h["ref_codes"], h["splits"], Ms)
        from instrain_b200.reads import compact_reads
        hc = compact_reads(hr)                          # compact transfer format: 3 bits per aligned base, no word offsets
        hcp = {"base2": pin(hc["base2"].view(np.int16)), "pass": pin(hc["pass"])}
        cbatch = _cabi.IsbReadsCompact(int(hr["n_segs"]), p(hrp["seg_start"]), p(hrp["seg_len"]), p(hrp["seg_pair"]),
                                       int(hc["n_units"]), p(hcp["base2"]), p(hcp["pass"]), int(hr["max_seg_len"]), 0,
                                       len(hr["nev_pos"]), p(hr["nev_pos"]), p(hr["nev_pair"]), len(hb["pair_mm"]),
                                       p(h["pair_mm"]), 0, Ls, p(h["ref_codes"]), len(hb["splits"]), p(h["splits"]), Ms, 0)

        from instrain_b200.reads import delta_reads
        hd = delta_reads(hr, hb["ref_codes"])           # reference-delta transfer format: event bits + mismatch entries
        hdp = {"pass": pin(hd["pass"]), "mis_word": pin(hd["mis_word"].view(np.int32)), "mis_code": pin(hd["mis_code"])}
        dbatch = _cabi.IsbReadsDelta(int(hr["n_segs"]), p(hrp["seg_start"]), p(hrp["seg_len"]), p(hrp["seg_pair"]),
                                     int(hd["n_units"]), p(hdp["pass"]), len(hd["mis_word"]), p(hdp["mis_word"]), p(hdp["mis_code"]),
                                     int(hr["max_seg_len"]), 0, len(hr["nev_pos"]), p(hr["nev_pos"]), p(hr["nev_pair"]),
                                     len(hb["pair_mm"]), p(h["pair_mm"]), 0, Ls, p(h["ref_codes"]), len(hb["splits"]),
                                     p(h["splits"]), Ms, 0)

        def time_call(fn, b):
            ts = []
            for it in range(2 + 3):
                torch.cuda.synchronize()
                a = time.time()
                rc = fn(ctx, C.byref(b), C.byref(prm), C.byref(hres))
                torch.cuda.synchronize()
                if rc != 0:
                    raise RuntimeError(lib.isb_last_error(ctx).decode())
                if it >= 2:
                    ts.append(time.time() - a)
            return float(np.median(ts))

        dt_col = time_call(lib.isb_profile_batch, hbatch)
        dt_pk = time_call(lib.isb_profile_batch_packed, pbatch)
        dt_rd = time_call(lib.isb_profile_reads, rbatch)
        dt_rc = time_call(lib.isb_profile_reads_compact, cbatch)
        dt_rdl = time_call(lib.isb_profile_reads_delta, dbatch)
        via_reads = use_reads or use_cols             # host buffers cross PCIe in a read-major transfer format either way
        main_fn, main_b, dt = (lib.isb_profile_reads_delta, dbatch, dt_rdl) if via_reads else (lib.isb_profile_batch_packed, pbatch, dt_pk)

        # Two contexts on two host threads (each call is still host buffers -> C-ABI -> host tables): the H2D copy of one
        # call overlaps the kernels and the D2H copy of the other, which is how a host pipeline feeds the GPU.
        dt_pipe = None
        try:
            eng2 = Engine(local_rank, lut, dflt)
            o2 = {k: torch.empty_like(v).pin_memory() for k, v in o.items()}
            hres2 = _cabi.IsbResult(None, None, p(o2["covT"]), p(o2["clonT"]), p(o2["flags"]), p(o2["snv"]),
                                    o2["snv"].numel() // 32, p(o2["ld"]), o2["ld"].numel() // 48, 0, 0, 0, 0)
            n_it = 4
            errs = []

            def worker(c, r):
                for _ in range(n_it):
                    if main_fn(c, C.byref(main_b), C.byref(prm), C.byref(r)) != 0:
                        errs.append(lib.isb_last_error(c).decode())

            for rep in range(2):                         # first repetition warms the second context's scratch buffers
                torch.cuda.synchronize()
                t_a = time.time()
                ths = [threading.Thread(target=worker, args=(ctx, hres)), threading.Thread(target=worker, args=(eng2.ctx, hres2))]
                [t.start() for t in ths]
                [t.join() for t in ths]
                torch.cuda.synchronize()
                dt_pipe = (time.time() - t_a) / (2 * n_it)
            if errs:
                raise RuntimeError(errs[0])
            eng2.close()
        except Exception as ex:                          # noqa: BLE001 - the pipelined figure is optional
            dt_pipe = None
            print("e2e pipelined leg skipped: %r" % (ex,), file=sys.stderr)
        common = len(hb["pair_mm"]) + Ls + hb["splits"].nbytes
        h2d_pk = pk["n_events"] + (Ls + 1) * 8 + Ls * 4 + len(pk["esc_evt"]) * 12 + common
        h2d_rd = hr["n_words"] * 4 + hr["n_segs"] * (4 + 2 + 4 + 8) + common
        h2d_rc = hc["n_units"] * 3 + hr["n_segs"] * (4 + 2 + 4) + common
        h2d_rdl = hd["n_units"] + len(hd["mis_word"]) * 5 + hr["n_segs"] * (4 + 2 + 4) + common
        d2h = Ls * Ms * 8 + Ls + int(hres.n_snv) * 32 + int(hres.n_ld) * 48
        best = min(dt, dt_pipe) if dt_pipe else dt
        e2e = {"value": Ls / best, "unit": UNIT, "h2d_bytes_per_step": int(h2d_rdl if via_reads else h2d_pk),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": best * 1e3,
               "api": ("isb_profile_reads_delta (read-major aligned segments in the reference-delta transfer format: event bits "
                       "+ one entry per base that differs from the reference, K0d rebuilds the stream on the device)" if via_reads else
                       "isb_profile_batch_packed (packed transfer format, K0 expands on the device)"),
               "single_context": {"value": Ls / dt, "ms_per_step": dt * 1e3},
               "two_contexts_pipelined": ({"value": Ls / dt_pipe, "ms_per_step": dt_pipe * 1e3} if dt_pipe else None),
               "slice": "%d of the %d scaffolds per step, pinned host buffers -> C-ABI -> pinned host result tables" % (n_sc, args.scaffolds),
               "other_host_formats": {
                   "read_major_delta": {"value": Ls / dt_rdl, "ms_per_step": dt_rdl * 1e3, "h2d_bytes_per_step": int(h2d_rdl)},
                   "read_major_compact": {"value": Ls / dt_rc, "ms_per_step": dt_rc * 1e3, "h2d_bytes_per_step": int(h2d_rc)},
                   "read_major_segments": {"value": Ls / dt_rd, "ms_per_step": dt_rd * 1e3, "h2d_bytes_per_step": int(h2d_rd)},
                   "packed_events": {"value": Ls / dt_pk, "ms_per_step": dt_pk * 1e3, "h2d_bytes_per_step": int(h2d_pk)},
                   "columnar_events": {"value": Ls / dt_col, "ms_per_step": dt_col * 1e3,
                                       "h2d_bytes_per_step": int(len(hb["ref_pos"]) * 10 + common)}}}

    # -------------------------------------------------------------------------------- CPU baseline (oracle port) beside
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sc = max(1, min(args.cpu_scaffolds, args.scaffolds))
        ds = synth.generate(local_rank, args.L, n_sc, args.cov, args.dens, SEED + rank, skip_mm=not args.mm)
        hb = synth.to_host_batch(ds, 0, n_sc)
        del ds
        a = time.time()
        cpu_oracle_pass(hb, lut, dflt, 1)
        dt = time.time() - a
        cpu = {"value": n_sc * args.L / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d scaffold(s) x %d bp at %dx (%d events) of the same workload, one pass, 1 thread of oracle/oracle.c (orc_profile_mt)"
                         % (n_sc, args.L, args.cov, len(hb["ref_pos"]))}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 counts / f64 statistics", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "position_major_path": pos_major,
            "rows": {"n_events_per_gpu": n, "n_pairs_per_gpu": npairs, "M": M, "n_snv": int(res.n_snv), "n_ld": int(res.n_ld),
                     "n_sites": int(res.n_sites), "n_site_pairs": int(res.n_site_pairs)},
            "setup_s": round(t_gen, 1)}))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
