#!/usr/bin/env python
"""bench.py -- genome positions profiled / second (pileup + SNV + linkage) on B200, CPU reference beside.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

Workload (BASELINE.json configs[2] / north_star): synthetic metagenome, `--scaffolds` x `--L` bp at `--cov` x coverage,
`--dens` SNV density, min_cov 5, min_freq 0.05, min_snp 20, window_length 10000, --skip_mm_profiling (M = 1) unless --mm.
Defaults: 100 x 1 Mb x 100x = the full 100 Mb configuration (1e10 aligned bases).  Scaffold k is generated on the device
from seed SEED + k (instrain_b200/synth.py): the data set is the same whatever the number of GPUs.
Other BASELINE configurations: C2 `--scaffolds 1 --cov 50 --skip-linkage`, C5 `--scaffolds 1 --L 10000000 --cov 500
--dens 0.05`, the reference's default mode (mm profiling on) `--mm`.

A "step" = one pass of the whole hot path over the resident data set, FROM BAM-ORDER ALIGNED SEGMENTS (what the host
packer emits from a BAM: one 4-bit code per aligned base, stored once per read; include/instrain_b200.h,
isb_reads_batch): isb_profile_reads = K1f (pileup + SNV call + bit rows of the linkage sites, one kernel) -> linkage back
end.  Nothing is pre-transposed or pre-filtered outside the timed region.  Inputs (5.6 GB) are far larger than the 126 MB
L2: no flush between steps.  `--layout cols | events` time the older resident layouts instead (column words need a
conversion that is NOT part of their step; reported for comparison under `other_layouts` on a subset by default).

Multi-GPU: scaffolds are independent (the reference farms splits to processes, profile_controller.py:243-271).  The SAME
scaffolds are partitioned over the ranks (LPT on aligned bases, instrain_b200/shard.py): strong scaling, no data-path
collective; the final SNV / linkage tables are gathered to rank 0 over NCCL inside the timed region, on a side stream
underneath the next step's kernels.  `weak_scaling` (every rank its own --scaffolds) is reported beside for N > 1.
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
    ap.add_argument("--skip-linkage", action="store_true", help="pileup + SNV call only (BASELINE config 2)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: shard the same --scaffolds over the ranks (strong) or give every rank its own (weak)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaling line")
    ap.add_argument("--e2e-scaffolds", type=int, default=4, help="scaffolds in the bounded host-buffer (e2e) slice per rank")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (configurations whose one scaffold is too large for a bounded slice)")
    ap.add_argument("--cpu-scaffolds", type=int, default=1, help="scaffolds in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained leg in seconds (0 = skip)")
    ap.add_argument("--layout", default="reads", choices=["reads", "cols", "events"],
                    help="resident input layout: BAM-order aligned segments (default), column words or event columns")
    ap.add_argument("--keep-counts", action="store_true",
                    help="ask for the full counts / nmask arrays (M = 1: disables the fused pileup + SNV kernels)")
    ap.add_argument("--also-layouts", type=int, default=10,
                    help="also time the column-word and event-column layouts on this many scaffolds (0 = skip; N = 1 only)")
    ap.add_argument("--from-bam-scaffolds", type=int, default=8,
                    help="scaffolds of the from-BAM leg (a real BAM of the workload's shape is written, then profile_bam runs on it; "
                         "0 = skip; N = 1 only)")
    ap.add_argument("--from-bam-L", type=int, default=250000, help="scaffold length of the from-BAM leg")
    ap.add_argument("--seg-words", type=int, default=None, help="words per segment block of the generated read-major batch (21 or 22)")
    return ap.parse_args()


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING a timed region (B200_PROFILING.md's clocks line)."""
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

    def window(self, t0, t1):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


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


def build_dataset(synth, device, ids, args):
    """Scaffolds `ids` of the data set (scaffold k <- seed SEED + k) as ONE read-major batch in HBM: coordinates, pair ids
    and word offsets of the per-scaffold batches re-based and concatenated (the layout rules of isb_reads_batch hold: one
    leading zero word, [data words + separator] per segment, zero padding to a multiple of 4 words)."""
    import torch
    dev = torch.device("cuda", device)
    parts, n_pairs, n_words, n_ev = [], 0, 1, 0
    spw = (synth.READLEN + 14) // 8 + 1 if args.seg_words is None else int(args.seg_words)
    for j, k in enumerate(ids):
        g = synth.generate(device, args.L, 1, args.cov, args.dens, SEED + int(k), skip_mm=not args.mm, events=False, reads=True,
                           seg_words=args.seg_words)
        rd = g["reads"]
        n = int(rd["n_segs"])
        parts.append(dict(seg_start=rd["seg_start"] + j * args.L, seg_len=rd["seg_len"], seg_pair=rd["seg_pair"] + n_pairs,
                          seg_word=rd["seg_word"] + (n_words - 1), words=rd["words"][1:1 + n * spw], pair_mm=g["pair_mm"],
                          ref_codes=g["ref_codes"], splits=g["splits"] + j * args.L))
        n_pairs += g["pair_mm"].numel()
        n_words += n * spw
        n_ev += int(g["n_events"])
        del g, rd

    def cat(key, dt, shape=(0,)):
        return torch.cat([p[key] for p in parts]) if parts else torch.empty(shape, dtype=dt, device=dev)

    pad = (-n_words) % 4
    words = torch.cat([torch.zeros(1, dtype=torch.int32, device=dev)] + [p["words"] for p in parts] +
                      [torch.zeros(pad, dtype=torch.int32, device=dev)])
    reads = dict(n_segs=sum(p["seg_start"].numel() for p in parts), n_words=n_words + pad, max_seg_len=synth.READLEN, seg_words=spw,
                 seg_start=cat("seg_start", torch.int32), seg_len=cat("seg_len", torch.int16), seg_pair=cat("seg_pair", torch.int32),
                 seg_word=cat("seg_word", torch.int64), words=words,
                 nev_pos=torch.empty(0, dtype=torch.int32, device=dev), nev_pair=torch.empty(0, dtype=torch.int32, device=dev))
    d = dict(reads=reads, pair_mm=cat("pair_mm", torch.uint8), ref_codes=cat("ref_codes", torch.uint8),
             splits=cat("splits", torch.int32, (0, 2)), L=args.L, n_scaffolds=len(ids), n_events=n_ev)
    del parts[:]
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return d


def clip_dataset_to_run(d, lo, hi):
    """Split-run sharding of ONE scaffold: keep the part of a device-resident read-major data set that lies
    in [lo, hi) -- the device-side twin of instrain_b200.reads.clip_reads.  The segments that overlap the range are a
    contiguous slice of the start-sorted table; the few that stick out (coverage many at each border) are cut in place:
    start / length / first word moved, the nibbles outside the range zeroed.  The batch is then profiled with start =
    lo & ~7 and L = hi - start; coordinates, pair ids and the keys of the re-drawn outputs stay those of the whole."""
    import torch
    rd = d["reads"]
    s = rd["seg_start"].to(torch.int64)
    n = rd["seg_len"].to(torch.int64) & 0xffff
    keep = torch.nonzero((s < hi) & (s + n > lo)).flatten()
    i0, i1 = (int(keep[0]), int(keep[-1]) + 1) if keep.numel() else (0, 0)
    origin = lo & ~7
    words = rd["words"].clone()
    seg_start, seg_len = rd["seg_start"][i0:i1].clone(), rd["seg_len"][i0:i1].clone()
    seg_pair, seg_word = rd["seg_pair"][i0:i1].clone(), rd["seg_word"][i0:i1].clone()
    s, n = s[i0:i1], n[i0:i1]
    cut = torch.nonzero((s < lo) | (s + n > hi)).flatten()
    if cut.numel():                                         # the border segments, fixed on the host (a few hundred)
        cs, cn, cw = s[cut].cpu().numpy(), n[cut].cpu().numpy(), seg_word[cut].cpu().numpy()
        nw_old = ((cs & 7) + cn + 7) >> 3
        idx = np.concatenate([w0 + np.arange(k) for w0, k in zip(cw, nw_old)])
        old = words[torch.from_numpy(idx).to(words.device)].cpu().numpy().astype(np.uint32).astype(np.uint64)
        s2, e2 = np.maximum(cs, lo), np.minimum(cs + cn, hi)
        c_first, c_last = (s2 >> 3) - (cs >> 3), ((e2 - 1) >> 3) - (cs >> 3)
        new = np.zeros_like(old)
        off = np.concatenate([[0], np.cumsum(nw_old)])
        for j in range(len(cs)):
            w = old[off[j]:off[j + 1]].copy()
            w[:c_first[j]] = 0
            w[c_last[j] + 1:] = 0
            w[c_first[j]] &= ~((np.uint64(1) << np.uint64(4 * (s2[j] & 7))) - np.uint64(1))
            w[c_last[j]] &= (np.uint64(1) << np.uint64(4 * (((e2[j] - 1) & 7) + 1))) - np.uint64(1)
            new[off[j]:off[j + 1]] = w
        words[torch.from_numpy(idx).to(words.device)] = torch.from_numpy((new & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32)).to(words.device)
        dev = seg_start.device
        seg_start[cut] = torch.from_numpy(s2.astype(np.int32)).to(dev)
        seg_len[cut] = torch.from_numpy((e2 - s2).astype(np.int16)).to(dev)
        seg_word[cut] = torch.from_numpy((cw + c_first).astype(np.int64)).to(dev)
    out = dict(d)
    out["reads"] = dict(rd, n_segs=i1 - i0, seg_start=seg_start, seg_len=seg_len, seg_pair=seg_pair, seg_word=seg_word, words=words)
    sp = d["splits"]
    out["splits"] = sp[(sp[:, 0] >= lo) & (sp[:, 1] < hi)].contiguous()
    out["ref_codes"] = d["ref_codes"][origin:hi].contiguous()
    out["start"], out["L_run"] = origin, hi - origin
    out["n_events"] = int((seg_len.to(torch.int64) & 0xffff).sum().item())
    return out


class Job:
    """One rank's share of the workload resident in HBM + the result buffers; step() = one pass of the hot path."""

    def __init__(self, eng, d, args, dev, layout="reads", cols=None, two_sets=False):
        import torch
        from instrain_b200 import _cabi
        self.eng, self.d, self.args, self.dev, self.layout = eng, d, args, dev, layout
        self._cabi = _cabi
        p = _cabi.ptr
        self.Ltot = d.get("L_run", d["L"] * d["n_scaffolds"])   # a run of splits of one scaffold: its own start / length
        start_ = int(d.get("start", 0))
        self.npairs = d["pair_mm"].numel()
        self.M = int(d["pair_mm"].max().item()) + 1 if self.npairs else 1
        L_, M_ = self.Ltot, self.M
        # no raw counts asked (the reference stores none either): at M = 1 the SNV call runs inside the pileup kernel
        self.lean = layout in ("reads", "cols") and not args.keep_counts
        self.counts = None if self.lean else torch.empty((L_, M_, 4), dtype=torch.int32, device=dev)
        self.nmask = None if self.lean else torch.empty(L_, dtype=torch.int64, device=dev)
        self.covT = torch.empty((L_, M_), dtype=torch.int32, device=dev)
        self.clonT = torch.empty((L_, M_), dtype=torch.float32, device=dev)
        # rarefied clonality: part of the reference's output (ISB_BENCH_NO_CLONTR=1: A/B switch, not a bench configuration)
        self.no_clontr = os.environ.get("ISB_BENCH_NO_CLONTR", "0") == "1"
        self.clonTR = None if self.no_clontr else torch.empty((L_, M_), dtype=torch.float32, device=dev)
        self.flags = torch.empty(L_, dtype=torch.uint8, device=dev)
        self.snv_cap = max(1 << 16, (L_ // 16) * (1 if M_ == 1 else 4))
        self.ld_cap = max(1 << 18, (L_ // 2) * (1 if M_ == 1 else 4))
        self.n_sets = 2 if two_sets else 1
        self.sets = []
        self._alloc_rows()
        if layout == "reads":
            rd = d["reads"]
            self.batch = _cabi.IsbReadsBatch(int(rd["n_segs"]), p(rd["seg_start"]), p(rd["seg_len"]), p(rd["seg_pair"]),
                                             p(rd["seg_word"]), int(rd["n_words"]), p(rd["words"]), int(rd["max_seg_len"]), 0,
                                             0, None, None, self.npairs, p(d["pair_mm"]), start_, L_, p(d["ref_codes"]),
                                             len(d["splits"]), p(d["splits"]), M_, 0)
            self.entry = eng.lib.isb_profile_reads
        elif layout == "cols":
            self.cols = cols
            self.batch = _cabi.IsbColsBatch(cols["n_groups"], p(cols["grp_off"]), cols["n_chunks"], p(cols["words"]), p(cols["ids"]), 0,
                                            None, None, self.npairs, p(d["pair_mm"]), 0, L_, p(d["ref_codes"]), len(d["splits"]),
                                            p(d["splits"]), M_, 0)
            self.entry = eng.lib.isb_profile_cols
        else:
            self.batch = _cabi.IsbBatch(int(d["n_events"]), p(d["ref_pos"]), p(d["base"]), p(d["qual"]), p(d["read_id"]), self.npairs,
                                        p(d["pair_mm"]), 0, L_, p(d["ref_codes"]), d["splits"].shape[0], p(d["splits"]), M_)
            self.entry = eng.lib.isb_profile_batch
        flags = _cabi.ISB_SKIP_LINKAGE if args.skip_linkage else 0
        self.prm = _cabi.IsbParams(5, 20, 30, flags, 0.05, 0 if self.no_clontr else 50, 0, SEED)
        # the timed loops enqueue the steps without a host round trip per call (ISB_NO_SYNC: capacities are settled in the
        # synchronous warm-up steps; errors and row counts are checked once after the loop, isb_synchronize)
        self.prm_async = _cabi.IsbParams(5, 20, 30, flags | _cabi.ISB_NO_SYNC, 0.05, 0 if self.no_clontr else 50, 0, SEED)
        self.counts4 = torch.zeros(4, dtype=torch.int64, device=dev)
        self.step_no, self.res = 0, None

    def _alloc_rows(self):
        import torch
        p = self._cabi.ptr
        self.sets = []
        for _ in range(self.n_sets):
            s_ = torch.empty(self.snv_cap * 32, dtype=torch.uint8, device=self.dev)
            l_ = torch.empty(self.ld_cap * 64, dtype=torch.uint8, device=self.dev)
            r_ = self._cabi.IsbResult(p(self.counts), p(self.nmask), p(self.covT), p(self.clonT), p(self.flags), p(s_), self.snv_cap,
                                      p(l_), self.ld_cap, 0, 0, 0, 0, p(self.clonTR))
            self.sets.append((s_, l_, r_))

    def step_async(self):
        """One step enqueued on the stream, no host synchronisation (row counts stay on the device)."""
        cur = self.step_no % self.n_sets
        rc = self.entry(self.eng.ctx, C.byref(self.batch), C.byref(self.prm_async), C.byref(self.sets[cur][2]))
        if rc != 0:
            raise RuntimeError(self.eng.lib.isb_last_error(self.eng.ctx).decode())
        self.step_no += 1
        return cur

    def check_async(self):
        """After a loop of step_async: device-side error flags, row counts of the last step (one host round trip)."""
        rc = self.eng.lib.isb_row_counts_async(self.eng.ctx, self._cabi.ptr(self.counts4))
        if rc == 0:
            rc = self.eng.lib.isb_synchronize(self.eng.ctx)
        if rc != 0:
            raise RuntimeError(self.eng.lib.isb_last_error(self.eng.ctx).decode())
        c = self.counts4.tolist()
        if c[0] > self.snv_cap or c[1] > self.ld_cap:
            raise RuntimeError("row buffers too small in the timed loop: %r" % (c,))
        r_ = self.sets[(self.step_no - 1) % self.n_sets][2]
        r_.n_snv, r_.n_ld, r_.n_sites, r_.n_site_pairs = c
        self.res = r_

    def step(self):
        cur = self.step_no % self.n_sets
        r_ = self.sets[cur][2]
        rc = self.entry(self.eng.ctx, C.byref(self.batch), C.byref(self.prm), C.byref(r_))
        if rc == self._cabi.ISB_ERR_CAPACITY:                 # only in warm-up: the row buffers grow to the counted need
            self.snv_cap, self.ld_cap = max(self.snv_cap, int(r_.n_snv) + 1024), max(self.ld_cap, int(r_.n_ld) + 1024)
            self._alloc_rows()
            cur, r_ = 0, self.sets[0][2]
            rc = self.entry(self.eng.ctx, C.byref(self.batch), C.byref(self.prm), C.byref(r_))
        if rc != 0:
            raise RuntimeError(self.eng.lib.isb_last_error(self.eng.ctx).decode())
        self.res = r_
        self.step_no += 1
        return cur


class Gather:
    """NCCL gather of the final SNV / linkage rows of a step to rank 0 -- the only collective of the job.  Fixed-size slabs
    (row counts are known after warm-up: the same data every step), receive buffers allocated once, row counts exchanged on
    the device: no host synchronisation and no allocation in the timed loop.  Enqueued on a side stream: it runs underneath
    the next step's kernels; a row-buffer set is reused only after its gather has completed."""

    def __init__(self, job, world, rank, stream, dev):
        import torch
        import torch.distributed as dist
        self.job, self.world, self.rank, self.stream, self.dev, self.dist, self.torch = job, world, rank, stream, dev, dist, torch
        self.side = torch.cuda.Stream(device=dev)
        self.done = [None] * job.n_sets
        self.counts_in = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(job.n_sets)]
        self.counts_all = torch.zeros(world * 4, dtype=torch.int64, device=dev)
        self.slab, self.recv, self.bytes_per_step = None, None, 0

    def fix_sizes(self):
        """After warm-up: slab sizes = the largest row counts over the ranks, rounded up (one-off host sync)."""
        torch, dist = self.torch, self.dist
        r_ = self.job.res
        mine = torch.tensor([int(r_.n_snv), int(r_.n_ld)], dtype=torch.int64, device=self.dev)
        dist.all_reduce(mine, op=dist.ReduceOp.MAX)
        mx = mine.tolist()
        self.slab = [min((mx[0] + 4095) // 4096 * 4096, self.job.snv_cap) * 32, min((mx[1] + 4095) // 4096 * 4096, self.job.ld_cap) * 64]
        if self.rank == 0:
            self.recv = [[torch.empty(self.slab[t], dtype=torch.uint8, device=self.dev) for _ in range(self.world)] for t in range(2)]
        self.bytes_per_step = sum(self.slab) * (self.world - 1)

    def wait_set(self, cur):
        if self.done[cur] is not None:
            self.stream.wait_event(self.done[cur])

    def launch(self, cur):
        torch, dist = self.torch, self.dist
        s_, l_, r_ = self.job.sets[cur]
        # row counts of the step: device -> device on the step's stream (isb_row_counts_async), exchanged on the side stream
        rc = self.job.eng.lib.isb_row_counts_async(self.job.eng.ctx, self.job._cabi.ptr(self.counts_in[cur]))
        if rc != 0:
            raise RuntimeError(self.job.eng.lib.isb_last_error(self.job.eng.ctx).decode())
        self.side.wait_stream(self.stream)
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(self.counts_all, self.counts_in[cur])
            for t, buf in enumerate((s_, l_)):
                dist.gather(buf[:self.slab[t]], self.recv[t] if self.rank == 0 else None, dst=0)
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.done[cur] = ev

    def finish(self):
        self.stream.wait_stream(self.side)


def time_steps(job, gather, steps, stream, world, dist, torch):
    """Exactly `steps` steps between two events on the step stream, barrier + synchronize on both sides.  Returns
    (total ms as the max over ranks, wall t0, wall t1)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        if gather is not None:
            gather.wait_set(job.step_no % job.n_sets)
        cur = job.step_async()
        if gather is not None:
            gather.launch(cur)
    if gather is not None:
        gather.finish()                                          # the last gathers belong to the timed region
    e1.record(stream)
    torch.cuda.synchronize()
    job.check_async()
    if world > 1:
        dist.barrier()
    w1 = time.time()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=job.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), w0, w1


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
    from instrain_b200.shard import lpt_partition
    lut, dflt = load_lut()
    M_label = "per-pair mm levels" if args.mm else "M=1 (--skip_mm_profiling)"
    stages = "pileup + SNV call (K1+K2)" if args.skip_linkage else "full profile (K1+K2+K3)"
    workload = "synthetic metagenome %d x %d bp, %dx coverage, %.3g SNV density, %s, %s" % (
        args.scaffolds, args.L, args.cov, args.dens, M_label, stages)
    strong = args.scaling == "strong"
    n_ranks = max(world, args.gpus) if args.impl == "reference" else world
    per_gpu = args.scaffolds / (n_ranks if strong else 1)
    config = {"workload": workload, "scaffolds": args.scaffolds, "scaffold_len": args.L, "coverage": args.cov,
              "snv_density": args.dens, "min_cov": 5, "min_freq": 0.05, "min_snp": 20, "window_length": 10000,
              "sharding": ((("the same %d scaffold(s) cut into contiguous runs of splits, one run per rank (strong scaling; reads that reach "
                             "into a run are cut at its borders), NCCL gather of SNV/linkage rows to rank 0" % args.scaffolds)
                            if args.scaffolds < n_ranks else
                            ("the same %d scaffolds LPT-partitioned over %d ranks (strong scaling), NCCL gather of SNV/linkage rows to rank 0"
                             % (args.scaffolds, n_ranks))) if strong else
                           ("%d scaffolds per rank (weak scaling), NCCL gather of SNV/linkage rows to rank 0" % args.scaffolds))
              if n_ranks > 1 else "single GPU",
              "layout": {"reads": "BAM-order aligned segments (4-bit one-hot code per aligned base, stored once per read: what the host packer emits)",
                         "cols": "column words (the same codes regrouped per 8-position column; conversion not in the step)",
                         "events": "position-major event columns (10 B per event)"}[args.layout],
              "l2": "inputs (%.1f GB per GPU) larger than L2; no flush needed" % (
                  per_gpu * args.L * args.cov * {"cols": 1.07, "reads": 0.56, "events": 10}[args.layout] / 1e9)}

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
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": args.scaling if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "int32 counts / f64 statistics", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference algorithm = oracle/oracle.c (C restatement pinned on the reference's goldens); the "
                    "reference itself is pure Python + pysam and cannot run on this box", "rows": list(rows)}))
        return 0

    # ------------------------------------------------------------------------------------------------ B200 arm
    from instrain_b200.engine import Engine
    t_gen = time.time()
    run = None
    if world > 1 and strong and args.scaffolds < world:
        # fewer scaffolds than GPUs (BASELINE config 5: one 10 Mb scaffold): every scaffold is cut into contiguous runs of
        # splits, one run per rank of its group (instrain_b200.shard.split_runs); a rank profiles its run from the reads
        # that overlap it (read halo, cut at the run's borders)
        from instrain_b200.shard import split_runs
        from instrain_b200.synth import iterate_splits
        per = world // args.scaffolds                                   # ranks per scaffold (the last group takes the rest)
        sc = min(rank // per, args.scaffolds - 1)
        grp = list(range(sc * per, world if sc == args.scaffolds - 1 else (sc + 1) * per))
        sp = iterate_splits(args.L, 10000)
        a_, b_ = split_runs(sp, [e - s_ + 1 for s_, e in sp], len(grp))[grp.index(rank)]
        run = (sp[a_][0], sp[b_ - 1][1] + 1)
        my_ids = [sc]
    elif world > 1 and strong:
        # LPT over aligned bases (all scaffolds of the synthetic set weigh the same: contiguous blocks come out)
        my_ids = lpt_partition([float(args.L) * args.cov] * args.scaffolds, world)[rank]
    else:
        my_ids = list(range(rank * args.scaffolds, (rank + 1) * args.scaffolds))
    if args.layout == "reads":
        d = build_dataset(synth, local_rank, my_ids, args)
        if run is not None:
            d = clip_dataset_to_run(d, run[0], run[1])
    else:                                                # the older layouts: one generator call, N = 1 or weak scaling only
        if world > 1 and strong:
            raise SystemExit("--layout %s is timed with --scaling weak only" % args.layout)
        d = synth.generate(local_rank, args.L, args.scaffolds, args.cov, args.dens, SEED + rank, skip_mm=not args.mm,
                           events=args.layout == "events", reads=args.layout == "cols", seg_words=args.seg_words)
    torch.cuda.synchronize()
    eng = Engine(local_rank, lut, dflt)
    cols = None
    if args.layout == "cols":
        cols = synth.reads_to_cols_device(eng, d)
        cols["n_real_words"] = int((cols["ids"] >= 0).sum().item())
        del d["reads"]
        eng.close()
        eng = Engine(local_rank, lut, dflt)
        torch.cuda.empty_cache()
    t_gen = time.time() - t_gen
    stream = torch.cuda.Stream(device=dev)                 # the steps' stream (not the legacy default stream: handle 0 would mean
    eng.set_stream(stream.cuda_stream)                     # "the context's own stream" to isb_set_stream)
    job = Job(eng, d, args, dev, layout=args.layout, cols=cols, two_sets=world > 1)
    gather = Gather(job, world, rank, stream, dev) if world > 1 else None
    n_ev, npairs, Ltot, M = int(d["n_events"]), job.npairs, job.Ltot, job.M
    L_job = args.scaffolds * args.L * (1 if strong or world == 1 else world)     # positions of the whole job

    for _ in range(max(3, args.warmup)):
        job.step()
    if gather is not None:
        gather.fix_sizes()
        for _ in range(2):
            gather.wait_set(job.step_no % job.n_sets)
            gather.launch(job.step())
        gather.finish()
    eng.enable_timing(True)
    eng.stage_times()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = eng.launch_count
    ms_total, w0, w1 = time_steps(job, gather, args.steps, stream, world, dist, torch)
    stage_ms, stage_calls = eng.stage_times()
    eng.enable_timing(False)
    launches = eng.launch_count - launches0
    clocks = sampler.window(w0, w1) if sampler else None
    ms_step = ms_total / args.steps
    value = L_job / (ms_step / 1e3)
    res = job.res
    rows_out = {"n_events_per_gpu": n_ev, "n_pairs_per_gpu": npairs, "M": M, "n_snv": int(res.n_snv), "n_ld": int(res.n_ld),
                "n_sites": int(res.n_sites), "n_site_pairs": int(res.n_site_pairs)}
    gather_bytes = gather.bytes_per_step if gather is not None else 0

    # sustained leg: the same step for >= --sustain-s seconds (power / clocks settle), reported beside
    sustained = None
    if args.sustain_s > 0:
        n_s = max(args.steps, int(args.sustain_s * 1e3 / ms_step) + 1)
        ms_s, s0, s1 = time_steps(job, gather, n_s, stream, world, dist, torch)
        sustained = {"value": L_job / (ms_s / n_s / 1e3), "unit": UNIT, "steps": n_s, "seconds": ms_s / 1e3,
                     "ms_per_step": ms_s / n_s, "clocks": sampler.window(s0, s1) if sampler else None}

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

    def rl(kernel, alg, ms, key, units, bytes_def, **extra):
        out = {"kernel": kernel, "bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak, "unit": "GB/s",
               "frac": alg / ms / 1e6 / peak, "traffic": traffic_of(key, units), "peak_source": peak_src,
               "algorithmic_bytes_per_launch": alg, "bytes_def": bytes_def, "launch_ms": ms}
        out.update(extra)
        return out

    k1_ms = stage_ms[0] / max(1, stage_calls[0])
    if args.layout == "reads":
        rd = d["reads"]
        fused = M == 1 and job.lean and os.environ.get("ISB_K1F", "1") != "0"
        in_bytes = int(rd["n_words"]) * 4 + int(rd["n_segs"]) * 14              # nibble stream + (start i32, len u16, word i64) per segment
        if fused:
            out_bytes = Ltot * (1 + 4 + 4 + 4 + 1)
            roofline = rl("k1f_pileup<M=1, fused: coverage + single-allele sites + site queue>", in_bytes + out_bytes, k1_ms, "K1f_fused_M1", Ltot,
                          "4 bits per aligned base (nibble stream incl. separators) + 14 B per segment + 1 B reference per position in; "
                          "covT 4 + clonT 4 + clonTR 4 + site_flags 1 B per position out (the ~12 % of the sites that go to the site queue "
                          "get a 24-byte queue entry instead of the last three: not counted)",
                          achieved_survey_def=(10.0 * n_ev + (8 * M + 1) * Ltot) / k1_ms / 1e6,
                          survey_def="SURVEY 8(d): 10 B per aligned base (ref_pos, base, qual, read_id columns) + 8 M + 1 B per position, "
                                     "fused; the layout here is 16x smaller, so this figure can exceed the HBM peak",
                          note="pileup from BAM-order segments; the step's other kernels: k2q_sites (general SNV call on the queued sites, "
                               "stage k2_snv), k3f_site_rows + k3_enum_pairs_tiles + k3_pair_stats_dev + k3_self_edges (stage k3_linkage)")
        else:
            alg = in_bytes + int(rd["n_segs"]) * (4 if M > 1 else 0) + 16 * M * Ltot + 8 * Ltot
            roofline = rl("k1f_pileup<M=1>" if M == 1 else "k1f_pileup<M>1>", alg, k1_ms, "K1r_M1" if M == 1 else "K1r_Mgt1", Ltot,
                          "4 bits per aligned base + 14-18 B per segment in, 16*M B/position counts + 8 B/position nmask out",
                          achieved_survey_def=(10.0 * n_ev + 16 * M * Ltot) / k1_ms / 1e6)
    elif args.layout == "cols":
        fused = M == 1 and job.lean and os.environ.get("ISB_K1C_FUSE", "1") != "0"
        in_bytes = cols["n_real_words"] * (4 if M == 1 else 8) + (cols["n_groups"] + 1) * 8
        out_bytes = (Ltot * (1 + 4 + 4 + 1) + int(res.n_snv) * 32 + int(res.n_sites) * 16) if fused else (16 * M * Ltot + 8 * Ltot)
        roofline = rl("k1c_pileup_m1<fused SNV call>" if fused else ("k1c_pileup_m1" if M == 1 else "k1c_pileup_mm"),
                      in_bytes + out_bytes, k1_ms, "K1c_fused_M1" if fused else ("K1c_M1" if M == 1 else "K1c_Mgt1"), Ltot,
                      "column words (4 bits per aligned base, padding not counted)%s + offsets in; %s out" % (
                          " + 4 B pair id per word" if M > 1 else "", "covT + clonT + site_flags, SNV rows, counts of linkage sites"
                          if fused else "16*M B/position counts + 8 B/position nmask"),
                      note="resident pre-transposed layout: the reads -> columns conversion is NOT in this step")
    else:
        ev_bytes = 10 if M > 1 else 6
        alg = n_ev * ev_bytes + 16 * M * Ltot + 8 * Ltot
        roofline = rl("k1_pileup_tiles_tma", alg, k1_ms, "M1" if M == 1 else "Mgt1", n_ev,
                      "%d B/event + 16*M B/position counts + 8 B/position nmask" % ev_bytes)
    roofline["stage_ms_per_step"] = stage_per_step
    # the same bytes over the whole step (all stages): what the job as a whole achieves against the HBM peak
    roofline["step_frac"] = roofline["algorithmic_bytes_per_launch"] / ms_step / 1e6 / peak if world == 1 or not strong else None

    # -------------------------------------------------------------------------------- weak scaling beside (N > 1)
    weak = None
    if world > 1 and strong and not args.no_weak and args.layout == "reads" and run is None:
        del job, gather, d
        torch.cuda.empty_cache()
        dw = build_dataset(synth, local_rank, list(range(rank * args.scaffolds, (rank + 1) * args.scaffolds)), args)
        jw = Job(eng, dw, args, dev, two_sets=True)
        gw = Gather(jw, world, rank, stream, dev)
        for _ in range(3):
            jw.step()
        gw.fix_sizes()
        n_w = max(5, args.steps // 2)
        ms_w, _, _ = time_steps(jw, gw, n_w, stream, world, dist, torch)
        weak = {"value": world * args.scaffolds * args.L / (ms_w / n_w / 1e3), "unit": UNIT, "ms_per_step": ms_w / n_w,
                "scaffolds_per_rank": args.scaffolds, "steps": n_w, "gather_bytes_per_step": gw.bytes_per_step}
        d, job, gather = dw, jw, gw                              # the e2e slice below is cut from this rank's data

    # -------------------------------------------------------------------------------- other resident layouts, for comparison
    other = None
    if world == 1 and args.layout == "reads" and args.also_layouts > 0 and not args.mm and not args.skip_linkage:
        other = {}
        n_sc = min(args.also_layouts, args.scaffolds)
        de = synth.generate(local_rank, args.L, n_sc, args.cov, args.dens, SEED, skip_mm=True, events=True, reads=True,
                            seg_words=args.seg_words)
        for lay in ("cols", "events"):
            e2 = Engine(local_rank, lut, dflt)
            e2.set_stream(stream.cuda_stream)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            conv_ms, cd2 = None, None
            if lay == "cols":
                synth.reads_to_cols_device(e2, de)                                 # warm (scratch allocation)
                torch.cuda.synchronize()
                ev0.record(stream)
                cd2 = synth.reads_to_cols_device(e2, de)
                ev1.record(stream)
                torch.cuda.synchronize()
                conv_ms = ev0.elapsed_time(ev1)
            j2 = Job(e2, de, args, dev, layout=lay, cols=cd2)
            for _ in range(3):
                j2.step()
            e2.enable_timing(True)
            e2.stage_times()
            ms2, _, _ = time_steps(j2, None, 5, stream, 1, dist, torch)
            sm, _ = e2.stage_times()
            other[lay] = {"value": n_sc * args.L / (ms2 / 5 / 1e3), "unit": UNIT, "ms_per_step": ms2 / 5, "scaffolds": n_sc,
                          "stage_ms_per_step": {"k1_pileup": sm[0] / 5, "k2_snv": sm[1] / 5, "k3_linkage": sm[2] / 5}}
            if conv_ms is not None:
                other[lay]["conversion_ms_not_in_step"] = conv_ms
                other[lay]["value_incl_conversion"] = n_sc * args.L / ((ms2 / 5 + conv_ms) / 1e3)
            del j2, cd2
            e2.close()
        del de
        torch.cuda.empty_cache()

    # -------------------------------------------------------------------------------- e2e: host buffers through the C-ABI
    # The call a user of the library makes with HOST memory: pinned host buffers in (the packer's reference-delta transfer
    # format of the same BAM-order segments), isb_profile_reads_delta (H2D + K0d + the same K1f / linkage kernels + D2H of
    # every result table), pinned host tables out.  Every rank runs its own slice; the job's e2e is the sum over ranks.
    e2e = None
    if args.layout == "reads" and not args.no_e2e and run is None:    # (a run of splits: no bounded host slice is cut)
        from instrain_b200.reads import delta_reads_host
        n_sc = max(1, min(args.e2e_scaffolds, d["n_scaffolds"]))
        hs = synth.reads_to_host(d, 0, n_sc)
        hr = hs["reads"]
        Ls = n_sc * args.L
        Ms = int(hs["pair_mm"].max()) + 1 if len(hs["pair_mm"]) else 1
        t_enc = time.time()
        hd = delta_reads_host(hr, hs["ref_codes"])          # the C++ encoder the packer runs (one host thread)
        t_enc = time.time() - t_enc
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h = {k: pin(hs[k]) for k in ("pair_mm", "ref_codes", "splits")}
        hseg = {"seg_start": pin(hr["seg_start"]), "seg_len": pin(hr["seg_len"].view(np.int16)), "seg_pair": pin(hr["seg_pair"])}
        hdp = {"pass": pin(hd["pass"]), "mis_word": pin(hd["mis_word"].view(np.int32)), "mis_code": pin(hd["mis_code"])}
        p = _cabi.ptr
        dbatch = _cabi.IsbReadsDelta(int(hr["n_segs"]), p(hseg["seg_start"]), p(hseg["seg_len"]), p(hseg["seg_pair"]),
                                     int(hd["n_units"]), p(hdp["pass"]), len(hd["mis_word"]), p(hdp["mis_word"]), p(hdp["mis_code"]),
                                     int(hr["max_seg_len"]), 0, 0, None, None, len(hs["pair_mm"]), p(h["pair_mm"]), 0, Ls,
                                     p(h["ref_codes"]), len(hs["splits"]), p(h["splits"]), Ms, 0)

        def host_result():
            o = dict(covT=torch.empty((Ls, Ms), dtype=torch.int32).pin_memory(),
                     clonT=torch.empty((Ls, Ms), dtype=torch.float32).pin_memory(),
                     clonTR=torch.empty((Ls, Ms), dtype=torch.float32).pin_memory(),
                     flags=torch.empty(Ls, dtype=torch.uint8).pin_memory(),
                     snv=torch.empty(max(1 << 16, (Ls // 16) * (1 if Ms == 1 else 16)) * 32, dtype=torch.uint8).pin_memory(),
                     ld=torch.empty(max(1 << 18, (Ls // 2) * (1 if Ms == 1 else 8)) * 64, dtype=torch.uint8).pin_memory())
            return o, _cabi.IsbResult(None, None, p(o["covT"]), p(o["clonT"]), p(o["flags"]), p(o["snv"]), o["snv"].numel() // 32,
                                      p(o["ld"]), o["ld"].numel() // 64, 0, 0, 0, 0, p(o["clonTR"]))

        o1, hres = host_result()
        prm = _cabi.IsbParams(5, 20, 30, _cabi.ISB_SKIP_LINKAGE if args.skip_linkage else 0, 0.05, 50, 0, SEED)
        fn = eng.lib.isb_profile_reads_delta

        def call(ctx_, res_):
            rc_ = fn(ctx_, C.byref(dbatch), C.byref(prm), C.byref(res_))
            if rc_ != 0:
                raise RuntimeError(eng.lib.isb_last_error(ctx_).decode())

        for _ in range(3):
            call(eng.ctx, hres)
        torch.cuda.synchronize()
        a = time.time()
        call(eng.ctx, hres)
        t_one = time.time() - a
        n_calls = int(min(400, max(50, 1.0 / max(t_one, 1e-4))))
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(n_calls):
            call(eng.ctx, hres)
        ev1.record(stream)
        torch.cuda.synchronize()
        dt = ev0.elapsed_time(ev1) / 1e3 / n_calls

        # Two (and three) contexts on as many host threads (each call is still host buffers -> C-ABI -> host tables): the H2D
        # copy of one call overlaps the kernels and the D2H copy of the others, which is how a host pipeline feeds the GPU.
        dt_pipe, dt_pipe3 = None, None
        try:
            extra = [(Engine(local_rank, lut, dflt),) + host_result() for _ in range(2 if world == 1 else 1)]     # (engine, host tables, isb_result)
            errs = []

            def worker(c, r, n):
                try:
                    for _ in range(n):
                        call(c, r)
                except Exception as ex:                      # noqa: BLE001
                    errs.append(str(ex))

            for n_ctx in ((2, 3) if world == 1 else (2,)):          # N > 1: the ranks already share the host's cores and PCIe root
                ctxs = [(eng.ctx, hres)] + [(e_[0].ctx, e_[2]) for e_ in extra[:n_ctx - 1]]
                for n_it in (4, max(25, n_calls // 2)):     # the first repetition warms the extra contexts' scratch buffers
                    torch.cuda.synchronize()
                    t_a = time.time()
                    ths = [threading.Thread(target=worker, args=(c_, r_, n_it)) for c_, r_ in ctxs]
                    [t.start() for t in ths]
                    [t.join() for t in ths]
                    torch.cuda.synchronize()
                    d_ = (time.time() - t_a) / (n_ctx * n_it)
                if n_ctx == 2:
                    dt_pipe = d_
                else:
                    dt_pipe3 = d_
            if errs:
                raise RuntimeError(errs[0])
            for e_ in extra:
                e_[0].close()
        except Exception as ex:                              # noqa: BLE001 - the pipelined figures are optional
            dt_pipe = dt_pipe3 = None
            print("e2e pipelined leg skipped: %r" % (ex,), file=sys.stderr)
        best = min([d_ for d_ in (dt, dt_pipe, dt_pipe3) if d_])
        common = len(hs["pair_mm"]) + Ls + hs["splits"].nbytes
        h2d = hd["n_units"] + len(hd["mis_word"]) * 5 + hr["n_segs"] * (4 + 2 + 4) + common
        d2h = Ls * Ms * 12 + Ls + int(hres.n_snv) * 32 + int(hres.n_ld) * 64
        tot = torch.tensor([Ls / best, Ls / dt, Ls / dt_pipe if dt_pipe else 0.0, Ls / dt_pipe3 if dt_pipe3 else 0.0],
                           dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        tot = tot.tolist()
        e2e = {"value": tot[0], "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": best * 1e3, "calls_timed": n_calls,
               "timer": "CUDA events on the context stream around %d back-to-back calls (single context); wall clock around the multi-context legs" % n_calls,
               "api": "isb_profile_reads_delta: BAM-order segments in the reference-delta transfer format (event bits + one entry per base "
                      "that differs from the reference) -> K0d -> the same K1f / linkage kernels as `value` -> host tables",
               "ranks": world, "per_rank_slice": "%d scaffolds (%d positions) per call, pinned host buffers -> C-ABI -> pinned host result tables" % (n_sc, Ls),
               "single_context": {"value": tot[1], "ms_per_step": dt * 1e3},
               "two_contexts_pipelined": ({"value": tot[2], "ms_per_step": dt_pipe * 1e3} if dt_pipe else None),
               "three_contexts_pipelined": ({"value": tot[3], "ms_per_step": dt_pipe3 * 1e3} if dt_pipe3 else None),
               "host_encode": {"what": "reference-delta encoding of the slice on ONE host core (isb_reads_delta_host), outside the timed "
                                       "region: the packer threads do it while the GPU works", "positions_per_s_per_core": Ls / t_enc}}

    # -------------------------------------------------------------------------------- from a BAM file on disk
    # The whole product call, file in / SNVprofile directory out: a coordinate-sorted, indexed BAM of the workload's shape
    # (instrain_b200/synth_bam.py writes BGZF / BAM / BAI itself) -> profile_bam: C++ read filter -> C++ packer on every
    # host core (BGZF inflate, htslib's overlap tweak, CIGAR expansion) -> H2D -> the same kernels -> tables -> .hd5 store.
    from_bam = None
    if rank == 0 and world == 1 and args.from_bam_scaffolds > 0 and args.layout == "reads":
        import shutil
        import tempfile
        from instrain_b200 import synth_bam
        from instrain_b200.packer import BamPacker
        from instrain_b200.profile import iter_batches, profile_bam
        from instrain_b200.read_filter import filter_reads
        tmp = tempfile.mkdtemp(prefix="isb_bench_")
        try:
            n_sc, cores = args.from_bam_scaffolds, host_cores()
            path = os.path.join(tmp, "synth.bam")
            t0 = time.time()
            Lsc = args.from_bam_L
            info = synth_bam.write_bam(path, Lsc, n_sc, args.cov, args.dens, SEED)
            t_write = time.time() - t0
            Lb = n_sc * Lsc
            kw = dict(s2s=info["seqs"], packer_threads=cores, skip_mm_profiling=not args.mm, seed=SEED)
            profile_bam(path, None, None, os.path.join(tmp, "warm.IS"), **kw)        # warm: CUDA context, scratch, page cache
            t0 = time.time()
            out = profile_bam(path, None, None, os.path.join(tmp, "run.IS"), **kw)
            t_run = time.time() - t0
            tm = out.result.timing
            # the packer alone on all cores: aligned bases per second and per core
            with BamPacker(path) as bp:
                names = bp.ref_names
            t0 = time.time()
            r2m, _, _ = filter_reads(path, names)
            t_filter = time.time() - t0
            if not args.mm:
                r2m = {k: set(v) for k, v in r2m.items()}
            t0 = time.time()
            n_bases = sum(b["n_events"] for k, b in iter_batches(path, r2m, info["seqs"], packer_threads=cores) if k == "batch")
            t_pack = time.time() - t0
            t0 = time.time()
            n_bases1 = sum(b["n_events"] for k, b in iter_batches(path, r2m, info["seqs"], packer_threads=1) if k == "batch")
            t_pack1 = time.time() - t0
            from_bam = {"value": Lb / t_run, "unit": UNIT, "seconds": t_run, "positions": Lb, "scaffolds": n_sc, "scaffold_len": Lsc,
                        "coverage": args.cov, "bam_bytes": info["bytes"],
                        "reads": info["n_reads"], "aligned_bases": info["aligned_bases"], "host_threads": cores,
                        "what": "profile_bam(bam, Fdb=None, sR2M=None, ISP_loc, packer_threads=%d): file in, SNVprofile directory out "
                                "(read filter + mapping_info, packer, engine, tables, .hd5 store)" % cores,
                        "breakdown_s": {"read_filter_and_report": tm.get("read_filter_s"), "pack_plus_engine_plus_tables": tm.get("profile_scaffolds_s"),
                                        "engine_calls": tm.get("engine_s"), "store": tm.get("store_s")},
                        "gpu_idle_fraction": 1.0 - tm.get("engine_s", 0.0) / t_run,
                        "packer": {"aligned_bases_per_s_all_threads": n_bases / t_pack, "aligned_bases_per_s_one_thread": n_bases1 / t_pack1,
                                   "per_core_at_%d_threads" % cores: n_bases / t_pack / cores, "seconds_all_threads": t_pack,
                                   "read_filter_s_one_thread": t_filter},
                        "snv_rows": len(out.result.raw_snp_table), "linkage_rows": len(out.result.raw_linkage_table),
                        "bam_write_s_not_timed": t_write}
            if not args.no_cpu_baseline:                             # the oracle port on the same file's events (checker / baseline)
                evs, off, npair = [], 0, 0
                with BamPacker(path) as bp:
                    while True:
                        tid = bp.peek_tid()
                        if tid < 0:
                            break
                        nm_ = bp.ref_names[tid]
                        ev = bp.pack_scaffold(tid, r2m.get(nm_, {}), pos_offset=off, pair_id_offset=npair)
                        evs.append(ev)
                        off += Lsc
                        npair += len(ev["pair_mm"])
                hb = {k: np.concatenate([e[k] for e in evs]) for k in ("ref_pos", "base", "qual", "read_id", "pair_mm")}
                from instrain_b200.profile import encode_reference
                hb["ref_codes"] = np.concatenate([encode_reference(info["seqs"][n_]) for n_ in info["names"]])
                hb["splits"] = synth.batch_splits(Lsc, n_sc)
                t0 = time.time()
                rows_c = cpu_oracle_pass(hb, lut, dflt, cores)
                t_cpu = time.time() - t0
                from_bam["cpu_port_same_bam"] = {"value": Lb / t_cpu, "unit": UNIT, "cores": cores, "seconds_compute_only": t_cpu,
                                                 "rows": list(rows_c), "note": "oracle/oracle.c on the events the C++ packer makes of the same "
                                                                               "BAM; BAM decoding not included"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    # -------------------------------------------------------------------------------- CPU baseline (oracle port) beside
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sc = max(1, min(args.cpu_scaffolds, args.scaffolds))
        ds = synth.generate(local_rank, args.L, n_sc, args.cov, args.dens, SEED, skip_mm=not args.mm)
        hb = synth.to_host_batch(ds, 0, n_sc)
        del ds
        a = time.time()
        cpu_oracle_pass(hb, lut, dflt, 1)
        dt_c = time.time() - a
        cpu = {"value": n_sc * args.L / dt_c, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d scaffold(s) x %d bp at %dx (%d events) of the same workload, one pass, 1 thread of oracle/oracle.c (orc_profile_mt)"
                         % (n_sc, args.L, args.cov, len(hb["ref_pos"]))}
    if sampler:
        sampler.stop()

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": "int32 counts / f64 statistics", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "sustained": sustained,
            "weak_scaling": weak, "gather_bytes_per_step": gather_bytes, "from_bam": from_bam, "other_layouts": other, "rows": rows_out,
            "setup_s": round(t_gen, 1)}))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
