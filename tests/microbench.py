"""Per-stage device timings on synthetic data (development tool; run on the GPU box through gpurun).

    python tests/microbench.py [--L 200000] [--cov 100] [--rep 10] [--mm]
Events are generated on the host with oracle/synth.py (test infrastructure used as a data generator only),
replicated `rep` times on the device with shifted coordinates, and each stage is timed with CUDA events on the
context's stream.  Prints one JSON line per measurement.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from instrain_b200 import _cabi  # noqa: E402
from instrain_b200.engine import Engine  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=200000)
    ap.add_argument("--cov", type=int, default=100)
    ap.add_argument("--dens", type=float, default=0.01)
    ap.add_argument("--rep", type=int, default=10)
    ap.add_argument("--mm", action="store_true", help="keep per-pair mm (M ~ 12) instead of --skip_mm_profiling")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--seglen", type=int, default=150, help="max segment length of the read-major batch")
    ap.add_argument("--reads", action="store_true", help="also time the read-major path (K1r, isb_profile_reads)")
    args = ap.parse_args()

    t0 = time.time()
    b = synth.make_batch(args.L, args.cov, args.dens, 20260103, skip_mm=not args.mm)
    n1, L1, np1 = len(b["ref_pos"]), len(b["ref_codes"]), len(b["pair_mm"])
    M = int(b["pair_mm"].max()) + 1
    dev = torch.device("cuda:0")
    R = args.rep
    pos = torch.from_numpy(b["ref_pos"]).to(dev)
    rid = torch.from_numpy(b["read_id"]).to(dev)
    pos = torch.cat([pos + r * L1 for r in range(R)]).contiguous()
    rid = torch.cat([rid + r * np1 for r in range(R)]).contiguous()
    base = torch.from_numpy(b["base"]).to(dev).repeat(R).contiguous()
    qual = torch.from_numpy(b["qual"]).to(dev).repeat(R).contiguous()
    mm = torch.from_numpy(b["pair_mm"].astype(np.uint8)).to(dev).repeat(R).contiguous()
    ref = torch.from_numpy(b["ref_codes"]).to(dev).repeat(R).contiguous()
    splits = torch.from_numpy(np.concatenate([b["splits"] + r * L1 for r in range(R)]).astype(np.int32)).to(dev)
    n, L, npairs = n1 * R, L1 * R, np1 * R
    print(json.dumps(dict(setup_s=round(time.time() - t0, 1), n_events=n, L=L, n_pairs=npairs, M=M)), flush=True)

    eng = Engine(0)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    counts = torch.empty((L, M, 4), dtype=torch.int32, device=dev)
    nmask = torch.empty(L, dtype=torch.int64, device=dev)
    covT = torch.empty((L, M), dtype=torch.int32, device=dev)
    clonT = torch.empty((L, M), dtype=torch.float32, device=dev)
    flags = torch.empty(L, dtype=torch.uint8, device=dev)
    snv = torch.empty(max(1024, L // 4) * 32, dtype=torch.uint8, device=dev)
    ld = torch.empty(max(1 << 16, L // 2) * 64, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = dict(ref_pos=pos, base=base, qual=qual, read_id=rid, pair_mm=mm)
    p = _cabi.ptr
    lib, ctx = eng.lib, eng.ctx

    def timed(name, fn, bytes_alg):
        ts = []
        for it in range(args.iters + 2):
            flush.zero_()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            z.record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(z))
        ms = float(np.median(ts))
        print(json.dumps(dict(stage=name, ms=round(ms, 4), gbs=round(bytes_alg / ms / 1e6, 1),
                              mpos_s=round(L / ms / 1e3, 1))), flush=True)

    import ctypes as C
    nrows = C.c_int64(0)

    def k1(flags_=0):
        rc = lib.isb_pileup_counts(ctx, n, p(pos), p(base), p(qual), p(rid), npairs, p(mm), 0, L, M, 30, flags_,
                                   p(counts), p(nmask))
        assert rc == 0, lib.isb_last_error(ctx)

    def k2():
        rc = lib.isb_call_snvs(ctx, L, M, p(counts), p(nmask), p(ref), 0, 5, 0.05, p(covT), p(clonT), p(flags), p(snv),
                               snv.numel() // 32, C.byref(nrows))
        assert rc == 0, lib.isb_last_error(ctx)

    def k3():
        rc = lib.isb_linkage(ctx, n, p(pos), p(base), p(qual), p(rid), npairs, p(mm), 0, L, M, 30, p(counts), p(nmask),
                             p(flags), splits.shape[0], p(splits), 20, p(ld), ld.numel() // 64, C.byref(nrows))
        assert rc == 0, lib.isb_last_error(ctx)

    ev_bytes = (10 if M > 1 else 6) * n
    timed("k1_tiles", lambda: k1(0), ev_bytes + 16 * M * L + 8 * L)
    timed("k1_atomic", lambda: k1(_cabi.ISB_K1_ANY_ORDER), ev_bytes + 16 * M * L)
    k1(0)
    timed("k2_snv", k2, (16 * M + 1 + 8) * L + 8 * M * L + L)
    print(json.dumps(dict(n_snv=int(nrows.value))), flush=True)
    timed("k3_linkage", k3, 0)
    print(json.dumps(dict(n_ld=int(nrows.value))), flush=True)

    if args.reads:
        from instrain_b200 import reads as reads_mod
        t1 = time.time()
        rd1 = reads_mod.events_to_reads(b, max_len=args.seglen)
        nw1 = rd1["n_words"]
        seg_start = torch.cat([torch.from_numpy(rd1["seg_start"]).to(dev) + r * L1 for r in range(R)]).contiguous()
        seg_len = torch.from_numpy(rd1["seg_len"].view(np.int16)).to(dev).repeat(R).contiguous()
        seg_pair = torch.cat([torch.from_numpy(rd1["seg_pair"]).to(dev) + r * np1 for r in range(R)]).contiguous()
        seg_word = torch.cat([torch.from_numpy(rd1["seg_word"]).to(dev) + r * nw1 for r in range(R)]).contiguous()
        words = torch.from_numpy(rd1["words"].view(np.int32)).to(dev).repeat(R).contiguous()
        print(json.dumps(dict(reads_setup_s=round(time.time() - t1, 1), n_segs=rd1["n_segs"] * R, n_words=nw1 * R,
                              max_seg_len=rd1["max_seg_len"])), flush=True)
        rb = _cabi.IsbReadsBatch(rd1["n_segs"] * R, p(seg_start), p(seg_len), p(seg_pair), p(seg_word), nw1 * R, p(words),
                                 rd1["max_seg_len"], 0, 0, None, None, npairs, p(mm), 0, L, p(ref), splits.shape[0], p(splits), M, 0)
        counts_r = torch.empty_like(counts)
        nmask_r = torch.empty_like(nmask)

        def k1r():
            rc = lib.isb_pileup_reads(ctx, C.byref(rb), p(counts_r), p(nmask_r))
            assert rc == 0, lib.isb_last_error(ctx)

        k1(0)
        timed("k1r_reads", k1r, nw1 * R * 4 + 22 * rd1["n_segs"] * R + 16 * M * L + 8 * L)
        print(json.dumps(dict(k1r_equal=bool(torch.equal(counts, counts_r)), nmask_equal=bool(torch.equal(nmask, nmask_r)))),
              flush=True)
        prm_r = _cabi.IsbParams(5, 20, 30, 0, 0.05)
        res_r = _cabi.IsbResult(p(counts_r), p(nmask_r), p(covT), p(clonT), p(flags), p(snv), snv.numel() // 32, p(ld),
                                ld.numel() // 64, 0, 0, 0, 0)

        def full_r():
            rc = lib.isb_profile_reads(ctx, C.byref(rb), C.byref(prm_r), C.byref(res_r))
            assert rc == 0, lib.isb_last_error(ctx)

        eng.enable_timing(True)
        timed("profile_reads", full_r, nw1 * R * 4 + 40 * M * L)
        ms, calls = eng.stage_times()
        print(json.dumps(dict(reads_stage_ms=[round(ms[k] / max(calls[k], 1), 4) for k in range(3)],
                              n_snv=int(res_r.n_snv), n_ld=int(res_r.n_ld))), flush=True)
        eng.enable_timing(False)

    batch = _cabi.IsbBatch(n, p(pos), p(base), p(qual), p(rid), npairs, p(mm), 0, L, p(ref), splits.shape[0], p(splits), M)
    prm = _cabi.IsbParams(5, 20, 30, 0, 0.05)
    res = _cabi.IsbResult(p(counts), p(nmask), p(covT), p(clonT), p(flags), p(snv), snv.numel() // 32, p(ld),
                          ld.numel() // 64, 0, 0, 0, 0)

    def full():
        rc = lib.isb_profile_batch(ctx, C.byref(batch), C.byref(prm), C.byref(res))
        assert rc == 0, lib.isb_last_error(ctx)

    timed("profile_batch", full, 10 * n + 40 * M * L + L)
    print(json.dumps(dict(n_snv=int(res.n_snv), n_ld=int(res.n_ld), n_sites=int(res.n_sites),
                          n_site_pairs=int(res.n_site_pairs), launches=eng.launch_count)), flush=True)


if __name__ == "__main__":
    main()
