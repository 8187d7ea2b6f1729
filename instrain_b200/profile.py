"""Host-side mirror of the reference's profiling entry points for the hot path.

    profile_bam(bam, Fdb, sR2M, ISP_loc, **kwargs)          <- inStrain.profile.profile_bam   (profile/__init__.py:7-18)

Same arguments and keyword names as the reference (`kwargs` = vars(args) + s2s + s2p, controller.py:337-343).
Where the reference farms (scaffold, split) commands to worker processes (profile_controller.py:243-271) and each worker
runs profile_split's per-column Python loop (profile_utilities.py:115-266), this shim
  1. streams the BAM once through the C++ host packer (instrain_b200/packer.py),
  2. concatenates scaffolds into batches (one int32 coordinate space per batch) of READ-MAJOR aligned segments
     (4-bit code per aligned base; instrain_b200/reads.py),
  3. runs K1r -> K2 -> K3 for the whole batch through the C-ABI (isb_profile_reads),
  4. turns the row arrays into the reference's per-scaffold tables (instrain_b200/tables.py).
There is no CPU fallback: without the CUDA library / a GPU this raises.
"""
import logging
import os
import time

import numpy as np
import pandas as pd

from . import summary, tables
from .engine import Engine
from . import reads as reads_mod
from .packer import BamPacker
from .synth import iterate_splits

_REF_LUT = np.full(256, 4, dtype=np.uint8)
for _i, _b in enumerate("ACTG"):
    _REF_LUT[ord(_b)] = _i


def encode_reference(seq):
    """Upper-cased reference string (fasta.py:25-27) -> base codes 0..3 (A,C,T,G), 4 = anything else."""
    return _REF_LUT[np.frombuffer(seq.encode(), dtype=np.uint8)]


class ScaffoldProfile:
    """What the reference's merged ScaffoldSplitObject carries for one scaffold (profile_utilities.py:719-820):
    raw_snp_table, raw_linkage_table, covT, clonT, clonTR (+ length).  clonTR and the normalized linkage columns are
    re-drawn quantities: unseeded in the reference, keyed by kwargs["seed"] here."""

    def __init__(self, scaffold, length):
        self.scaffold, self.length = scaffold, length
        self.raw_snp_table = self.raw_linkage_table = None
        self.covT, self.clonT, self.clonTR = {}, {}, {}
        self.pileup_counts = None       # [length, 4] A,C,T,G counts over all mm levels, with kwargs["store_everything"] only


class ProfileResult:
    """Container returned by profile_bam when inStrain.SNVprofile is not importable: same attribute names the
    reference stores (profile_utilities.py:670-715)."""

    def __init__(self):
        self.scaffold_list = []
        self.scaffolds = {}
        self.raw_snp_table = self.raw_linkage_table = self.cumulative_scaffold_table = self.cumulative_snv_table = None
        self.timing = {}
        self.failures = []
        self.store = None                     # SNVprofileStore at ISP_loc once profile_bam has written it

    def get(self, name):
        return getattr(self, name)


class _SplitTable:
    """Split table per scaffold: from the reference's Fdb (fasta.py:30-52) when given -- grouped ONCE (a filter per
    scaffold is quadratic in the number of scaffolds, and a metagenome assembly has 1e5 .. 1e6 of them) -- else the same
    geometry computed from the length."""

    def __init__(self, Fdb, window_length):
        self.window_length = window_length
        self.by_scaffold = None
        if Fdb is not None:
            db = Fdb.sort_values(["scaffold", "start"], kind="stable")
            names = db["scaffold"].values
            starts, ends = db["start"].values.astype(np.int64), db["end"].values.astype(np.int64)
            cut = np.nonzero(names[1:] != names[:-1])[0] + 1 if len(names) else np.zeros(0, np.int64)
            lo = np.concatenate([[0], cut]).astype(np.int64)
            hi = np.concatenate([cut, [len(names)]]).astype(np.int64)
            self.by_scaffold = {names[a]: (starts[a:b], ends[a:b]) for a, b in zip(lo, hi) if b > a}

    def __call__(self, scaffold, length):
        if self.by_scaffold is not None:
            st, en = self.by_scaffold.get(scaffold, ((), ()))
            return [(int(a), int(b)) for a, b in zip(st, en)]
        return iterate_splits(length, self.window_length)


def _fdb_splits(Fdb, scaffold, length, window_length):
    """Split table of one scaffold (one-off lookups; the batch stream uses _SplitTable)."""
    return _SplitTable(Fdb if Fdb is None else Fdb[Fdb["scaffold"] == scaffold], window_length)(scaffold, length)


def _rss():
    try:
        import psutil
        return psutil.Process(os.getpid()).memory_info().rss
    except Exception:                                                     # noqa: BLE001 - psutil is optional here
        return 0


def worker_log(worker_type, unit, status, t=None):
    """A line of the reference's run log (inStrain/logUtils.py:940-975, get_worker_log): "WorkerLog worker_type unit status
    RAM time PID" -- what its log parser (logUtils.py:167, 401-470) turns into the per-split / per-scaffold run report.
    Emitted for every (scaffold, split) of a batch (SplitProfile, profile_utilities.py:135,213) and every scaffold
    (MergeProfile, :756-802) with the batch's start / end times: a batch is the unit of work here."""
    assert status in ("start", "end")
    return "\nWorkerLog {0} {1} {2} {3} {4} {5}".format(worker_type, unit, status, _rss(), time.time() if t is None else t, os.getpid())


def log_checkpoint(log_class, name, status):
    """"Checkpoint class task status RAM" (inStrain/logUtils.py:903-937)."""
    assert status in ("start", "end") and len(name.split()) == 1
    logging.debug("Checkpoint {0} {1} {2} {3}".format(log_class, name, status, _rss()))


def _r2m_levels(r2m):
    """mm levels a scaffold's R2M needs on the device (1 in set mode)."""
    if isinstance(r2m, dict) and r2m:
        return min(int(max(r2m.values())), 63) + 1
    return 1


def _new_batch():
    return dict(names=[], off=[], ref=[], splits=[], n_splits=[], parts=[], pair_mm=[], pair_off=[], n_events=0, L=0,
                n_pairs=0, M=1, pad=0)


def _add_to_batch(batch, name, seq, ev, splits):
    """Append one packed scaffold (positions already offset by batch["L"]; pair ids by batch["n_pairs"])."""
    L = len(seq)
    batch["names"].append(name)
    batch["off"].append(batch["L"])
    batch["pair_off"].append(batch["n_pairs"])
    batch["ref"].append(encode_reference(seq))
    batch["splits"].extend((s + batch["L"], e + batch["L"]) for s, e in splits)
    batch["n_splits"].append(len(splits))
    batch["parts"].append(ev)
    batch["pair_mm"].append(ev["pair_mm"])
    batch["n_events"] += ev["n_events"]
    batch["L"] += L
    batch["n_pairs"] += len(ev["pair_mm"])
    if len(ev["pair_mm"]):
        batch["M"] = max(batch["M"], int(ev["pair_mm"].max()) + 1)


def _sub_batch(batch, i0, i1):
    """Scaffolds [i0, i1) of a batch as a batch of their own (coordinates, pair ids and splits re-based to 0): what a
    failed batch is bisected into, so that one bad scaffold does not take its neighbours down."""
    sub = _new_batch()
    d_pair = batch["pair_off"][i0]
    # the packed nibble words are aligned to 8-position columns of the BATCH coordinates: coordinates may only be shifted
    # by a multiple of 8, so the sub-batch starts with up to 7 unused positions (`pad`: no reads, reference code "other")
    sub["pad"] = sub["L"] = batch["off"][i0] & 7
    d_pos = batch["off"][i0] - sub["pad"]
    s0 = sum(batch["n_splits"][:i0])
    for k in range(i0, i1):
        ev = dict(batch["parts"][k])
        for key, d in (("seg_start", d_pos), ("nev_pos", d_pos), ("seg_pair", d_pair), ("nev_pair", d_pair)):
            ev[key] = ev[key] - np.int32(d) if d else ev[key]
        n_sp = batch["n_splits"][k]
        splits = [(s - batch["off"][k], e - batch["off"][k]) for s, e in batch["splits"][s0:s0 + n_sp]]
        s0 += n_sp
        sub["names"].append(batch["names"][k])
        sub["off"].append(sub["L"])
        sub["pair_off"].append(sub["n_pairs"])
        sub["ref"].append(batch["ref"][k])
        sub["splits"].extend((s + sub["L"], e + sub["L"]) for s, e in splits)
        sub["n_splits"].append(n_sp)
        sub["parts"].append(ev)
        sub["pair_mm"].append(ev["pair_mm"])
        sub["n_events"] += ev["n_events"]
        sub["L"] += len(batch["ref"][k])
        sub["n_pairs"] += len(ev["pair_mm"])
        if len(ev["pair_mm"]):
            sub["M"] = max(sub["M"], int(ev["pair_mm"].max()) + 1)
    return sub


# Dense per-position outputs cost 24 bytes per (position, mm level) on the device (counts + covT + clonT, isb_api.cu) and
# 8 on the host; a batch is closed before L * M exceeds this many cells (8e8 cells = 19 GB of HBM, 6.4 GB of host memory)
MAX_BATCH_CELLS = 800_000_000


def iter_batches(bam, sR2M, s2s, Fdb=None, window_length=10000, max_batch_events=400_000_000, packer_threads=1, debug=False,
                 max_batch_cells=MAX_BATCH_CELLS):
    """Stream the BAM through the host packer and yield ("batch", batch dict) for every batch of scaffolds (one int32
    coordinate space each) and ("failure", scaffold) for the reference's fault-injection scaffold
    (profile_utilities.py:137-139, test_profile_17).

    packer_threads == 1: one sequential pass over the BAM (batches are closed by the number of events actually packed).
    packer_threads > 1 : scaffolds are packed concurrently by that many host threads, each seeking through the .bai
    index (instrain_b200.packer.pack_scaffolds_parallel); batches are planned up front from the scaffold lengths and the
    R2M sizes (300 aligned bases assumed per read pair).  Both give the same batches whenever everything fits one."""
    split_table = _SplitTable(Fdb, window_length)

    def full(b_L, b_M, b_events, L, M):
        """Would adding a scaffold of L positions / M levels overflow the coordinate space, the event or the cell budget?"""
        return b_events > 0 and (b_L + L >= 2 ** 31 - 1 or b_events > max_batch_events or
                                 (b_L + L) * max(b_M, M) > max_batch_cells)

    if packer_threads <= 1:
        batch = _new_batch()
        with BamPacker(bam) as bp:
            while True:
                tid = bp.peek_tid()
                if tid < 0:
                    break
                name = bp.ref_names[tid]
                if name not in sR2M or name not in s2s:
                    bp.pack_scaffold_reads(tid, {})                            # consume and drop
                    continue
                if name == "FailureScaffoldHeaderTesting" and debug:
                    bp.pack_scaffold_reads(tid, {})
                    yield "failure", name
                    continue
                L = len(s2s[name])
                if full(batch["L"], batch["M"], batch["n_events"], L, _r2m_levels(sR2M[name])):
                    yield "batch", batch
                    batch = _new_batch()
                ev = bp.pack_scaffold_reads(tid, sR2M[name], pos_offset=batch["L"], pair_id_offset=batch["n_pairs"])
                _add_to_batch(batch, name, s2s[name], ev, split_table(name, L))
        if batch["names"]:
            yield "batch", batch
        return

    from .packer import find_bai, pack_scaffolds_parallel, read_bai, scan_scaffold_offsets
    with BamPacker(bam) as bp:
        ref_names = bp.ref_names
    bai = find_bai(bam)
    first = read_bai(bai) if bai is not None else scan_scaffold_offsets(bam)
    # plan: scaffolds in file (= tid) order, batch index and position offset of each
    plan, b_idx, b_L, b_est, b_M = [], 0, 0, 0, 1
    for tid, name in enumerate(ref_names):
        if first[tid] is None or name not in sR2M or name not in s2s:
            continue
        if name == "FailureScaffoldHeaderTesting" and debug:
            plan.append((tid, name, None, None))
            continue
        L = len(s2s[name])
        M_sc = _r2m_levels(sR2M[name])
        if full(b_L, b_M, b_est, L, M_sc):
            b_idx, b_L, b_est, b_M = b_idx + 1, 0, 0, 1
        plan.append((tid, name, b_idx, b_L))
        b_L += L
        b_M = max(b_M, M_sc)
        b_est += 300 * len(sR2M[name])
    jobs = [(tid, sR2M[name], off) for tid, name, bi, off in plan if bi is not None]
    packed = pack_scaffolds_parallel(bam, jobs, packer_threads)
    batch, cur = _new_batch(), 0
    for tid, name, bi, off in plan:
        if bi is None:
            yield "failure", name
            continue
        ev = next(packed)
        if bi != cur:
            if batch["names"]:
                yield "batch", batch
            batch, cur = _new_batch(), bi
        assert off == batch["L"]
        if batch["n_pairs"]:                                                   # pair ids were numbered from 0 per scaffold
            ev["seg_pair"] = ev["seg_pair"] + np.int32(batch["n_pairs"])
            ev["nev_pair"] = ev["nev_pair"] + np.int32(batch["n_pairs"])
        _add_to_batch(batch, name, s2s[name], ev, split_table(name, len(s2s[name])))
    if batch["names"]:
        yield "batch", batch


def profile_scaffolds(bam, sR2M, s2s, Fdb=None, engine=None, device=0, max_batch_events=400_000_000, **kwargs):
    """Profile every scaffold of `sR2M` found in the BAM.  Returns (ProfileResult, engine).

    kwargs (reference names and defaults, argumentParser.py:107-173): min_cov 5, min_freq 0.05, min_snp 20,
    window_length 10000, skip_mm_profiling False (then sR2M values are sets), model_file/fdr for the null model.
    Own keywords: packer_threads (host threads packing scaffolds concurrently; default 1), b200_transfer.
    """
    min_cov = int(kwargs.get("min_cov", 5))
    min_freq = float(kwargs.get("min_freq", 0.05))
    min_snp = int(kwargs.get("min_snp", 20))
    window_length = int(kwargs.get("window_length", 10000))
    fdr = float(kwargs.get("fdr", 1e-6)) or 1e-6                      # 0 -> 1e-6 (controller.py:207-208)
    rarefied_coverage = int(kwargs.get("rarefied_coverage", 50))      # argumentParser.py:168
    seed = int(kwargs.get("seed", 0) or 0)                            # own keyword: key of the re-drawn outputs (clonTR, normalized LD)
    transfer = kwargs.get("b200_transfer") or os.environ.get("ISB_TRANSFER", "segments")
    if transfer not in ("segments", "delta", "cols"):
        raise ValueError("b200_transfer must be 'segments', 'delta' or 'cols'")
    own = engine is None
    if own:
        engine = Engine(device, model_file=kwargs.get("model_file"), fdr=fdr)
    res = ProfileResult()
    t0 = time.time()
    snp_tabs, ld_tabs, sum_tabs = [], [], []
    # keep_rows = {scaffold: global index}: also keep the raw row arrays with scaffold-relative positions, the global
    # scaffold index and the reference character per row -- what a rank sends to rank 0 (profile_bam_distributed)
    store_everything = bool(kwargs.get("store_everything", False))     # --store_everything: keep the raw pileup counts
    keep_rows = kwargs.get("keep_rows")
    res.rows = dict(snv=[], snv_sidx=[], snv_ref=[], ld=[], ld_sidx=[]) if keep_rows is not None else None

    def flush(batch):
        """Profile one batch; a failing batch is logged in the reference's format and skipped, the run continues
        (split_profile_wrapper_groups, profile_utilities.py:92-112)."""
        if not batch["names"]:
            return
        try:
            _flush(batch)
        except Exception as e:                                            # noqa: BLE001 - mirror of the reference's catch-all
            # The reference loses ONE split to an exception (SplitException, profile_utilities.py:104-111).  A batch holds
            # many scaffolds, so a failing batch is bisected until the offending scaffold stands alone: an input outside
            # the device envelope (ISB_ERR_UNSUPPORTED), an out-of-memory condition or a bad scaffold costs that
            # scaffold only.
            n = len(batch["names"])
            if n > 1:
                logging.warning("instrain_b200: batch of %d scaffolds failed (%s); retrying in halves", n, e)
                flush(_sub_batch(batch, 0, n // 2))
                flush(_sub_batch(batch, n // 2, n))
                return
            t = time.strftime("%m-%d %H:%M")
            msg = "\n{1} DEBUG FAILURE SplitException {0} {2}\n".format(batch["names"][0], t, 0)
            logging.error(msg + str(e))
            res.failures.append(batch["names"][0])

    def _flush(batch):
        cat = np.concatenate
        t_batch = time.time()
        pad = batch.get("pad", 0)
        ref_codes = cat(([np.full(pad, 4, np.uint8)] if pad else []) + batch["ref"])
        offs = np.array(batch["off"], dtype=np.int64)
        # host -> device as read-major aligned segments (4 bits per aligned base); K1r transposes on the device.
        # Opt-in (kwargs["b200_transfer"] / ISB_TRANSFER): "delta" sends the reference-delta transfer format (about a
        # quarter of the bytes; K0d rebuilds the stream), "cols" lays the batch out as column words on the host (the
        # packer's transposition) and runs the streaming K1c.  Results are identical.
        rd = reads_mod.concat_streams(batch["parts"])
        fmt = {}
        if transfer == "delta":
            fmt["reads"] = reads_mod.delta_reads_host(rd, ref_codes)
        elif transfer == "cols":
            from .cols import reads_to_cols
            fmt["cols"] = reads_to_cols(rd, len(ref_codes))
        else:
            fmt["reads"] = rd
        t_e = time.time()
        out = engine.profile_batch(dict(pair_mm=cat(batch["pair_mm"])), ref_codes, np.array(batch["splits"], np.int32),
                                   min_cov=min_cov, min_freq=min_freq, min_snp=min_snp,
                                   want=("covT", "clonT", "clonTR", "nmask", "snv", "ld") + (("counts",) if store_everything else ()),
                                   rarefied_coverage=rarefied_coverage, seed=seed, **fmt)
        # merge-stage summary (K4): cumulative_scaffold_table rows of this batch
        bounds = np.append(offs, len(ref_codes)).astype(np.int32)
        if pad:                                                           # the unused leading positions: a dummy segment
            bounds = np.concatenate([[0], bounds]).astype(np.int32)
        k4 = engine.scaffold_summary(out["covT"], out["clonT"], out["nmask"], bounds)
        k4r = engine.scaffold_summary(out["covT"], out["clonTR"], out["nmask"], bounds) if "clonTR" in out else None
        res.timing["engine_s"] = res.timing.get("engine_s", 0.0) + time.time() - t_e      # host<->device copies + kernels
        res.timing["positions"] = res.timing.get("positions", 0) + len(ref_codes)
        res.timing["aligned_bases"] = res.timing.get("aligned_bases", 0) + int(batch["n_events"])
        if pad:
            k4 = k4[out["M"]:]
            k4r = k4r[out["M"]:] if k4r is not None else None
        sum_tabs.append(summary.summary_table(k4, out["snv"], batch["names"], offs, out["M"], rows_rarefied=k4r))
        seqs = {n: s2s[n] for n in batch["names"]}
        snp_tabs.append(tables.snv_table(out["snv"], batch["names"], offs, seqs, ref_codes=ref_codes))
        ld_tabs.append(tables.linkage_table(out["ld"], batch["names"], offs))
        if keep_rows is not None:
            gidx = np.array([keep_rows[n] for n in batch["names"]], dtype=np.int32)
            names_arr = np.asarray(batch["names"], dtype=object)
            r = out["snv"].copy()
            sidx, rel = tables._locate(r["pos"].astype(np.int64), offs)
            ref_ch = tables.snv_ref_chars(r, sidx, rel, names_arr, seqs, ref_codes)
            r["pos"] = rel
            res.rows["snv"].append(r)
            res.rows["snv_sidx"].append(gidx[sidx])
            res.rows["snv_ref"].append(np.array([ord(c) for c in ref_ch], dtype=np.uint8))
            q = out["ld"].copy()
            sidx, rel_a = tables._locate(q["pos_a"].astype(np.int64), offs)
            q["pos_b"] = q["pos_b"].astype(np.int64) - offs[sidx]
            q["pos_a"] = rel_a
            res.rows["ld"].append(q)
            res.rows["ld_sidx"].append(gidx[sidx])
        if logging.getLogger().isEnabledFor(logging.DEBUG):                # the reference's per-split / per-scaffold log lines
            t_end, msg = time.time(), []
            for name, n_sp in zip(batch["names"], batch["n_splits"]):
                for k in range(n_sp):
                    msg.append(worker_log("SplitProfile", "{0}.{1}".format(name, k), "start", t_batch))
                    msg.append(worker_log("SplitProfile", "{0}.{1}".format(name, k), "end", t_end))
                msg.append(worker_log("MergeProfile", name, "start", t_end))
                msg.append(worker_log("MergeProfile", name, "end", t_end))
            logging.debug("".join(msg))
        for name, off in zip(batch["names"], offs):
            sp = ScaffoldProfile(name, len(s2s[name]))
            sl = slice(int(off), int(off) + len(s2s[name]))
            levels = tables.present_levels(out["covT"][sl], out["nmask"][sl])
            sp.covT = tables.basewise(out["covT"][sl], "coverage", levels)
            sp.clonT = tables.basewise(out["clonT"][sl], "clonality", levels)
            if "clonTR" in out:
                sp.clonTR = tables.basewise(out["clonTR"][sl], "clonality", levels)
            if store_everything:                                          # pileup_counts[RelPosition] = total counts (profile_utilities.py:167-168,258-259)
                sp.pileup_counts = out["counts"][sl].sum(axis=1).astype(np.int64)
            res.scaffolds[name] = sp
            res.scaffold_list.append(name)

    for kind, payload in iter_batches(bam, sR2M, s2s, Fdb=Fdb, window_length=window_length,
                                      max_batch_events=max_batch_events, packer_threads=int(kwargs.get("packer_threads", 1) or 1),
                                      debug=kwargs.get("debug", False),
                                      max_batch_cells=int(kwargs.get("max_batch_cells", MAX_BATCH_CELLS))):
        if kind == "failure":
            logging.error("\n{1} DEBUG FAILURE SplitException {0} {2}\n".format(payload, time.strftime("%m-%d %H:%M"), 1))
            res.failures.append(payload)
        else:
            flush(payload)
    res.raw_snp_table = pd.concat(snp_tabs, ignore_index=True) if snp_tabs else pd.DataFrame(columns=tables.SNV_COLUMNS)
    res.raw_linkage_table = pd.concat(ld_tabs, ignore_index=True) if ld_tabs else pd.DataFrame(columns=tables.LD_COLUMNS)
    res.cumulative_snv_table = tables.cumulative_snv_table(res.raw_snp_table)
    res.cumulative_scaffold_table = (pd.concat(sum_tabs, ignore_index=True) if sum_tabs
                                     else pd.DataFrame(columns=summary.COLUMNS))
    empty_snp, empty_ld = res.raw_snp_table.iloc[0:0], res.raw_linkage_table.iloc[0:0]
    by_snp = {k: v for k, v in res.raw_snp_table.groupby("scaffold", sort=False)}       # grouped once, not filtered per scaffold
    by_ld = {k: v for k, v in res.raw_linkage_table.groupby("scaffold", sort=False)}
    for name, sp in res.scaffolds.items():
        sp.raw_snp_table = by_snp.get(name, empty_snp)
        sp.raw_linkage_table = by_ld.get(name, empty_ld)
    res.timing["profile_scaffolds_s"] = time.time() - t0
    logging.debug("instrain_b200: profiled %d scaffolds in %.2fs", len(res.scaffold_list), res.timing["profile_scaffolds_s"])
    if own:
        engine.close()
    return res


def profile_scaffold_run(bam, name, r2m, seq, run_splits, engine, **kwargs):
    """One contiguous run of splits of ONE scaffold (a scaffold too large for a balanced scaffold-wise
    partition is sharded by runs of splits): the scaffold is packed, the reads that overlap the run are cut at its borders
    (instrain_b200.reads.clip_reads) and the run is profiled as a batch of its own with start = its origin, so rows come
    back in scaffold coordinates.  run_splits: [(start, end)] of the run, in order.  Returns dict(snv, ld: row arrays;
    lo, hi; covT, clonT, clonTR, nmask: dense arrays of the positions [lo, hi); M)."""
    from .packer import find_bai, read_bai, read_bai_linear, seek_offset
    lo, hi = int(run_splits[0][0]), int(run_splits[-1][1]) + 1
    with BamPacker(bam) as bp:
        tid = bp.ref_names.index(name)
        bai = find_bai(bam)
        if bai is not None:                                               # straight to the first read that can overlap the run
            first = read_bai(bai)[tid]
            if first is not None:
                bp.seek(seek_offset(read_bai_linear(bai)[tid], first, lo))
        while True:
            t = bp.peek_tid()
            if t < 0 or t >= tid:
                break
            bp.pack_scaffold_reads(t, {})                                  # no index: skip the earlier scaffolds
        # only the reads that overlap the run are read and packed (the read halo); reading stops behind the run
        part = bp.pack_scaffold_reads(tid, r2m, region=(lo, hi)) if bp.peek_tid() == tid else None
    L_run = hi - (lo & ~7)
    if part is None or len(part["seg_start"]) == 0:
        M = _r2m_levels(r2m)
        return dict(snv=np.zeros(0, _cabi_mod().SNV_DT), ld=np.zeros(0, _cabi_mod().LD_DT), lo=lo, hi=hi, M=M,
                    covT=np.zeros((hi - lo, M), np.int32), clonT=np.full((hi - lo, M), np.nan, np.float32),
                    clonTR=np.full((hi - lo, M), np.nan, np.float32), nmask=np.zeros(hi - lo, np.uint64))
    sub, origin = reads_mod.clip_reads(reads_mod.concat_streams([part]), lo, hi)
    ref_codes = encode_reference(seq)[origin:hi]
    out = engine.profile_batch(dict(pair_mm=part["pair_mm"]), ref_codes, np.asarray(run_splits, np.int32), start=origin,
                               min_cov=int(kwargs.get("min_cov", 5)), min_freq=float(kwargs.get("min_freq", 0.05)),
                               min_snp=int(kwargs.get("min_snp", 20)), want=("covT", "clonT", "clonTR", "nmask", "snv", "ld"),
                               rarefied_coverage=int(kwargs.get("rarefied_coverage", 50)), seed=int(kwargs.get("seed", 0) or 0),
                               reads=sub)
    assert len(ref_codes) == L_run
    d = lo - origin
    return dict(snv=out["snv"], ld=out["ld"], lo=lo, hi=hi, M=int(out["M"]), covT=out["covT"][d:], clonT=out["clonT"][d:],
                clonTR=out["clonTR"][d:] if "clonTR" in out else None, nmask=out["nmask"][d:])


def _cabi_mod():
    from . import _cabi
    return _cabi


def profile_bam(bam, Fdb, sR2M, ISP_loc, **kwargs):
    """Drop-in for inStrain.profile.profile_bam (profile/__init__.py:7-18).  Writes the SNVprofile directory at ISP_loc
    (instrain_b200/store.py: attributes.tsv, csv.gz tables, covT / clonT .hd5) and returns the on-disk object the
    reference's ProfileController keeps using as `self.ISP` (`.generate / .get / .store / .get_location`): inStrain's own
    SNVprofile when importable, else instrain_b200.store.ProfileStore.  The in-memory tables of the run are its `.result`
    (a ProfileResult; its attributes are also reachable directly on a ProfileStore).  With store=False (or ISP_loc None)
    the bare ProfileResult is returned."""
    s2s = kwargs.pop("s2s", None)
    report = None
    t_start = time.time()
    if s2s is None:
        raise ValueError("profile_bam needs kwargs['s2s'] (scaffold -> sequence), as ProfileController.run_profile passes it")
    if sR2M is None:
        # no Rdic from the caller: run the reference's default read filter ourselves (C++ host filter, one BAM pass)
        from .packer import BamPacker as _BP
        from .read_filter import filter_reads
        with _BP(bam) as bp:
            names = bp.ref_names
        fkw = {k: kwargs[k] for k in ("min_read_ani", "min_mapq", "max_insert_relative", "min_insert", "pairing_filter")
               if k in kwargs}
        pr = kwargs.get("priority_reads")
        if pr is not None and not isinstance(pr, (set, list, tuple)):      # the CLI hands a file name (argumentParser.py:95-97)
            from .read_filter import load_priority_reads
            pr = load_priority_reads(pr)
        sR2M, _, _, report = filter_reads(bam, names, priority_reads=pr or (), with_report=True,      # one pass: sR2M + mapping_info
                                          threads=int(kwargs.get("packer_threads", 1) or 1), **fkw)
        if kwargs.get("skip_mm_profiling"):
            sR2M = {s: set(d) for s, d in sR2M.items()}
    t_filter = time.time() - t_start
    log_checkpoint("Profile", "B200_profile_scaffolds", "start")
    res = profile_scaffolds(bam, sR2M, s2s, Fdb=Fdb, **kwargs)
    log_checkpoint("Profile", "B200_profile_scaffolds", "end")
    res.timing["read_filter_s"] = t_filter
    # the SNVprofile directory at ISP_loc (gen_snv_profile, profile_utilities.py:670-706), written natively:
    # inStrain.SNVprofile.SNVprofile(ISP_loc) of the reference opens it unchanged
    if ISP_loc is None or not kwargs.get("store", True):
        return res
    from .store import store_profile
    fdef = dict(min_read_ani=0.95, min_mapq=-1, max_insert_relative=3, min_insert=50)    # this shim's filter defaults
    t_store = time.time()
    log_checkpoint("Profile", "B200_store", "start")
    S = store_profile(ISP_loc, bam, res, mapping_info=report, **{k: kwargs.get(k, v) for k, v in fdef.items()})
    log_checkpoint("Profile", "B200_store", "end")
    res.timing["store_s"] = time.time() - t_store
    res.timing["total_s"] = time.time() - t_start
    res.store = S
    return _as_snvprofile(S, res)


def _as_snvprofile(S, res):
    """The object handed back to ProfileController.run_profile (controller.py:341-350), which goes on to call
    `.generate(...)`, `.get(...)`, `.store(...)` and `.get_location(...)` on it: the reference's own
    inStrain.SNVprofile.SNVprofile opened on the directory just written when that class can be imported (it needs h5py),
    otherwise this package's ProfileStore (same methods, same directory).  Either way the in-memory ProfileResult rides
    along as `.result`."""
    if os.environ.get("ISB_NATIVE_STORE", "0") != "1":
        try:
            import h5py  # noqa: F401 - the reference's SNVprofile cannot load covT / clonT without it
            from inStrain.SNVprofile import SNVprofile
            isp = SNVprofile(S.location)
            isp.result = res
            return isp
        except Exception:                                                # noqa: BLE001 - any import problem: native store
            pass
    return S


def profile_bam_distributed(bam, Fdb, sR2M, ISP_loc, **kwargs):
    """profile_bam over several GPUs of one node: one process per GPU (torchrun; torch.distributed initialised by the
    caller, NCCL on GPUs / gloo in the CPU tests).  The reference farms (scaffold, split) tasks of ONE input to worker
    processes, heaviest scaffolds first (profile_controller.py:243-271, fasta.py:103-105); here the scaffolds of the one
    BAM are partitioned over the ranks by longest-processing-time packing on their filtered pairs
    (instrain_b200.shard.lpt_partition), every rank profiles its share on its own GPU -- no collective on the data path --
    and the final tables travel to rank 0 (instrain_b200.shard.gather_rows for the row tables; the per-scaffold coverage /
    clonality series as pickled objects).  Rank 0 writes the SNVprofile directory and returns what profile_bam returns;
    the other ranks return None."""
    import torch.distributed as dist
    from . import _cabi
    from .shard import gather_rows, lpt_partition
    rank, world = dist.get_rank(), dist.get_world_size()
    s2s = kwargs.get("s2s")
    if s2s is None or sR2M is None:
        raise ValueError("profile_bam_distributed needs sR2M and kwargs['s2s'] (run the read filter once, on rank 0 or on every rank)")
    order = list(sR2M)                                                    # the same on every rank
    weights = [float(len(sR2M[s])) for s in order]
    # A scaffold heavier than a rank's fair share cannot be balanced scaffold-wise: it is cut into contiguous runs of its
    # splits (one unit each, weight by length), profiled from the reads that overlap the run (profile_scaffold_run) and
    # put back together on rank 0.  kwargs["split_runs"] = False keeps every scaffold whole.
    units, heavy = [], {}                                                 # unit = (scaffold index, run index or None, weight)
    fair = sum(weights) / world if world > 1 else float("inf")
    st = _SplitTable(Fdb, int(kwargs.get("window_length", 10000)))
    for i, name in enumerate(order):
        k = int(np.ceil(weights[i] / fair)) if (kwargs.get("split_runs", True) and weights[i] > 1.25 * fair) else 1
        sp = st(name, len(s2s[name])) if k > 1 and name in s2s else None
        if sp is not None and len(sp) > 1:
            from .shard import split_runs
            runs = split_runs(sp, [e - a + 1 for a, e in sp], min(k, world))
            heavy[i] = (sp, runs)
            tot_len = float(sum(e - a + 1 for a, e in sp))
            for r_, (a_, b_) in enumerate(runs):
                units.append((i, r_, weights[i] * sum(e - a + 1 for a, e in sp[a_:b_]) / tot_len))
        else:
            units.append((i, None, weights[i]))
    mine_u = [units[j] for j in lpt_partition([u[2] for u in units], world)[rank]]
    mine = [u[0] for u in mine_u if u[1] is None]
    my_runs = [(u[0], u[1]) for u in mine_u if u[1] is not None]
    sub = {order[i]: sR2M[order[i]] for i in mine}
    kw = dict(kwargs)
    kw.pop("s2s")
    kw["keep_rows"] = {s: i for i, s in enumerate(order)}
    kw.setdefault("packer_threads", 2)                                    # > 1: seek through the .bai instead of reading the whole BAM
    device = kw.pop("device", None)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    gdev = None
    if dist.get_backend() == "nccl":
        import torch
        gdev = torch.device("cuda", device)
    run_engine = None
    if my_runs or (rank == 0 and heavy):                                  # runs, and their merge on rank 0, need an engine of their own
        run_engine = Engine(device, model_file=kwargs.get("model_file"), fdr=float(kwargs.get("fdr", 1e-6)) or 1e-6)
    res = profile_scaffolds(bam, sub, s2s, Fdb=Fdb, device=device, **kw)
    run_parts = []
    for i, r_ in my_runs:
        sp, runs = heavy[i]
        a_, b_ = runs[r_]
        try:
            rr = profile_scaffold_run(bam, order[i], sR2M[order[i]], s2s[order[i]], sp[a_:b_], run_engine, **kwargs)
        except Exception as e:                                           # noqa: BLE001 - the reference's per-split catch-all
            logging.error("\n{1} DEBUG FAILURE SplitException {0} {2}\n".format(order[i], time.strftime("%m-%d %H:%M"), r_) + str(e))
            res.failures.append(order[i])
            continue
        seq = s2s[order[i]]
        res.rows["snv"].append(rr["snv"])
        res.rows["snv_sidx"].append(np.full(len(rr["snv"]), i, np.int32))
        res.rows["snv_ref"].append(np.array([ord(seq[int(p_)]) for p_ in rr["snv"]["pos"]], dtype=np.uint8))
        res.rows["ld"].append(rr["ld"])
        res.rows["ld_sidx"].append(np.full(len(rr["ld"]), i, np.int32))
        run_parts.append(dict(sidx=i, run=r_, lo=rr["lo"], hi=rr["hi"], M=rr["M"], covT=rr["covT"], clonT=rr["clonT"],
                              clonTR=rr["clonTR"], nmask=rr["nmask"]))
    cat = lambda parts, dt: np.concatenate(parts) if parts else np.zeros(0, dtype=dt)
    snv = gather_rows(cat(res.rows["snv"], _cabi.SNV_DT), device=gdev)
    snv_sidx = gather_rows(cat(res.rows["snv_sidx"], np.int32), device=gdev)
    snv_ref = gather_rows(cat(res.rows["snv_ref"], np.uint8), device=gdev)
    ld = gather_rows(cat(res.rows["ld"], _cabi.LD_DT), device=gdev)
    ld_sidx = gather_rows(cat(res.rows["ld_sidx"], np.int32), device=gdev)
    small = dict(scaffold_list=res.scaffold_list, scaffolds=res.scaffolds, failures=res.failures,
                 summary=res.cumulative_scaffold_table, seconds=res.timing.get("profile_scaffolds_s"), runs=run_parts)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(small, parts, dst=0)
    if rank != 0:
        if run_engine is not None:
            run_engine.close()
        return None
    out = ProfileResult()
    names = np.asarray(order, dtype=object)
    o = np.lexsort((snv["mm"], snv["pos"], snv_sidx))
    snv, snv_sidx, snv_ref = snv[o], snv_sidx[o], snv_ref[o]
    out.raw_snp_table = tables.snv_frame(snv, names[snv_sidx], snv["pos"].astype(np.int64),
                                         np.array([chr(c) for c in snv_ref], dtype=object))
    o = np.lexsort((ld["mm"], ld["pos_b"], ld["pos_a"], ld_sidx))
    ld, ld_sidx = ld[o], ld_sidx[o]
    out.raw_linkage_table = tables.linkage_frame(ld, names[ld_sidx], ld["pos_a"].astype(np.int64), ld["pos_b"].astype(np.int64))
    out.cumulative_snv_table = tables.cumulative_snv_table(out.raw_snp_table)
    sums = [p["summary"] for p in parts if p["summary"] is not None and len(p["summary"])]
    out.cumulative_scaffold_table = pd.concat(sums, ignore_index=True) if sums else pd.DataFrame(columns=summary.COLUMNS)
    for r, p in enumerate(parts):
        out.scaffold_list.extend(p["scaffold_list"])
        out.scaffolds.update(p["scaffolds"])
        out.failures.extend(p["failures"])
        out.timing["rank%d_profile_scaffolds_s" % r] = p["seconds"]
    # the scaffolds that were profiled by runs: dense series put back together, basewise tables and the merge-stage
    # summary (K4) computed on rank 0
    for i, (sp, runs) in sorted(heavy.items()):
        name = order[i]
        pieces = sorted((q for p in parts for q in p["runs"] if q["sidx"] == i), key=lambda q: q["lo"])
        if len(pieces) != len(runs) or name in out.failures:
            if name not in out.failures:
                out.failures.append(name)
            continue
        Ls, Ms = len(s2s[name]), max(q["M"] for q in pieces)
        covT = np.zeros((Ls, Ms), np.int32)
        clonT = np.full((Ls, Ms), np.nan, np.float32)
        clonTR = np.full((Ls, Ms), np.nan, np.float32)
        nmask = np.zeros(Ls, np.uint64)
        for q in pieces:
            sl = slice(q["lo"], q["hi"])
            covT[sl, :q["M"]] = q["covT"]
            clonT[sl, :q["M"]] = q["clonT"]
            if q["clonTR"] is not None:
                clonTR[sl, :q["M"]] = q["clonTR"]
            nmask[sl] = q["nmask"]
        bounds = np.array([0, Ls], dtype=np.int32)
        k4 = run_engine.scaffold_summary(covT, clonT, nmask, bounds)
        k4r = run_engine.scaffold_summary(covT, clonTR, nmask, bounds)
        sel = snv[snv_sidx == i]
        sums.append(summary.summary_table(k4, sel, [name], np.array([0]), Ms, rows_rarefied=k4r))
        spf = ScaffoldProfile(name, Ls)
        levels = tables.present_levels(covT, nmask)
        spf.covT = tables.basewise(covT, "coverage", levels)
        spf.clonT = tables.basewise(clonT, "clonality", levels)
        spf.clonTR = tables.basewise(clonTR, "clonality", levels)
        out.scaffolds[name] = spf
        out.scaffold_list.append(name)
    if run_engine is not None:
        run_engine.close()
    out.cumulative_scaffold_table = pd.concat(sums, ignore_index=True) if sums else pd.DataFrame(columns=summary.COLUMNS)
    by_snp = {k: v for k, v in out.raw_snp_table.groupby("scaffold", sort=False)}
    by_ld = {k: v for k, v in out.raw_linkage_table.groupby("scaffold", sort=False)}
    for name, sp in out.scaffolds.items():
        sp.raw_snp_table = by_snp.get(name, out.raw_snp_table.iloc[0:0])
        sp.raw_linkage_table = by_ld.get(name, out.raw_linkage_table.iloc[0:0])
    if ISP_loc is None or not kwargs.get("store", True):
        return out
    from .store import store_profile
    S = store_profile(ISP_loc, bam, out)
    out.store = S
    return _as_snvprofile(S, out)
