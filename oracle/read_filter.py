"""Restatement of the reference's read filter (test infrastructure only).

    get_paired_reads          inStrain/filter_reads.py:885-956   (pair2info from the BAM, per scaffold)
    paired_read_filter        inStrain/filter_reads.py:471-532   (paired_only / non_discordant / all_reads, priority reads)
    filter_scaff2pair2info    inStrain/filter_reads.py:201-300   (median insert, thresholds, tallies)
    evaluate_pair             inStrain/filter_reads.py:387-426
Output: sR2M (scaffold -> read-pair name -> summed NM), the hot path's input.  Pinned on the stored Rdic.json
(tests/test_oracle_golden.py, build container) and used as the twin of the C++ filter in instrain_b200/csrc/isb_host.cpp.
"""
import numpy as np

_MATCH = (0, 7, 8)


def _aligned_span(read):
    """(first, last) reference position of the aligned (M/=/X) bases, i.e. get_reference_positions()[0] / [-1]."""
    pos, first, last = read.pos, None, None
    for op, n in read.cigar:
        if op in _MATCH:
            if first is None:
                first = pos
            last = pos + n - 1
            pos += n
        elif op in (2, 3):
            pos += n
    return first, last


def _query_length(read):
    """pysam infer_query_length(): M, I, S, =, X (hard clips excluded)."""
    return sum(n for op, n in read.cigar if op in (0, 1, 4, 7, 8))


def pair2info(reads):
    """get_paired_reads for ONE scaffold: reads = its records in file order.  name -> [nm, insert, mapq, length, reads]."""
    info, span = {}, {}
    for r in reads:
        if r.flag & 0x4 or not r.cigar:
            continue
        first, last = _aligned_span(r)
        if first is None:
            continue
        nm = int(r.nm or 0)
        if r.name not in info:
            info[r.name] = [nm, -1, r.mapq, _query_length(r), 1]
            span[r.name] = (first, last)
        else:
            i = info[r.name]
            i[0] += nm
            i[4] += 1
            i[3] += _query_length(r)
            i[2] = max(i[2], r.mapq)
            if i[4] == 2:
                s0, e0 = span[r.name]
                i[1] = last - s0 if last > s0 else e0 - first
            else:
                i[1] = -1
            span[r.name] = (0, 0)
    return info


def _merge_info(i1, i2):
    """_merge_info (filter_reads.py:534-542): a pair whose mates map to two scaffolds, under pairing_filter='all_reads'.
    `max([a + b])` of the reference is the sum; insert becomes -2 (which then FAILS min_insert for the merged pair)."""
    return [i1[0] + i2[0], -2, i1[2] + i2[2], i1[3] + i2[3], i1[4] + i2[4]]


def pairing_filter_pairs(scaff2info, pairing_filter="paired_only", priority_reads=()):
    """paired_read_filter (filter_reads.py:471-532): which names go on to the thresholds, per scaffold (dict order kept).
    Also returns the `unfiltered_*` tallies it makes."""
    priority = set(priority_reads)
    out, where, tallies = {}, {}, {}
    for s, d in scaff2info.items():
        out[s] = {}
        t = dict(unfiltered_reads=0, unfiltered_pairs=0, unfiltered_singletons=0, unfiltered_priority_reads=0)
        for p, i in d.items():
            t["unfiltered_reads"] += i[4]
            t["unfiltered_pairs"] += i[4] == 2
            t["unfiltered_singletons"] += i[4] == 1
            t["unfiltered_priority_reads"] += p in priority
            if pairing_filter == "paired_only":
                if i[4] == 2 or p in priority:
                    out[s][p] = i
            elif pairing_filter == "non_discordant":
                if p not in where or p in priority:
                    out[s][p] = i
                    where[p] = s
                else:                                  # seen on another scaffold before: discordant, drop that one too
                    del out[where[p]][p]
            elif pairing_filter == "all_reads":
                if p in where:
                    mi = _merge_info(i, out[where[p]][p])
                    out[s][p] = mi
                    out[where[p]][p] = mi
                else:
                    where[p] = s
                    out[s][p] = i
            else:
                raise ValueError("unknown pairing_filter %r" % pairing_filter)
        tallies[s] = t
    return out, tallies


def filter_pairs(scaff2info, min_read_ani=0.95, min_mapq=-1, max_insert_relative=3, min_insert=50,
                 pairing_filter="paired_only", priority_reads=()):
    """paired_read_filter + filter_scaff2pair2info.  Returns (sR2M, tallies per scaffold, max_insert)."""
    priority = set(priority_reads)
    paired, pre = pairing_filter_pairs(scaff2info, pairing_filter, priority)
    inserts = [i[1] for d in paired.values() for i in d.values() if i[4] == 2]
    max_insert = float(np.median(inserts)) * max_insert_relative if inserts else float("nan")
    out, tallies = {}, {}
    for s, d in paired.items():
        t = dict(pre[s], pass_pairing_filter=0, pass_min_read_ani=0, pass_max_insert=0, pass_min_insert=0, pass_min_mapq=0,
                 filtered_pairs=0, filtered_singletons=0, filtered_priority_reads=0)
        out[s] = {}
        for p, i in d.items():
            t["pass_pairing_filter"] += 1
            f_ani = (1 - float(i[0]) / float(i[3])) > min_read_ani
            f_mapq = i[2] > min_mapq
            if i[4] == 2 and i[1] != -1:
                f_min, f_max = i[1] > min_insert, i[1] < max_insert
            else:
                f_min = f_max = True
            t["pass_min_read_ani"] += f_ani
            t["pass_max_insert"] += f_max
            t["pass_min_insert"] += f_min
            t["pass_min_mapq"] += f_mapq
            if f_ani and f_mapq and f_min and f_max:
                t["filtered_pairs"] += 1
                out[s][p] = i[0]
                t["filtered_singletons"] += i[4] == 1
                t["filtered_priority_reads"] += p in priority
        tallies[s] = t
    return out, tallies, max_insert
