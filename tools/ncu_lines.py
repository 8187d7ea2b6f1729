"""Per-CUDA-source-line instruction / stall-sample shares of one kernel of a `--set full --import-source on` report.

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [min_pct]
"""
import csv
import subprocess
import sys


def main(rep, kern, min_pct=0.4):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname, ix, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ix = {k: j for j, k in enumerate(r)}
        elif ix and r[0] not in ("", "Function Name", "Kernel Name") and r[0].isdigit():
            try:
                lines.append((fname, int(r[0]), r[1].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])))
            except (ValueError, IndexError):
                pass
    ti = sum(l[3] for l in lines) or 1
    ts = sum(l[4] for l in lines) or 1
    print("total warp instructions %d, samples %d" % (ti, ts))
    for f, n, src, i, s in lines:
        if 100.0 * i / ti >= min_pct or 100.0 * s / ts >= min_pct:
            print("%-18s %4d  inst %5.1f%%  samp %5.1f%%  %s" % (f[:18], n, 100.0 * i / ti, 100.0 * s / ts, src[:110]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.4)


def sections(rep, kern, spec):
    """spec: 'name:file:lo-hi,file:lo-hi;name2:...' -> instruction / sample share per named group of line ranges."""
    import io, contextlib
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname, ix, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ix = {k: j for j, k in enumerate(r)}
        elif ix and r[0].isdigit():
            try:
                lines.append((fname, int(r[0]), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])))
            except (ValueError, IndexError):
                pass
    ti = sum(l[2] for l in lines) or 1
    ts = sum(l[3] for l in lines) or 1
    used = set()
    for grp in spec.split(";"):
        name, rest = grp.split(":", 1)
        i = s = 0
        for rng in rest.split(","):
            f, lh = rng.split(":")
            lo, hi = (int(v) for v in lh.split("-"))
            for k, l in enumerate(lines):
                if l[0].startswith(f) and lo <= l[1] <= hi and k not in used:
                    used.add(k); i += l[2]; s += l[3]
        print("%-14s inst %5.1f%%  samp %5.1f%%" % (name, 100.0 * i / ti, 100.0 * s / ts))
    i = sum(l[2] for k, l in enumerate(lines) if k not in used); s = sum(l[3] for k, l in enumerate(lines) if k not in used)
    print("%-14s inst %5.1f%%  samp %5.1f%%" % ("(other)", 100.0 * i / ti, 100.0 * s / ts))
    rest = sorted(((l[2], l) for k, l in enumerate(lines) if k not in used), reverse=True)[:12]
    for _, l in rest:
        print("    other: %s:%d inst %.1f%% samp %.1f%%" % (l[0], l[1], 100.0 * l[2] / ti, 100.0 * l[3] / ts))
