"""Aggregate an .ncu-rep source page by line ranges: python tools/ncu_phases.py REPORT KERNEL_REGEX FILE_PREFIX name:lo-hi ..."""
import csv, io, subprocess, sys, collections

def main():
    rep, kre, fpre = sys.argv[1:4]
    ranges = []
    for spec in sys.argv[4:]:
        name, r = spec.rsplit(":", 1)
        lo, hi = r.split("-")
        ranges.append((name, int(lo), int(hi)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                          "regex:" + kre], capture_output=True, text=True).stdout
    hdr, fname = None, ""
    samp, inst = collections.Counter(), collections.Counter()
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr is None or not r[0].strip().isdigit():
            continue
        ln = int(r[0])
        key = "(inlined from other files)"
        if fname.startswith(fpre):
            key = next((n for n, lo, hi in ranges if lo <= ln <= hi), "(other lines)")
        def f(name):
            try:
                return float(r[ix[name]])
            except (ValueError, KeyError):
                return 0.0
        samp[key] += f("# Samples")
        inst[key] += f("Instructions Executed")
    ts, ti = sum(samp.values()), sum(inst.values())
    for k in list(dict.fromkeys([n for n, _, _ in ranges] + sorted(samp))):
        print(f"{k:34s} samples {100 * samp[k] / ts:5.1f}%   inst {100 * inst[k] / ti:5.1f}%")

main()
