"""Per-source-line summary of one kernel in an .ncu-rep: instructions, stall samples, shared wavefronts, L2 sectors.

    python tools/ncu_lines.py REPORT [kernel-regex] [top]
"""
import csv, io, subprocess, sys, collections

def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]
    if kre:
        cmd += ["--kernel-name", "regex:" + kre]
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    agg = collections.OrderedDict()
    tot = collections.Counter()
    hdr, ix, fname = None, None, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr is None or len(r) < len(hdr) - 3 or not r[0].strip().isdigit():
            continue  # SASS rows: the source-line rows already carry the per-line totals
        key = (fname, int(r[0]), r[1].strip())
        a = agg.setdefault(key, collections.Counter())
        for k, name in (("inst", "Instructions Executed"), ("samp", "# Samples"), ("wf", "L1 Wavefronts Shared"),
                        ("sect", "L2 Theoretical Sectors Global"), ("thr", "Thread Instructions Executed")):
            try:
                v = float(r[ix[name]] or 0)
            except (ValueError, KeyError):
                v = 0.0
            a[k] += v
            tot[k] += v
    print(f"total inst {tot['inst']:.3g} samples {tot['samp']:.0f} smem wavefronts {tot['wf']:.3g} L2 sectors {tot['sect']:.3g}"
          f" avg threads {tot['thr'] / max(tot['inst'], 1):.1f}")
    items = sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]
    for (fn, ln, src), a in sorted(items):
        print(f"{fn[:14]:14s}{ln:5d} inst {100 * a['inst'] / tot['inst']:5.1f}% samp {100 * a['samp'] / max(tot['samp'], 1):5.1f}% "
              f"wf {100 * a['wf'] / max(tot['wf'], 1):5.1f}% sect {100 * a['sect'] / max(tot['sect'], 1):5.1f}%  {src[:100]}")

main()
