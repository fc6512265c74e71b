"""Hot source lines of one kernel from an ncu report (needs -lineinfo and --import-source on):
  python tools/ncu_source_hot.py gpurun_out/x.ncu-rep regex:kernel_name [top_n] [launch_skip]
Aggregates warp-stall samples per CUDA source line (file:line) and prints the dominant stall reason."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    cur_file = None
    hdr = None
    agg = defaultdict(lambda: [0, defaultdict(int), ""])
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("Function Name",) or len(r) < len(hdr) or r[2] != "-":
            continue  # only the per-source-line summary rows (Address == "-")
        try:
            line = int(r[0])
            samples = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        a = agg[(cur_file, line)]
        a[0] += samples
        a[2] = r[1].strip()[:100]
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    a[1][h] += int(r[i])
                except ValueError:
                    pass
    total = sum(a[0] for a in agg.values()) or 1
    print(f"{kern}: {total} samples")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        reasons = sorted(a[1].items(), key=lambda kv: -kv[1])[:2]
        rs = ", ".join(f"{k[6:]}={v}" for k, v in reasons if v)
        print(f"{100 * a[0] / total:5.1f}%  {f}:{l:<5d} [{rs}]  {a[2]}")


if __name__ == "__main__":
    main()
