"""Per-source-line summary of an ncu report's source page (helper; run where ncu is installed):
   python profiles/srcprof.py gpurun_out/x.ncu-rep [kernel-regex] [top]
Aggregates, per CUDA source line, the executed warp instructions, stall samples and shared-memory
wavefronts of the SASS instructions attributed to it (needs -lineinfo)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # sections: kernel header, then column header starting with "Line No"
    sect, cur, kern, fname = [], None, None, '?'
    for r in rows:
        if r and r[0] == "Kernel Name":
            kern = r[1]
        elif r and r[0] == "File Name":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            cur = {"kernel": kern, "file": fname, "hdr": r, "rows": []}
            sect.append(cur)
        elif cur is not None and r and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    agg = defaultdict(lambda: [0, 0, 0, 0, 0, ""])
    tot_inst = tot_samp = 0
    for s in sect:
        h = s["hdr"]
        ix = {name: i for i, name in reversed(list(enumerate(h)))}
        line, src = None, ""
        for r in s["rows"]:
            if r[0]:
                line, src = (s["file"], int(r[0])), r[1]
            a = agg[line]
            a[5] = src
            def f(name):
                try:
                    return int(float(r[ix[name]] or 0))
                except (ValueError, KeyError):
                    return 0
            a[0] += f("Instructions Executed")
            a[1] += f("# Samples")
            a[2] += f("L1 Wavefronts Shared")
            a[3] += f("stall_long_sb") + f("stall_lg")
            a[4] += f("stall_barrier") + f("stall_short_sb") + f("stall_mio")
            tot_inst += f("Instructions Executed")
            tot_samp += f("# Samples")
    print(f"{tot_inst} warp instructions, {tot_samp} samples")
    print(f"{'file:line':>20} {'inst%':>6} {'samp%':>6} {'smem wf':>9} {'lsb+lg':>7} {'bar+ssb+mio':>11}  source")
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{line[0][:14]+':'+str(line[1]):>20} {100*a[0]/max(1,tot_inst):6.2f} {100*a[1]/max(1,tot_samp):6.2f} {a[2]:>9} {a[3]:>7} {a[4]:>11}  {a[5].strip()[:100]}")


if __name__ == "__main__":
    main()
