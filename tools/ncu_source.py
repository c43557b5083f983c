"""Hottest source lines of a kernel in an .ncu-rep captured with --import-source on: warp stall samples per CUDA source line
(cuda,sass view), with the dominant stall reasons.     python tools/ncu_source.py rep.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    lines = []
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0] not in ("", "Function Name") and r[0].isdigit():
            lines.append((fname, r))
    si = hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[si]) for _, r in lines if r[si].isdigit())
    print(f"# {rep}: {total} warp stall samples; top {top} source lines")
    lines.sort(key=lambda fr: -int(fr[1][si]) if fr[1][si].isdigit() else 0)
    for fname, r in lines[:top]:
        n = int(r[si])
        stalls = sorted(((int(r[i]) if r[i].isdigit() else 0, h[6:]) for i, h in stall_cols), reverse=True)[:3]
        why = ", ".join(f"{h} {v}" for v, h in stalls if v)
        print(f"{100.0 * n / max(total, 1):5.1f}%  {fname}:{r[0]:>5s}  {r[1].strip()[:110]}   [{why}]")


if __name__ == "__main__":
    main()
