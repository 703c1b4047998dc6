"""Per-CUDA-line instruction and stall-sample shares from an ncu report (captured with --import-source on).

    ncu -i report.ncu-rep --page source --print-source cuda,sass --csv | python tools/ncu_lines.py [top]
"""
import csv
import sys


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rows = list(csv.reader(sys.stdin))
    path, hdr, out = "", None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            path = r[1].rsplit('/', 1)[-1]
        elif "Line No" in r[:1] and "Instructions Executed" in r:
            hdr = r
        elif hdr is not None and r and r[0].strip().isdigit():
            try:
                out.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), path, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot, ts = sum(o[0] for o in out), max(sum(o[1] for o in out), 1)
    print("warp instructions %d, stall samples %d" % (tot, ts))
    for n, s, p, ln, src in sorted(out, reverse=True)[:top]:
        print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100.0 * n / tot, 100.0 * s / ts, p, ln, src[:110]))


if __name__ == "__main__":
    main()
