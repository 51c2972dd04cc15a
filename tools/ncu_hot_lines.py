"""Hottest CUDA source lines of one kernel of a `ncu --set full --import-source on` report (warp-stall samples and executed
warp instructions per line).

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass -k regex:NAME > page.csv; python tools/ncu_hot_lines.py page.csv [N]
"""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    cur, out = None, []
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0].strip().isdigit() and len(r) > 7:
            try:
                out.append((int(r[4]), int(r[7]), cur, int(r[0]), r[1].strip()[:96]))
            except ValueError:
                pass
    ts, ti = sum(o[0] for o in out), sum(o[1] for o in out)
    print(f"samples {ts}  warp instructions {ti}")
    for s, i, f, l, src in sorted(out, reverse=True)[:top]:
        print(f"{100 * s / ts:5.1f}% samples {100 * i / ti:5.1f}% inst  {f}:{l}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
