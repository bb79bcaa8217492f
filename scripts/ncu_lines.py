"""Per-CUDA-source-line instruction counts from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    cur_file = None
    out = []
    total = 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 7 and r[0].isdigit():
            try:
                n = int(r[7]); s = int(r[4])
            except ValueError:
                continue
            out.append((n, s, cur_file, int(r[0]), r[1].strip()[:110]))
            total += n
    out.sort(reverse=True)
    print("total warp instructions attributed:", total)
    for n, s, f, ln, src in out[:top]:
        print(f"{100*n/total:5.1f}% {n:>11d} smp {s:>6d} {f}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
