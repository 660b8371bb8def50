"""Heaviest source lines of one kernel from an ncu report's source page in the `cuda,sass` view:
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> --launch-count 1 > v.csv
    python tools/ncu_lines.py v.csv [top-n] [nodes]
Prints executed warp instructions per source line (share of the kernel, thread instructions per mesh node)."""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    nodes = float(sys.argv[3]) if len(sys.argv) > 3 else 8193 * 1025
    cur, lines, infunc = None, [], False
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            infunc = True
        elif r[0] == "Line No":
            hdr = r
            iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif infunc and r[0] not in ("", "Line No") and len(r) > iI and r[iI].isdigit():
            lines.append((int(r[iI]), int(r[iS]) if r[iS].isdigit() else 0, cur, r[0], r[1].strip()))
    tot = sum(x[0] for x in lines)
    print(f"warp instructions {tot}  thread-instr/node {tot * 32 / nodes:.1f}")
    for c, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"  {100 * c / tot:5.1f}%  {c * 32 / nodes:6.1f}/node  smp {smp:5d}  {f}:{ln}  {src[:90]}")


if __name__ == "__main__":
    main()
