"""Basic-block weight profile of a kernel from an ncu report's source page (SASS view): executed warp
instructions per block, stall samples per block, and the stall reasons of the whole kernel.
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_blocks.py src.csv [kernel-index]"""
import csv
import sys


def sections(path):
    rows = list(csv.reader(open(path)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    return secs


def main():
    secs = sections(sys.argv[1])
    which = [int(sys.argv[2])] if len(sys.argv) > 2 else range(len(secs))
    nodes = float(sys.argv[3]) if len(sys.argv) > 3 else 8193 * 1025
    for w in which:
        s = secs[w]
        h = s["hdr"]
        iI, iS, isrc, ib = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source"), h.index("stall_barrier")
        tot = sum(int(r[iI]) for r in s["rows"])
        print(f"== {s['name']}  warp insts {tot}  thread-instr/node {tot * 32 / nodes:.1f}")
        reasons = {}
        for r in s["rows"]:
            for j in range(ib, ib + 18):
                if r[j] not in ("", "0") and "Not Issued" not in h[j]:
                    reasons[h[j]] = reasons.get(h[j], 0) + int(r[j])
        print("  ", sorted(reasons.items(), key=lambda kv: -kv[1]))
        blocks, start, prev = [], 0, None
        for k, r in enumerate(s["rows"]):
            c = int(r[iI])
            if prev is not None and c != prev:
                blocks.append((start, k - 1, prev))
                start = k
            prev = c
        blocks.append((start, len(s["rows"]) - 1, prev))
        for a, b, c in blocks:
            wgt = (b - a + 1) * c
            if wgt > 0.004 * tot:
                smp = sum(int(s["rows"][k][iS]) for k in range(a, b + 1))
                print(f"   {a:5d}-{b:5d} n={b - a + 1:4d} count={c:9d} share={100 * wgt / tot:5.1f}% samples={smp:6d}  "
                      f"{s['rows'][a][isrc].strip()[:44]}")


if __name__ == "__main__":
    main()
