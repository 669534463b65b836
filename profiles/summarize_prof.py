"""Summarise bench.py --prof-dump output: per-shape launches, ms, TF/s (GEMM) or GB/s (prep), share.
usage: python profiles/summarize_prof.py gpurun_out/x/prof.tsv [step_ms] [steps]"""
import sys


def main(path, step_ms=None, steps=1):
    rows = []
    for line in open(path):
        cat, tag, n, ms, val = line.rstrip("\n").split("\t")
        rows.append((int(cat), tag, int(n), float(ms), float(val)))
    tot = {}
    for c, _, n, ms, v in rows:
        t = tot.setdefault(c, [0, 0.0, 0.0])
        t[0] += n; t[1] += ms; t[2] += v
    names = {0: "conv fwd/dgrad GEMM", 1: "wgrad GEMM", 2: "operand-plane prep", 3: "memory-bound helpers (per C-ABI entry point)"}
    for c, (n, ms, v) in sorted(tot.items()):
        unit = "GB/s" if c >= 2 else "TF/s"
        rate = v / ms / 1e6 if c >= 2 else v / ms / 1e9
        share = f", {100 * ms / steps / step_ms:.1f} % of step" if step_ms else ""
        print(f"## {names[c]}: {n // steps} launches/step, {ms / steps:.1f} ms/step, {rate:.1f} {unit}{share}\n")
        print("| shape | launches/step | ms/step | avg us | rate | share of category |")
        print("|---|---:|---:|---:|---:|---:|")
        for cc, tag, nn, mms, vv in sorted((r for r in rows if r[0] == c), key=lambda r: -r[3])[:40]:
            r = vv / mms / 1e6 if c >= 2 else vv / mms / 1e9
            print(f"| {tag} | {nn // steps} | {mms / steps:.2f} | {1e3 * mms / nn:.1f} | {r:.1f} {unit} | {100 * mms / ms:.1f} % |")
        print()


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 1)
