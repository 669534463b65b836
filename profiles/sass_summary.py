"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
UTMALDG = TMA loads, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, plus the legacy HMMA / FFMA for contrast.
usage: python profiles/sass_summary.py [path/to/libdvdgan_b200.so] > profiles/r2/sass_summary.txt   (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "dvdgan_b200", "lib", "libdvdgan_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.splitlines()
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "MUFU.EX2"]
rows, cur, k = [], None, -1
counts = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur is not None:
            rows.append((cur, counts))
        k += 1
        cur, counts = names[k] if k < len(names) else m.group(1), collections.Counter()
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        for p in pats:
            if op.startswith(p):
                counts[p] += 1
                break
if cur is not None:
    rows.append((cur, counts))
print(f"# {os.path.basename(lib)}: {len(rows)} kernels; arch:",
      ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))))
print("| kernel | " + " | ".join(pats) + " |")
print("|---|" + "---:|" * len(pats))
tot = collections.Counter()
for name, c in rows:
    tot.update(c)
    if not any(c[p] for p in pats[:6]):
        continue            # list the tensor-core / TMA kernels; the rest are summed below
    short = re.sub(r"\((int|bool)\)", "", name)
    short = re.sub(r">\(.*", ">", short) if ">(" in short else re.sub(r"\(.*", "", short)
    print(f"| `{short[:110]}` | " + " | ".join(str(c[p]) for p in pats) + " |")
print("| **all kernels** | " + " | ".join(str(tot[p]) for p in pats) + " |")
