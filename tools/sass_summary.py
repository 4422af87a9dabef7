"""Opcode census of libb200mm.so per kernel (cuobjdump -sass; runs without a GPU):  python tools/sass_summary.py > profiles/sass_summary_r02.txt
Columns: tcgen05 MMA issue (UTCHMMA / UTCQMMA ...), TMEM loads / stores (LDTM / STTM), TMA loads / stores (UTMALDG / UTMASTG), tcgen05.commit
(UTCBAR), mbarrier (SYNCS), MUFU, and the legacy warp-level tensor instruction HMMA (mma.sync) — which must be 0 everywhere."""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "ant-multi-modal-framework_b200", "libb200mm.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU", "HMMA"]
tot = collections.Counter()
print(f"# {os.path.relpath(lib, root)}: {len(funcs)} kernels, sm_100a SASS (cuobjdump -sass), opcode counts per kernel")
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for (f, c), n in zip(funcs.items(), names):
    n = re.sub(r"\(.*", "", n).replace("b200mm::", "")
    print(f"| `{n[:70]}` | {sum(c.values())} | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
    for k in cols:
        tot[k] += c.get(k, 0)
print("| **total** | " + str(sum(sum(c.values()) for c in funcs.values())) + " | " + " | ".join(str(tot[k]) for k in cols) + " |")
print()
print(f"HMMA (mma.sync) instructions in the library: {tot['HMMA']}  — every tensor-core contraction is issued as tcgen05.mma (UTCHMMA), operands arrive "
      f"by TMA (UTMALDG), accumulators are read with tcgen05.ld (LDTM). UTMASTG = {tot['UTMASTG']}: results leave through st.global after a shared-memory "
      "transpose (no TMA store yet).")
