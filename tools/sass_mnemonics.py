"""Mnemonic counts per kernel of the built library (cuobjdump -sass; no GPU needed) -> profiles/r2_sass_mnemonics.txt."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = ROOT / "interactvlm_b200" / "libivlm_b200.so"
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKPF", "SYNCS", "HMMA", "LDSM", "MUFU.EX2", "SHFL"]
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
counts, order, cur = collections.defaultdict(collections.Counter), [], None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", next(it))
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for c in cols:
            if op == c or op.startswith(c + "."):
                counts[cur][c] += 1
out = ["# SASS mnemonics per kernel of interactvlm_b200/libivlm_b200.so (cuobjdump -sass, sm_100a), round 2 final tree",
       "# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG = TMA tensor load, UTMAPF / UBLKPF = TMA / bulk L2 prefetch,",
       "# SYNCS = mbarrier, HMMA/LDSM = mma.sync/ldmatrix (weight-streaming decode kernels, decoder tail, legacy flash attention), SHFL = warp shuffles",
       f"{'kernel':<72}" + "".join(f"{c:>9}" for c in cols)]
for k in order:
    if any(counts[k].values()):
        out.append(f"{k[:71]:<72}" + "".join(f"{counts[k][c]:>9}" for c in cols))
(ROOT / "profiles" / "r2_sass_mnemonics.txt").write_text("\n".join(out) + "\n")
print("\n".join(out[:3]), f"\n{len(out) - 4} kernels", file=sys.stderr)
