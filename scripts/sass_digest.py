"""profiles/sass_digest.md: per-kernel counts of the SASS opcodes that prove the Blackwell-native path
(tcgen05 -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP, mbarrier -> SYNCS, tcgen05.commit -> UTCBAR) next to the
legacy tensor path (HMMA) that must stay at zero.   python scripts/sass_digest.py"""
import collections, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "twisterl_b200" / "lib" / "libtwisterl_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
ops = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "ATOM", "RED"]
counts, total, fn = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        fn = re.sub(r"\(anonymous namespace\)::", "", fn)
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + "."):
                counts[fn][o] += 1
        if op.startswith("UTCHMMA") and ".2CTA" not in op:
            pass
out = ["# SASS digest of twisterl_b200/lib/libtwisterl_b200.so (cuobjdump -sass, sm_100a)", "",
       "`UTCHMMA` = tcgen05.mma kind::f16 (`.2CTA` = cta_group::2; counted in both columns), `UTCQMMA` = tcgen05.mma kind::f8f6f4 (the fp8 correction products of `f16f8c`),",
       "`LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = tensor-map TMA,",
       "`UBLKCP` = cp.async.bulk, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync (must be 0).", "",
       "| kernel | instrs | " + " | ".join(ops) + " |", "|---|---|" + "---|" * len(ops)]
for fn in sorted(total, key=lambda f: -total[f]):
    out.append(f"| `{fn}` | {total[fn]} | " + " | ".join(str(counts[fn][o]) for o in ops) + " |")
(ROOT / "profiles" / "sass_digest.md").write_text("\n".join(out) + "\n")
print("\n".join(out[:14]))
