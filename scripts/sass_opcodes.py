"""Per-kernel counts of the Blackwell-specific SASS opcodes in libadpst.so (cuobjdump -sass): tcgen05 MMAs (UTCHMMA/UTCQMMA...),
TMA loads/stores (UTMALDG/UTMASTG), tensor-memory loads/stores (LDTM/STTM), tcgen05 barriers/commits (UTCBAR), TMEM allocation
(UTCATOMSWS...), plus mbarrier (SYNCS) and classic MMA (HMMA) for contrast.  Writes profiles/sass_opcodes.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "automated-deep-photo-style-transfer_b200", "libadpst.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMXQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTCCP",
       "SYNCS", "HMMA", "DFMA", "LDG", "STG", "LDS", "STS", "LDGSTS", "REDG", "ATOMG", "BAR"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_all"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[kern][o] += 1
                total[o] += 1
lines = ["SASS opcode counts per kernel of libadpst.so (sm_100a), from `cuobjdump -sass` (scripts/sass_opcodes.py)", ""]
cols = [o for o in OPS if total[o]]
lines.append("%-72s %6s " % ("kernel", "instr") + " ".join("%8s" % c for c in cols))
for k, c in counts.items():
    lines.append("%-72s %6d " % (k[:72], c["_all"]) + " ".join("%8d" % c[o] for o in cols))
lines.append("")
lines.append("%-72s %6s " % ("total", "") + " ".join("%8d" % total[o] for o in cols))
text = "\n".join(lines) + "\n"
path = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
open(path, "w").write(text)
print(text[:3000])
