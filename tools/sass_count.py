"""Count SASS instructions per kernel and per opcode class from `cuobjdump -sass` on stdin (static probe helper)."""
import collections, re, sys
cur = None
counts = collections.OrderedDict()
for ln in sys.stdin:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and cur:
        txt = re.sub(r"^@!?U?P\d+\s+", "", m.group(1))
        op = txt.split()[0].split(".")[0]
        if op in ("NOP", "BRA", "EXIT"):
            continue
        counts[cur][op] += 1
for k, c in counts.items():
    fp = sum(v for o, v in c.items() if o in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2"))
    print("%-28s total %3d  fp %3d  %s" % (re.sub(r"^_Z\d+", "", k)[:28], sum(c.values()), fp,
                                           " ".join("%s=%d" % kv for kv in c.most_common(12))))
