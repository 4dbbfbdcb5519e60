"""Per-role summary of an ncu source page of conv1_tc_kernel (BUILD / MMA / EPI): warp-instructions per tile and the
stall-reason mix of each role, plus the source lines that collect the most stall samples.

usage: python tools/stem_roles.py <ncu --page source --csv> <nvdisasm -g -c of the cubin> <kernel substring> <clips>
(role boundaries are the line numbers of the three role branches in csrc/conv1_tc.cu, found by their marker comments)"""
import collections
import csv
import re
import sys

src_csv, sass_path, kern, clips = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
cu = open("audio_fewshot_b200/csrc/conv1_tc.cu").read().split("\n")
mark = {name: next(i + 1 for i, l in enumerate(cu) if tag in l)
        for name, tag in (("BUILD", "=== BUILD"), ("MMA", "=== MMA issuer"), ("EPI", "=== EPI"))}
end = next(i + 1 for i, l in enumerate(cu) if l.startswith("#undef BAR1"))


def role_of(line):
    if line < mark["BUILD"]:
        return "setup"
    if line < mark["MMA"]:
        return "BUILD"
    if line < mark["EPI"]:
        return "MMA"
    return "EPI" if line < end else "waits"


cur, in_fn, offs = None, False, {}
for ln in open(sass_path).read().split("\n"):
    if ln.startswith("\t.section") or ln.startswith(".section"):
        in_fn = kern in ln and ".text." in ln
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        offs[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ix["Address"]], 16)
agg = collections.defaultdict(collections.Counter)
lines = collections.Counter()
role = "setup"
for r in data:
    fl = offs.get(int(r[ix["Address"]], 16) - base)
    if fl and fl[0] == "conv1_tc.cu":
        role = role_of(fl[1])
    a = agg[role]
    a["inst"] += float(r[ix["Instructions Executed"]] or 0)
    a["samples"] += float(r[ix["# Samples"]] or 0)
    for h in stalls:
        a[h] += float(r[ix[h]] or 0)
    if fl:
        lines[fl] += float(r[ix["# Samples"]] or 0)
tiles = clips * 42 * 52 / 128.0
total = sum(a["samples"] for a in agg.values())
print("kernel %s, %d clips = %.0f tiles of 128 pooled pixels; %d stall samples" % (kern, clips, tiles, total))
print("role    warp-instr/tile  samples  share  issuing  top stall reasons")
for k in ("BUILD", "MMA", "EPI", "setup", "waits"):  # waits: the out-of-line mbarrier wait loops of all roles
    a = agg[k]
    if not a["samples"]:
        continue
    top = sorted(((h[6:], a[h]) for h in stalls if h != "stall_selected"), key=lambda kv: -kv[1])[:4]
    print("%-7s %10.0f %12.0f %6.1f%% %7.1f%%  %s" % (
        k, a["inst"] / tiles, a["samples"], 100 * a["samples"] / total, 100 * a["stall_selected"] / a["samples"],
        ", ".join("%s %.0f%%" % (n, 100 * v / a["samples"]) for n, v in top)))
print("source lines with the most stall samples:")
for (f, l), v in lines.most_common(8):
    text = cu[l - 1].strip()[:80] if f == "conv1_tc.cu" and 0 < l <= len(cu) else ""
    print("  %5.1f%%  %s:%d  %s" % (100 * v / total, f, l, text))
