"""Join an ncu source-page CSV (SASS view, --page source --csv) with nvdisasm -g line info:
per CUDA source line: executed warp-instructions and stall samples.
usage: python tools/sass_lines.py <ncu_source.csv> <nvdisasm -g -c output> <kernel-substring> [top]"""
import csv, re, sys, collections
ncu_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: find the function section, collect (offset -> (file,line))
lines = open(sass).read().split("\n")
cur = None; in_fn = False; offs = []
for ln in lines:
    if ln.startswith("\t.section") or ln.startswith(".section"):
        in_fn = kern in ln and ".text." in ln
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: offs.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]; ia = hdr.index("Address"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > ie]
base = int(data[0][ia], 16)
by_off = {o: (fl, txt) for o, fl, txt in offs}
agg = collections.defaultdict(lambda: [0.0, 0.0]); tot_i = tot_s = 0.0; miss = 0
for r in data:
    off = int(r[ia], 16) - base
    fl = by_off.get(off, (None, ""))[0]
    if fl is None: miss += 1
    n = float(r[ie] or 0); s = float(r[isamp] or 0)
    agg[fl][0] += n; agg[fl][1] += s; tot_i += n; tot_s += s
print("instructions", tot_i, "samples", tot_s, "unmapped rows", miss, "sass rows", len(data), "disasm", len(offs))
src_cache = {}
def src(fl):
    if fl is None: return ""
    f, l = fl
    import glob
    if f not in src_cache:
        c = glob.glob("audio_fewshot_b200/csrc/" + f)
        src_cache[f] = open(c[0]).read().split("\n") if c else []
    s = src_cache[f]
    return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""
for fl, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %-22s %s" % (100 * n / tot_i, 100 * s / max(tot_s, 1), "%s:%s" % fl if fl else "?", src(fl)))
