"""Aggregate the warp-stall samples of an ncu report (--set full --import-source on) of one kernel by source line.
Usage: python tools/ncu_lines.py report.ncu-rep mangled_kernel_substring [lo hi]   (needs ncu, nvdisasm, cuobjdump; CPU only)"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
rng = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "imagematching-oetr_b200", "liboetr_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("tc_kernels")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
end = next(i for i in range(start + 1, len(dis)) if dis[i].startswith("\t.section"))
cur, off2line = None, {}
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][0], 16)
byline, reasons, spill, ninst = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
tot_reason = collections.Counter()
for r in data:
    if len(r) < len(hdr):
        continue
    key = off2line.get(int(r[0], 16) - base) or ("?", 0)
    s = int(r[ix["# Samples"]] or 0)
    byline[key] += s
    ninst[key] += 1
    if "LDL" in r[ix["Source"]] or "STL" in r[ix["Source"]]:
        spill[key] += 1
    for c in stall:
        v = int(r[ix[c]] or 0)
        reasons[key][c] += v
        if not (key[0] == "tc_common.cuh" and 50 <= key[1] <= 62):
            tot_reason[c] += v
tot = sum(byline.values())
print("samples %d; SASS instructions %d; stall reasons outside the mbarrier spin loop:" % (tot, len(data)))
nn = sum(tot_reason.values())
print("  " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / nn) for k, v in tot_reason.most_common(10)))
items = byline.most_common(40) if rng is None else sorted((k, v) for k, v in byline.items() if k[0] == "tc_enc.cuh" and rng[0] <= k[1] <= rng[1])
for k, v in items:
    top = ", ".join("%s %d" % (a.replace("stall_", ""), c) for a, c in reasons[k].most_common(4) if c)
    print("%-18s %4d  %5d %5.1f%%  instrs %4d spill %3d  %s" % (k[0], k[1], v, 100 * v / tot, ninst[k], spill[k], top))
