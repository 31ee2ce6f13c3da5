"""Attributes the per-instruction metrics of an ncu SASS source page to CUDA source lines through
the line table of the cubin (nvdisasm -g), for reports whose CUDA view is empty (sources not on the box).
  ncu -i REP --page source --csv --print-source sass > sass.csv
  cuobjdump -xelf all OBJ.o; nvdisasm -g -c OBJ.sm_100a.cubin > dis.txt
  python scripts/ncu_lines.py sass.csv dis.txt KERNEL_SUBSTRING_NCU KERNEL_SUBSTRING_MANGLED [top]"""
import csv, re, sys
sass_csv, dis, want_ncu, want_mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
SORT = 0 if (len(sys.argv) > 6 and sys.argv[6] == "inst") else 1
# 1. line table: instruction index -> (file, line, inline chain)
lines = open(dis, errors="replace").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and want_mangled in l and l.rstrip().endswith(":"))
loc, table = ("?", 0), []
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        table.append(loc)
# 2. ncu rows of the wanted kernel (last occurrence = last captured launch)
rows = list(csv.reader(open(sass_csv)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
blk = [b for b in blocks if want_ncu in b["name"]][-1]
ix = {h: i for i, h in enumerate(blk["hdr"])}
data = blk["data"]
print(blk["name"][:90], "ncu instructions", len(data), "line-table instructions", len(table))
agg = {}
ti = ts = 0
for k, r in enumerate(data):
    key = table[k] if k < len(table) else ("?", 0)
    a = agg.setdefault(key, [0, 0, 0])
    ie, s, th = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), int(r[ix["Thread Instructions Executed"]])
    a[0] += ie; a[1] += s; a[2] += th
    ti += ie; ts += s
print("warp instructions %d, samples %d" % (ti, ts))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
    print("%-22s:%4d  inst %5.1f%%  samples %5.1f%%  lanes %4.1f" % (key[0], key[1], 100 * a[0] / ti, 100 * a[1] / max(ts, 1), a[2] / max(a[0], 1)))
