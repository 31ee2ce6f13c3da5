"""Summarises the SASS source page of an ncu report in chunks of 16 instructions:
share of issued instructions, share of stall samples, average active threads.
  ncu -i REP --page source --csv --print-source sass > src.csv; python scripts/ncu_chunks.py src.csv [chunk]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 16
want = sys.argv[3] if len(sys.argv) > 3 else ""  # substring of the kernel name (default: first kernel)
hdr, data, k, on = None, [], 0, False
for r in rows:
    if r and r[0] == "Kernel Name":
        if on:
            break
        on = want in r[1]
        if on:
            print(r[1][:100])
        continue
    if not on:
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
ti = sum(int(r[ix["Instructions Executed"]]) for r in data)
ts = sum(int(r[ix["# Samples"]]) for r in data)
tt = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
print("instructions %d, warp-level %d, thread-level %d (avg %.1f lanes), samples %d" % (len(data), ti, tt, tt / ti, ts))
for g in range(0, len(data), step):
    ch = data[g:g + step]
    ie = sum(int(c[ix["Instructions Executed"]]) for c in ch)
    s = sum(int(c[ix["# Samples"]]) for c in ch)
    th = sum(int(c[ix["Thread Instructions Executed"]]) for c in ch)
    if ie * 1000 < ti and s * 1000 < ts:
        continue
    print("%4d-%4d inst %5.1f%% samples %5.1f%% lanes %4.1f | %s" % (g, g + len(ch) - 1, 100 * ie / ti, 100 * s / ts, th / max(ie, 1), ch[0][ix["Source"]].strip()[:60]))
