"""Summarise an ncu source-page export: executed warp instructions and stall samples per SASS region between branch
targets, for one kernel.  usage: ncu_regions.py report.ncu-rep kernel_regex [min_pct]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; blocks.append((r[1], cur)); continue
    if cur is not None and r: cur.append(r)
name, b = blocks[0]
hdr, data = b[0], b[1:]
ia, isrc, iex, ith, isamp = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
base = int(data[0][ia], 16)
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[isamp]) for r in data)
print(name[:100]); print("warp-inst", tot, "samples", tots, "static instrs", len(data))
# regions: split where the executed count changes by more than 2%
regs = []; start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(int(data[i][iex]) - int(data[i - 1][iex])) > 0.02 * max(int(data[i - 1][iex]), 1):
        regs.append((start, i)); start = i
for s, e in regs:
    ex = sum(int(r[iex]) for r in data[s:e]); th = sum(int(r[ith]) for r in data[s:e]); sm = sum(int(r[isamp]) for r in data[s:e])
    if 100 * ex / tot >= min_pct or 100 * sm / tots >= min_pct:
        ops = collections.Counter(r[isrc].split()[0 if not r[isrc].strip().startswith('@') else 1].split('.')[0] for r in data[s:e])
        top = " ".join(f"{k}:{v}" for k, v in ops.most_common(5))
        print(f"{int(data[s][ia],16)-base:#06x}-{int(data[e-1][ia],16)-base:#06x} n={e-s:4d} execs/instr={int(data[s][iex]):8d} inst {100*ex/tot:5.1f}% thr/inst {th/max(ex,1):5.1f} samples {100*sm/tots:5.1f}%  {top}")
