"""Aggregate ncu warp-stall samples per CUDA source line for one kernel.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [launch-index] [lib.so]
Correlates the SASS page of the report (instruction order) with `nvdisasm -g` line info of the same cubin."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, ksub = sys.argv[1], sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib = sys.argv[4] if len(sys.argv) > 4 else "smart-calibration_b200/csrc/libsmartcal_b200.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
lines_by_func = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_f, cur_line, seq = None, None, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_f = m.group(1); seq = lines_by_func.setdefault(cur_f, []); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if seq is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            seq.append(cur_line)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", "0"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if r]
# the page prints one block per profiled launch: "Kernel Name" row, header row, instruction rows
starts = [i for i, r in enumerate(rows) if r[0] == "Kernel Name"]
blk = starts[min(launch, len(starts) - 1)]
nxt = [i for i in starts if i > blk]
kname = rows[blk][1]
H = rows[blk + 1]; data = rows[blk + 2:(nxt[0] if nxt else len(rows))]
isamp, iex = H.index("# Samples"), H.index("Instructions Executed")
func = [k for k in lines_by_func if ksub in k and len(lines_by_func[k]) == len(data)]
if not func:
    cands = {k: len(v) for k, v in lines_by_func.items() if ksub in k}
    sys.exit(f"no function with {len(data)} instructions matching {ksub}: {cands} (kernel {kname})")
seq = lines_by_func[func[0]]
agg, aggx = collections.Counter(), collections.Counter()
for r, ln in zip(data, seq):
    agg[ln] += int(r[isamp] or 0); aggx[ln] += int(r[iex] or 0)
tot, totx = sum(agg.values()), sum(aggx.values())
print(f"kernel {kname}  function {func[0]}  samples {tot}  instructions {totx}")
src_cache = {}
def src(ln):
    if ln is None: return ""
    f, n = ln
    for root in ("smart-calibration_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            L = src_cache.setdefault(p, open(p).read().splitlines())
            return L[n - 1].strip()[:90] if n <= len(L) else ""
    return ""
for ln, s in agg.most_common(45):
    print(f"{s/tot*100:5.1f}% samp {aggx[ln]/totx*100:5.1f}% inst  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<5d} {src(ln)}")
