"""Attribute ncu stall samples (SASS source page CSV) to CUDA source lines using nvdisasm line info.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <mangled-name substring> [libnsg.so]"""
import csv, os, re, subprocess, sys, tempfile
rep, kre, sub = sys.argv[1], sys.argv[2], sys.argv[3]
so = sys.argv[4] if len(sys.argv) > 4 else "navier-stokes-dealii_b200/libnsg.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# walk the function: '//## File "...", line N' markers precede instructions; instructions carry /*addr*/
line_of, cur, infn = {}, None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        infn = sub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
h = next(i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r)
hdr = rows[h]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
agg, tot, totex = {}, 0, 0
for r in rows[h + 1:]:
    if len(r) <= isamp or not r[ia]:
        continue
    if r[ia] == "Address":
        break  # next kernel instance
    a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    if base is None:
        base = a
    s, ex = int(r[isamp] or 0), int(r[iex] or 0)
    key = line_of.get(a - base, ("?", 0))
    d = agg.setdefault(key, [0, 0])
    d[0] += s
    d[1] += ex
    tot += s
    totex += ex
src_cache = {}
def src(f, n):
    for root in ("navier-stokes-dealii_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:100] if 0 < n <= len(src_cache[p]) else ""
    return ""
print(f"total samples {tot}, warp instructions {totex}")
for (f, n), (s, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "40"))]:
    print(f"{100 * s / max(tot, 1):5.1f}% smp {100 * ex / max(totex, 1):5.1f}% inst  {f}:{n:<4d} {src(f, n)}")
