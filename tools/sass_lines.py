"""Static instruction count per source line of a kernel body (development aid; no GPU): tools/sass_lines.py lib.so kernel-substring [lo hi]"""
import re, subprocess, sys, tempfile, os, glob, collections
so, kname = sys.argv[1:3]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
chain, inside, fresh = [], False, True
agg = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside, chain, fresh = kname in ln, [], True
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:
        key = None
        for f, l in reversed(chain):
            if f.endswith("rtz_kernels.cuh") and lo <= l <= hi:
                key = l
                break
        if key is None:
            key = chain[-1][1] if chain else -1
        agg[key] += 1
        op = m.group(2).split()
        op = op[1] if op[0].startswith("@") else op[0]
        ops[key][op.split(".")[0]] += 1
        fresh = True
src = open(os.path.join(os.path.dirname(os.path.abspath(so)), "rtz_kernels.cuh")).read().splitlines()
print("total", sum(agg.values()))
for k in sorted(agg):
    if agg[k] >= 4:
        print(f"{k:5d} {agg[k]:5d}  {src[k-1].strip()[:60]:60s} {dict(ops[k].most_common(5))}")
