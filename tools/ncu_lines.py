"""Attribute an .ncu-rep's per-instruction counts to CUDA source lines (read here, no GPU).

usage: ncu_lines.py <report.ncu-rep> <library.so> <kernel-name-substring> [body-function-first-line body-function-last-line]
Joins `ncu --page source --csv` (per SASS address) with `nvdisasm -gi` (address -> inlined line chain) and
prints, per source line of the kernel body, the executed warp-instructions, the issue-cycle model
(packed FP32x2 = 2, everything else = 1), the stall samples and the average active threads."""
import csv, io, re, subprocess, sys, tempfile, os, glob, collections
rep, so, kname = sys.argv[1:4]
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
INNER = os.environ.get('INNER') == '1'   # attribute to the innermost (inlined) frame instead
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
addr2chain, chain, inside, fresh = {}, [], False, True
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside, chain, fresh = kname in ln, [], True
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:  # consecutive "//## File" lines form one chain: innermost inlined frame first, kernel frame last
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:  # an instruction: it belongs to the chain above it (or to the previous instruction's chain)
        addr2chain[int(m.group(1), 16)] = list(chain)
        fresh = True
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
iA, iS, iSamp, iEx, iThr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
base = int(data[0][iA], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0.0])
tot = [0, 0, 0]
for r in data:
    a = int(r[iA], 16) - base
    ch = addr2chain.get(a, [])
    # the frame inside the kernel body: last chain entry whose line lies in [lo, hi]
    key = None
    for f, l in reversed(ch):
        if f.endswith("rtz_kernels.cuh") and lo <= l <= hi:
            key = l
            break
    if key is None:
        key = ch[-1][1] if ch else -1
    if INNER and ch:
        key = (ch[0][0], ch[0][1])
    op = r[iS].split()
    op = op[1] if op[0].startswith("@") else op[0]
    w = 2 if op in ("FFMA2", "FADD2", "FMUL2") else 1
    e, s = int(r[iEx]), int(r[iSamp])
    x = agg[key]
    x[0] += e; x[1] += e * w; x[2] += s; x[3] += float(r[iThr]) * e
    tot[0] += e; tot[1] += e * w; tot[2] += s
srcdir = os.path.dirname(os.path.abspath(so))
lines = open(os.path.join(srcdir, "rtz_kernels.cuh")).read().splitlines()
files = {f: open(os.path.join(srcdir, f)).read().splitlines() for f in ("rtz_kernels.cuh", "rtz_device.cuh")}
print(f"total inst {tot[0]}  model cycles {tot[1]}  samples {tot[2]}")
for k in sorted(agg, key=str):
    e, c, s, t = agg[k]
    if c / tot[1] < (0.0015 if INNER else 0.002):
        continue
    if isinstance(k, tuple):
        fl = files.get(k[0], [])
        txt = fl[k[1] - 1].strip()[:70] if 0 < k[1] <= len(fl) else ""
        k = f"{k[0][4:10]}:{k[1]}"
    else:
        txt = lines[k - 1].strip()[:70] if 0 < k <= len(lines) else ""
    print(f"{k!s:>12} inst {100*e/tot[0]:5.2f}% cyc {100*c/tot[1]:5.2f}% samp {100*s/tot[2]:5.2f}% thr {t/max(e,1):4.1f} | {txt}")
