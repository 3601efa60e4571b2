"""Histogram of the sweep's inner loop in the shipped SASS (development aid): tools/sass_loop.py [lib.so|cubin] [kernel-substring]."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "raytracing-with-zig_b200/csrc/librtz.so"
key = sys.argv[2] if len(sys.argv) > 2 else "trace_kernel_constILi128ELi6"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, rows = None, []
for line in out.splitlines():
    if "Function :" in line:
        name = line.split("Function :")[1].strip()
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and name and key in name:
        rows.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(rows)}
best = None
for i, (a, ins) in enumerate(rows):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", ins)
    if m:
        t = int(m.group(1), 16)
        if t < a and t in addr:
            body = rows[addr[t]:i + 1]
            packed = sum(1 for _, x in body if re.match(r"(@\S+\s+)?(FFMA2|FADD2)", x))
            if packed >= 32 and (best is None or len(body) < len(best)):
                best = body
print(f"{key}: {len(rows)} instructions")
if best:
    h = collections.Counter(re.sub(r"^@\S+\s+", "", x).split()[0].split(".")[0] for _, x in best)
    print(f"inner loop: {len(best)} instructions:", dict(h.most_common()))
