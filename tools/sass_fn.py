"""Prints the SASS of one kernel of libb200sgm.so (substring match on the mangled name) with an opcode histogram.
usage: python tools/sass_fn.py <name-substring> [start_hex end_hex]"""
import subprocess, sys, re, collections
lib = "i3dr_stereo_camera-ros_b200/libb200sgm.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None; body = {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); body[fn] = []; continue
    if fn: body[fn].append(line)
keys = [k for k in body if sys.argv[1] in k]
if len(keys) != 1:
    print("\n".join(keys)); sys.exit(1)
lines = body[keys[0]]
lo = int(sys.argv[2], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
hist = collections.Counter(); n = 0
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if not m: continue
    a = int(m.group(1), 16)
    if a < lo or a > hi: continue
    ins = m.group(2).strip()
    print("%05x  %s" % (a, ins))
    op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    hist[op.split(".")[0]] += 1; n += 1
print("---- %d instructions" % n, file=sys.stderr)
for k, v in hist.most_common(): print("%6d %s" % (v, k), file=sys.stderr)
