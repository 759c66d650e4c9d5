"""Static opcode histogram of one kernel's SASS (cuobjdump -sass): python profiles/sass_ops.py <obj|so> <substring of the mangled name>"""
import subprocess, sys, collections, re
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur = None; ops = collections.Counter(); n = 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1).split('.')[0] if not m.group(1).startswith('MUFU') else m.group(1)] += 1; n += 1
print(n, "instructions")
for o, c in ops.most_common(30): print(f"{o:14s} {c}")
