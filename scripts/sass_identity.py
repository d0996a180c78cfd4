import sys, re, hashlib, subprocess
def hashes(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, h = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); h[cur] = hashlib.sha1(); continue
        if cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            # instruction text without the address and encoding comments
            txt = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()
            h[cur].update(txt.encode())
    return {k: v.hexdigest() for k, v in h.items()}
# usage: python scripts/sass_identity.py old.o new.o
a, b = hashes(sys.argv[1]), hashes(sys.argv[2])
common = set(a) & set(b)
diff = [k for k in common if a[k] != b[k]]
print("kernels old/new/common:", len(a), len(b), len(common), "differing:", len(diff))
for k in diff[:10]: print("  DIFF", k)
print("only old:", [k for k in a if k not in b][:5], "...", len([k for k in a if k not in b]))
print("only new:", [k for k in b if k not in a][:8])
