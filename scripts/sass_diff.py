"""Compare two `cuobjdump -sass` dumps function by function (order-independent): prints the kernels whose SASS differs, and those only in one
dump.  Used to prove that a host-side or header refactor left the device code of the existing kernels untouched."""
import hashlib
import re
import sys


def functions(path):
    out, name, buf = {}, None, []
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = hashlib.md5("".join(buf).encode()).hexdigest()
            name, buf = re.sub(r"_GLOBAL__N__[0-9a-f]+_(\d+_\w+?_cu)_[0-9a-f]+", r"_GLOBAL__N__\1", m.group(1)), []     # anonymous namespaces carry a per-path hash
        elif name:
            buf.append(" ".join(line.split()) + "\n")       # cuobjdump pads the columns to the longest instruction of the whole dump
    if name:
        out[name] = hashlib.md5("".join(buf).encode()).hexdigest()
    return out


a, b = functions(sys.argv[1]), functions(sys.argv[2])
changed = sorted(k for k in a if k in b and a[k] != b[k])
print(f"{len(a)} / {len(b)} functions; changed: {len(changed)}; only in first: {len(set(a) - set(b))}; only in second: {len(set(b) - set(a))}")
for k in changed:
    print("  changed", k)
for k in sorted(set(a) ^ set(b)):
    print("  only in", "first" if k in a else "second", k)
