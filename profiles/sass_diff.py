"""Compare the instruction streams of two cuobjdump -sass dumps kernel by kernel (addresses and encodings ignored).
Used when a change must leave already-validated kernels untouched:
    cuobjdump -sass viamd_b200/build/sdf.o > /tmp/before.sass ; <edit, rebuild> ; cuobjdump -sass viamd_b200/build/sdf.o > /tmp/after.sass
    python profiles/sass_diff.py /tmp/before.sass /tmp/after.sass
"""
import re
import sys


def kernels(path):
    out, cur = {}, None
    for line in open(path):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); out[cur] = []; continue
        if cur is None:
            continue
        t = re.sub(r"/\*.*?\*/", "", line).strip()
        if t:
            out[cur].append(t)
    return out


def main(a, b):
    ka, kb = kernels(a), kernels(b); bad = 0
    for k in sorted(set(ka) | set(kb)):
        if k not in ka: print("added    ", k, len(kb[k]), "instructions")
        elif k not in kb: print("removed  ", k); bad += 1
        elif ka[k] != kb[k]: print("CHANGED  ", k, len(ka[k]), "->", len(kb[k])); bad += 1
        else: print("identical", k, len(ka[k]))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:3]))
