#!/usr/bin/env python
"""Compare two builds of libsod_b200.so kernel by kernel at the SASS level (no GPU needed).

    python tools/sass_diff.py OLD.so NEW.so

Used when a change is meant to leave already-validated kernels untouched (e.g. adding a template variant behind a
flag while no GPU is available): every kernel present in both libraries must have an identical instruction stream.
Template arguments that were appended with a default (`<T, 16>` → `<T, 16, false>`) are matched by dropping trailing
`, false` arguments.  Exit code 1 if any common kernel differs.
"""
import re
import subprocess
import sys


def kernels(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"(, false)+>\(", ">(", cur)
            out[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            out[cur].append(re.sub(r"/\*[0-9a-f]{4,}\*/", "", line).strip())
    return out


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    changed = [k for k in old if k in new and old[k] != new[k]]
    for k in sorted(set(old) - set(new)):
        print("removed :", k)
    for k in sorted(set(new) - set(old)):
        print("added   :", k)
    for k in changed:
        print("CHANGED :", k, f"({len(old[k])} → {len(new[k])} instructions)")
    print(f"{len(set(old) & set(new)) - len(changed)} kernels identical, {len(changed)} changed")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
