#!/usr/bin/env python
"""SASS of one kernel of libtinybvh_b200.so (cuobjdump, no GPU needed) with a static opcode histogram in front.
usage: tools/sass_extract.py <mangled-name-substring> <out.txt>"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "tinybvh_b200", "libtinybvh_b200.so")


def main():
    want, out = sys.argv[1], sys.argv[2]
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn = [m for m in re.findall(r"Function : (\S+)", names) if want in m]
    assert len(fn) == 1, f"{want!r} matches {fn}"
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], LIB], capture_output=True, text=True).stdout
    body = txt[txt.index("Function : " + fn[0]):]
    body = body[: body.index("....") + 100] if "\n\t\t.....\n" in body else body
    lines = [l for l in body.splitlines() if not re.match(r"^\s*/\* 0x[0-9a-f]{16} \*/\s*$", l)]
    lines = [re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", l) for l in lines]
    lines = [l for l in lines if not l.startswith("Fatbin") and not l.startswith("====") and not re.match(r"^(arch|code version|host|compile_size|identifier) =", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", l)
        if m:
            ops[m.group(1)] += 1
    demangled = subprocess.run(["cu++filt", fn[0]], capture_output=True, text=True).stdout.strip() or fn[0]
    with open(out, "w") as f:
        f.write(f"# {demangled}\n# cuobjdump -sass of tinybvh_b200/libtinybvh_b200.so (nvcc 12.9, sm_100a); {sum(ops.values())} instructions (static)\n")
        f.write("# opcode histogram (static): " + ", ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n\n")
        f.write("\n".join(l.rstrip() for l in lines if l.strip()) + "\n")
    print(out, sum(ops.values()), "instructions;", ", ".join(f"{k} {v}" for k, v in ops.most_common(8)))


if __name__ == "__main__":
    main()
