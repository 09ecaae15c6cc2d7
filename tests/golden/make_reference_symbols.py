"""Snapshot of the names GalerkinToolkit v0.6.3 defines at module level (functions, generic-function stubs, structs,
abstract types, constants), extracted from /root/reference/src/*.jl.  tests/test_julia_shim.py checks every `GT.<name>` the
Julia shim uses against this list (and, where /root/reference is present, against the sources themselves).

    python tests/golden/make_reference_symbols.py > tests/golden/reference_symbols.txt
"""
import glob
import re
import sys

PATTERNS = [
    r"^\s*(?:@inline\s+|@noinline\s+)?function\s+(?:[A-Za-z_][\w\.]*\.)?([A-Za-z_∫∇][\w!]*)\s*(?:\(|end|\{)",   # function f(...) / function f end
    r"^(?:[A-Za-z_][\w\.]*\.)?([A-Za-z_][\w!]*)\s*\([^=\n]*\)\s*(?:where\s+[^=\n]+)?=(?!=)",                  # f(x) = ...
    r"^\s*(?:mutable\s+)?struct\s+([A-Za-z_]\w*)",
    r"^\s*abstract\s+type\s+([A-Za-z_]\w*)",
    r"^\s*const\s+([A-Za-z_∫]\w*)\s*=",
    r"^\s*@enum\s+([A-Za-z_]\w*)((?:\s+[A-Za-z_]\w*(?:=\d+)?)*)",
]


def symbols(src_dir):
    names = set()
    for path in sorted(glob.glob(src_dir + "/*.jl")):
        for line in open(path, encoding="utf-8"):
            if line.lstrip().startswith("#"):
                continue
            for pat in PATTERNS:
                m = re.match(pat, line)
                if m:
                    names.add(m.group(1))
                    if pat.startswith(r"^\s*@enum"):
                        for item in m.group(2).split():
                            names.add(item.split("=")[0])
    return names


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src"
    for n in sorted(symbols(src)):
        print(n)
