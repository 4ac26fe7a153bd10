"""Compare the SASS of every conv_tc_kernel instance between two object files (addresses stripped): used to show that
adding opt-in kernel variants left the GPU-validated default instances bit-identical.

    python scripts/sass_diff.py old_conv_tc.o keras-ocr_b200/csrc/conv_tc.o [suffix-to-append-to-old-template-args]

e.g. after a new trailing template parameter `bool BOX16 = false` pass `ELb0` so that old `<64,64,3,true>` is compared
with new `<64,64,3,true,false>`.
"""
import collections
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    table, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            full = m.group(1)
            name = full[full.index("kernelI"):] if "kernelI" in full else full
            table[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and name:
            table[name].append(re.sub(r"\s+", " ", m.group(1)))
    return table


def main():
    old, new = functions(sys.argv[1]), functions(sys.argv[2])
    suffix = sys.argv[3] if len(sys.argv) > 3 else ""
    same = 0
    for name, body in old.items():
        other = name.replace("EEEv14CUtensorMap", suffix + "EEEv14CUtensorMap") if suffix else name
        if body == new.get(other):
            same += 1
        else:
            print("DIFFERENT" if other in new else "MISSING  ", name[:60])
    print(f"{same} of {len(old)} instances identical; {len(new) - len(old)} instances only in the second file")


if __name__ == "__main__":
    main()
