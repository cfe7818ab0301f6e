"""Maps local-memory (spill) instructions of one kernel instantiation to source lines.
   usage: python tools/spill_lines.py <cubin> <ITEMS>   (cubin: cuobjdump -xelf all <object built with -lineinfo>)"""
import collections
import re
import subprocess
import sys

cubin, items = sys.argv[1], sys.argv[2]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
inside, line = False, None
spills, insts = collections.Counter(), collections.Counter()
for l in sass:
    if l.startswith(".text."):
        inside = ("Li%sELi1E" % items) in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,8}\*/", l):
        insts[line] += 1
        if re.search(r"\b(STL|LDL)(\.|\b)", l):
            spills[line] += 1
print("instructions: %d, spill instructions: %d" % (sum(insts.values()), sum(spills.values())))
for k, v in sorted(spills.items(), key=lambda t: -t[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("%-18s line %5d  spill insts %4d of %5d" % (k[0], k[1], v, insts[k]))
