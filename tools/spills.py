"""List the kernels of a *.ptxas.log (written by build.py) that spill registers: name, stack / spill bytes, registers."""
import re
import subprocess
import sys

for path in sys.argv[1:]:
    t = open(path).read()
    ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for \1\n"
                      r"\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", t)
    for e in ents:
        if int(e[2]) > 0 or int(e[3]) > 0:
            name = subprocess.run(["c++filt", e[0]], capture_output=True, text=True).stdout.strip()
            print("%-40s %-70s stack %4s  spill st %4s ld %4s  regs %s" % (path.split("/")[-1], name[:70], e[1], e[2], e[3], e[4]))
