#!/usr/bin/env python
"""Print registers / spills of every compiled kernel from the ptxas logs in build/."""
import glob, os, re, sys
here = os.path.dirname(os.path.abspath(__file__))
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(here, "build", "*.ptxas.log"))):
    for e in re.split(r"Compiling entry function ", open(f).read())[1:]:
        name = re.search(r"_ZN3cbf\d+(\w+?)_kernel", e).group(1)
        if flt not in name:
            continue
        st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", e)
        rg = re.search(r"Used (\d+) registers", e)
        print(f"{os.path.basename(f)[:-10]:18s} {name:18s} stack={st.group(1):>5s} spill={st.group(2):>5s} regs={rg.group(1)}")
