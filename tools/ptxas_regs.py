"""Registers / spills of the tensor-path kernels from a build log:  python tools/ptxas_regs.py [log] [filter]"""
import re
import sys

log = sys.argv[1] if len(sys.argv) > 1 else 'cbf_ssm_b200/csrc/build/inst_4_2_2.ptxas.log'
flt = sys.argv[2] if len(sys.argv) > 2 else 'tc_kernel'
for b in open(log).read().split("ptxas info    : Compiling entry function ")[1:]:
    name = b.split("'")[1]
    if flt not in name:
        continue
    m = re.search(r"_ZN3cbf\d+(\w+?_kernel)I(.*?)EEv", name)
    args = re.findall(r"Li(\d+)E", m.group(2)) if m else []
    sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
    rg = re.search(r"Used (\d+) registers", b)
    print("%-24s %-22s regs %3s  spill %s/%s" % (m.group(1) if m else name[:40], ",".join(args), rg.group(1), sp.group(1), sp.group(2)))
