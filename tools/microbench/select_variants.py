"""Opcode counts of tools/microbench/select_variants.cu per kernel (static SASS, no GPU):
   python tools/microbench/select_variants.py"""
import collections
import sys
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
cubin = os.path.join(tempfile.mkdtemp(), "sel.cubin")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-cubin", "-o", cubin,
                       (sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "select_variants.cu"))])
sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", cubin], stdout=subprocess.PIPE).stdout.decode()
kern, ops = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\w+)", line)
    if m:
        kern = m.group(1)
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        ops[kern][m.group(1).split(".")[0]] += 1
SKIP = {"LDC", "S2R", "LDG", "STG", "EXIT", "BRA", "NOP", "LDCU", "IMAD", "ULDC", "S2UR", "MOV", "UMOV", "IADD3", "LEA"}
print("# select/clamp work per element (address arithmetic, loads, stores excluded); DSETP runs on the FP64 pipe")
for k in sorted(ops):
    work = {o: n for o, n in ops[k].items() if o not in SKIP}
    print("%-10s %2d instructions: %s" % (k, sum(work.values()), "  ".join("%s x%d" % (o, n) for o, n in sorted(work.items()))))
