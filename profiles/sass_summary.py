"""Opcode histogram per kernel of the shipped library: python profiles/sass_summary.py > profiles/r2_sass_summary.txt
(cuobjdump -sass pyqmc_b200/libqmcb200.so).  Shows what the kernels are made of: FP64 (DFMA/DADD/DMUL/MUFU.RCP64H),
bulk-async table staging (UBLKCP + SYNCS = cp.async.bulk + mbarrier), and one tensor-core kernel: DMMA in
k_gemm_tn_dmma, the SR overlap matrix (FP64 has no tcgen05 path, so no UTCMMA)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "pyqmc_b200", "libqmcb200.so")], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
arch = re.search(r"arch = (\S+)", out)
print(f"# cuobjdump -sass pyqmc_b200/libqmcb200.so   ({arch.group(1) if arch else '?'}); opcode counts per kernel")
total = collections.Counter()
for name, c in kernels.items():
    total.update(c)
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:100]
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(8))
    print(f"{sum(c.values()):7d}  {demangled}\n         {top}")
print("\n# whole library")
for k in ("DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "UTCMMA", "HMMA", "DMMA", "LDTM"):
    print(f"{k:8s} {total.get(k, 0)}")
