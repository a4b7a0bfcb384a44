#!/usr/bin/env python
"""SASS digest of the hot kernels of libtmc_b200.so (no GPU needed): per kernel the instruction count, registers and the
Blackwell-specific mnemonics that prove how it moves data (UTMALDG = tensor-map TMA, UBLKCP = 1-D bulk TMA, SYNCS =
mbarrier, LDGSTS = cp.async, FFMA2/FMUL2/FADD2 = packed fp32x2).  Usage: python tools/sass_digest.py > profiles/<tag>_sass_digest.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torch_motion_correction_b200", "libtmc_b200.so")
HOT = ["warp_tma_kernel", "warp_lattice_kernel", "local_loss_tile_kernel", "local_coefficient_kernel", "rows_forward_poly",
       "rows_inverse_argmax_poly", "rows_forward_real2n", "rows_inverse_argmax_real2n", "rows_forward_p2", "cols_forward_p2", "cols_inverse_p2", "rows_inverse_argmax_p2",
       "xc_leave_one_out_kernel", "stats_partial_kernel", "convert_stack_kernel", "lattice_xinterp_kernel"]
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "FFMA", "LDS", "STS", "LDG", "STG", "BAR.SYNC", "SHFL", "MUFU"]

res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+).*?SHARED:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print(f"# {os.path.basename(LIB)}: arch {sorted(set(re.findall(r'arch = (sm_\w+)', sass)))}, {len(regs)} functions")
print("kernel | registers | static smem | SASS instructions | " + " ".join(KEYS))
for block in sass.split("Function : ")[1:]:
    name = block.split("\n", 1)[0].strip()
    if not any(h in name for h in HOT):
        continue
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", block, flags=re.M)
    c = collections.Counter()
    for op in ops:
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("FFMA", "LDS", "STS", "LDG", "STG") and op.split(".")[0] == k):
                c[k] += 1
                break
    pretty = re.sub(r"\(anonymous namespace\)::", "", demangle(name))
    pretty = re.sub(r"\(.*", "", pretty).replace("void ", "")
    r = regs.get(name, (0, 0))
    print(f"{pretty} | {r[0]} | {r[1]} | {len(ops)} | " + " ".join(str(c[k]) for k in KEYS))
