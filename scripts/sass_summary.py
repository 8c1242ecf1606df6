#!/usr/bin/env python3
"""SASS and ptxas evidence for profiles/: which Blackwell-native instructions each kernel family of the shipped
liblcc_b200.so contains (cuobjdump -sass, counted per kernel family), and registers / spills per kernel from the saved
nvcc -Xptxas -v logs.  Runs in the build container (no GPU needed).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "largecrimsoncanine_b200", "lib", "liblcc_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMACCTL", "UBLKCP", "SYNCS", "DMMA", "HMMA", "FFMA", "DFMA", "LDG", "STG",
             "LDS", "STS", "LDL", "STL", "BAR", "REDUX", "SHFL"]
FAMILIES = [("K1 product (direct)", r"product_kernel(?!_aos_tma|_soa_tma)"), ("K1 product (AoS TMA tiles)", r"product_kernel_aos_tma"),
            ("K1 product (SoA TMA tiles)", r"product_kernel_soa_tma"), ("K1+K5 fused product+sum", r"product_sum_kernel"),
            ("K2 sandwich (direct)", r"sandwich_kernel(?!_aos_tma|_soa_tma)"), ("K2 sandwich (AoS TMA)", r"sandwich_kernel_aos_tma"),
            ("K2 sandwich (SoA TMA)", r"sandwich_kernel_soa_tma"), ("K2c PGA points", r"pga_points_sandwich_kernel"),
            ("K3 blade maps / addsub", r"blade_map|addsub"), ("K4 norms", r"norm"), ("K5 sums", r"sum_partial|sum_final"),
            ("K6 tcgen05 GEMM (CTA pairs)", r"fixed_product_tf32x3_pair_kernel"), ("K6 tcgen05 GEMM (single CTA)", r"fixed_product_tf32x3_kernel"),
            ("K6d DMMA GEMM", r"fixed_product_f64_dmma_kernel"), ("K7 TMA transpose", r"transpose_tma_kernel"), ("K7 scalar transpose", r"transpose_kernel"),
            ("K8 matrix representation", r"matrix_rep_kernel"), ("table product", r"table_product_kernel"),
            ("matrix builders", r"build_.*matrix|split_w|write_rotor"), ("points embed/extract, columns", r"points_|columns_kernel")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    kernels = collections.Counter()
    fam = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            fam = next((f for f, pat in FAMILIES if re.search(pat, name)), "other")
            kernels[fam] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fam:
            op = m.group(1)
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[fam][mn] += 1
                    break
    print("# SASS mnemonic counts per kernel family in largecrimsoncanine_b200/lib/liblcc_b200.so (cuobjdump -sass; sm_100a)")
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, SYNCS = mbarrier, DMMA = FP64 tensor MMA,")
    print("# HMMA = warp-level mma.sync (TF32 in K8)")
    total = collections.Counter()
    for f, _ in FAMILIES + [("other", "")]:
        if not kernels[f]:
            continue
        c = counts[f]
        total.update(c)
        print(f"{f:34s} kernels={kernels[f]:4d}  " + "  ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn]))
    print(f"{'TOTAL':34s} kernels={sum(kernels.values()):4d}  " + "  ".join(f"{mn}={total[mn]}" for mn in MNEMONICS if total[mn]))
    from largecrimsoncanine_b200 import _build

    print("\n# ptxas -v: registers / spills per kernel (kernels with spills, and the heaviest 25 by register count)")
    rows = []
    for ln in _build.ptxas_report().splitlines():
        m = re.search(r"^(\S+) (\S+) \| .*Used (\d+) registers.*?\| ?(.*)$", ln)
        if m:
            spill = re.search(r"(\d+) bytes spill stores", m.group(4))
            rows.append((int(m.group(3)), int(spill.group(1)) if spill else 0, m.group(1), m.group(2)))
    for regs, spill, tu, name in sorted(rows, key=lambda r: (-r[1], -r[0])):
        if spill:
            print(f"spill {spill:5d} B  regs {regs:3d}  {tu}  {name[:110]}")
    for regs, spill, tu, name in sorted(rows, key=lambda r: -r[0])[:25]:
        print(f"regs {regs:3d}  spill {spill:5d} B  {tu}  {name[:110]}")
    print(f"# {len(rows)} kernels in total, {sum(1 for r in rows if r[1])} with spills")


if __name__ == "__main__":
    main()
