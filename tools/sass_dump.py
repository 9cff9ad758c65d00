#!/usr/bin/env python
"""Refresh profiles/sass/<name>.sass for the kernels that changed this round (cuobjdump of the in-tree
objects, encodings stripped) and print the README entries:  python tools/sass_dump.py"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "lj_gpu_b200", "csrc", "build")
WANT = [  # (file, object, regex on the demangled name, description)
    ("celltile_fp64_double4", "lj_force_celltile.o", r"lj_celltile_force<1, false, 16, 1>", "lj_celltile_force<LJ_AOS_D4, MX=false, 16 consumers, 1 CTA/SM>: AUTO at N >= 3e5, FP64"),
    ("celltile_mixed_double4", "lj_force_celltile.o", r"lj_celltile_force<1, true, 16, 1>", "lj_celltile_force<LJ_AOS_D4, MX=true, 16 consumers, 1 CTA/SM>: LJ_PREC_MIXED on the cell-tile mirror"),
    ("tile_engine_count", "lj_nlist.o", r"k_tile_count", "k_tile_count: the one search pass of the list build (FP32 window scan, FP64 recheck in the error band)"),
    ("tile_engine_replay", "lj_nlist.o", r"k_tile_replay<false>", "k_tile_replay<32-bit pointers>: masks -> mirror list + public list, no second search"),
    ("newton3_ell_double4", "lj_force.o", r"lj_newton3_ell<1>", "lj_newton3_ell<LJ_AOS_D4>: Newton-3 half list in the column-major ELL layout"),
    ("gather_ellrows_g8_double4", "lj_force.o", r"lj_gather_ellrows<8, 1>", "lj_gather_ellrows<8, LJ_AOS_D4>: row-major padded list (reference sorted_list2d)"),
]
for name, obj, rx, desc in WANT:
    out = subprocess.run("cuobjdump -sass %s | c++filt" % os.path.join(OBJ, obj), shell=True, capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    hit = [b for b in blocks[1:] if re.search(rx, b.split("\n", 1)[0])]
    if not hit:
        print("!! no match for", name); continue
    b = hit[0]
    body = []
    for l in b.split("\n")[1:]:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m: body.append("/*%s*/  %s ;" % (m.group(1), m.group(2).strip()))
    with open(os.path.join(ROOT, "profiles", "sass", name + ".sass"), "w") as f:
        f.write("// %s\n// %s\n" % (desc, b.split("\n", 1)[0][:200]) + "\n".join(body) + "\n")
    ops = collections.Counter()
    for l in body:
        t = l.split("*/", 1)[1].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    keys = ["DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "I2FP", "MUFU", "LDS", "STS", "LDG", "STG", "REDG", "ATOMS", "SHFL", "UBLKCP", "SYNCS", "VOTE", "POPC", "LOP3", "IMAD", "ISETP", "BRA"]
    print("%s.sass  (%d instructions)\n  %s\n  %s\n" % (name, len(body), desc, "  ".join("%s=%d" % (k, ops[k]) for k in keys if ops[k])))
