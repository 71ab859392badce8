"""Static report of the built library (no GPU needed): registers / stack / shared memory per kernel from cuobjdump, and the
instruction mix of one kernel or of one noinline device function inside it.

    python scripts/sass_report.py                         # table of all kernels
    python scripts/sass_report.py k_query3ILi1E tiledILi0ELi3ELi1E   # mix of the K_BB / floor_k=3 / hybrid decade-tile body
"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SO = os.environ.get("RBF_B200_LIB") or os.path.join(ROOT, "new_bloom_filter_repo_b200", "librbf_b200.so")


def usage_table():
    out = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout.splitlines()
    rows = []
    for i, l in enumerate(out):
        m = re.match(r"\s*Function (\S+):", l)
        if m and i + 1 < len(out):
            f = dict(kv.split(":") for kv in out[i + 1].split() if ":" in kv)
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            rows.append((name, f.get("REG"), f.get("STACK"), f.get("SHARED")))
    w = max(len(r[0]) for r in rows)
    print("%-*s  regs stack  smem" % (w, "kernel"))
    for r in sorted(rows):
        print("%-*s  %4s %5s %5s" % (w, *r))


def mix(kernel_pat, inner_pat=None):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.startswith("rbf_kernels.") and f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()
    take, body = False, []
    for l in sass:
        lab = re.match(r"^(\$?[_A-Za-z0-9$]+):$", l)
        if lab:
            name = lab.group(1)
            if name.startswith(".L"):
                pass
            elif inner_pat:
                take = kernel_pat in name and inner_pat in name
            else:
                take = kernel_pat in name and "$" not in name[1:]
        elif l.startswith(".text.") or l.startswith("\t.section"):
            take = False
        if take:
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
            if m:
                body.append(m.group(1).strip())
    ops = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for t in body)
    print("%d instructions" % len(body))
    for op, c in ops.most_common(24):
        print("  %-10s %5d  %5.1f%%" % (op, c, 100.0 * c / max(1, len(body))))


def census():
    """Library-wide count of the mnemonics that show which hardware paths the kernels use."""
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout.splitlines()
    pats = [("UBLKCP (cp.async.bulk: TMA bulk copy engine, 1-D)", r"\bUBLKCP"), ("UTMALDG / UTMASTG (tensor-map TMA)", r"\bUTMA(LDG|STG)"),
            ("SYNCS (mbarrier)", r"\bSYNCS"), ("REDG (fire-and-forget global atomics)", r"\bREDG"), ("ATOMS (shared-memory atomics)", r"\bATOMS"),
            ("LDG ... .256 (32-byte vector loads)", r"\bLDG\.E[A-Z0-9.]*\.256"), ("VABSDIFF4 (byte-SIMD |a-b|)", r"\bVABSDIFF4"),
            ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("IMAD.WIDE", r"\bIMAD\.WIDE"), ("HMMA/UTCMMA (tensor cores; expected 0)", r"\b(HMMA|UTC\w*MMA|QGMMA|HGMMA)")]
    nk = sum(1 for l in sass if "Function :" in l)
    print("\n%d kernels, sm_100a; mnemonic census over the whole library:" % nk)
    for name, pat in pats:
        print("  %-58s %6d" % (name, sum(1 for l in sass if re.search(pat, l))))


if __name__ == "__main__":
    if len(sys.argv) == 1:
        usage_table()
        census()
    else:
        mix(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
