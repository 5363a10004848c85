"""Compile every .cu of the library with `-Xptxas -v` (no GPU needed) and tabulate registers / spills / static shared
memory per kernel, plus a census of the SASS mnemonics that prove the Blackwell paths (UTCHMMA = tcgen05.mma,
UBLKCP = cp.async.bulk, UTCBAR / SYNCS = tcgen05.commit / mbarrier, LDTM / STTM = tcgen05.ld / st)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dcl_net_b200", "csrc")
files = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
print("| file | kernel | registers | spill st/ld (B) | static smem (B) |")
print("|---|---|---:|---:|---:|")
for f in files:
    out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                          "--expt-relaxed-constexpr", "-Xptxas", "-v", "-c", f, "-o", "/tmp/_ptxas_report.o"],
                         cwd=CSRC, capture_output=True, text=True).stderr
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\n.*?Used (\d+) registers[^\n]*?(?:, (\d+) bytes smem)?\n", out):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
        print(f"| {f} | `{name}` | {m.group(5)} | {m.group(3)}/{m.group(4)} | {m.group(6) or 0} |")
lib = os.path.join(ROOT, "dcl_net_b200", "libdcl_b200.so")
if os.path.exists(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, census = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", re.sub(r"\(anonymous namespace\)::", "", fn)).replace("void ", "")
            continue
        for mn in ("UTCHMMA", "UBLKCP", "UTCBAR", "LDTM", "STTM", "SYNCS", "UCGABAR", "REDUX", "MATCH"):
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                census.setdefault(fn, {}).setdefault(mn, 0)
                census[fn][mn] += 1
    print("\n| kernel | Blackwell-path SASS mnemonics (count) |")
    print("|---|---|")
    for fn in sorted(census):
        print(f"| `{fn}` | " + ", ".join(f"{k} {v}" for k, v in sorted(census[fn].items())) + " |")
