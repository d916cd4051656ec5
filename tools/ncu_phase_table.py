"""Per-phase instruction / stall / shared-memory table of k_onchip from an ncu report taken with --import-source on.

The CSV of ncu's source page is SASS-level without line numbers; `nvdisasm -g` of the cubin inside libtno_b200.so gives
the line of every instruction.  Both are joined on the instruction offset and aggregated over the line ranges of the
numbered phases of the kernel.

usage: python tools/ncu_phase_table.py report.ncu-rep candidates_per_launch [top_lines]
(top_lines > 0 also lists the source lines with the most stall samples)
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
FUN = "_ZN3tno8k_onchipILi256ELi2ELi1ELb0ELb0E"


def main():
    rep, per = sys.argv[1], float(sys.argv[2])
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "compas_tno_b200", "libtno_b200.so")], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith("tno_onchip.") and f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True, check=True).stdout
    lines, cur, infun = {}, None, False
    for ln in sass.splitlines():
        if ln.startswith(".text." + FUN):
            infun = True
            continue
        if infun and ln.startswith("//---------------------"):
            break
        if not infun:
            continue
        m = re.search(r'//## File "(.*)", line (\d+)', ln)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith("tno_onchip.cu") else None
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    base = int(data[0][0], 16)
    src = open(os.path.join(ROOT, "compas_tno_b200", "csrc", "tno_onchip.cu")).read().splitlines()

    def find(s):
        for i, t in enumerate(src):
            if s in t:
                return i + 1
        raise KeyError(s)

    marks = [("0 variables", find("// ---- 0. variables")), ("1 q = B q_ind + d, funicular rows of g", find("// ---- 1. q = B q_ind")),
             ("2 assembly", find("// ---- 2. assemble")), ("3 numeric LDL^T: staged pull chunks", find("// ---- 3.")),
             ("4 scale to L", find("// ---- 4. L_ij")), ("5 phase-1 right-hand sides", find("// ---- 5. phase-1 right-hand")),
             ("6 solve 1 (call site)", find("// ---- 6. solve phase 1")), ("7 post-1: z, envelope g, reactions", find("// ---- 7. everything")),
             ("8 phase-2 right-hand sides (edge colours)", find("// ---- 8. phase-2 right-hand")),
             ("9 solve 2 (call site)", find("// ---- 9. solve phase 2")), ("10 reac rows, objective, gradient", find("// ---- 10.")),
             ("11 J envelope rows", find("// ---- 11.")), ("12 gmin / status", find("// ---- 12."))]
    kstart = find("__global__ void __launch_bounds__(THREADS * G, MINB) k_onchip")
    dense0 = find("// Dense work of one chain level")
    solve0 = find("// ---- kernels of the triangular solves")
    items0 = find("// Pull-form items of one staged piece of the factorisation")
    items1 = find("// Register-tile variant of the dense work of one chain level")

    def phase(line):
        if line is None:
            return "inlined library code (no line)"
        if items0 <= line < items1:
            return "3 numeric LDL^T: staged pull chunks"
        if solve0 <= line < dense0:
            return "6 + 9 triangular solves (pull / push / chain apply / gather)"
        if dense0 <= line < kstart:
            return "3 dense chain blocks (LDL^T + inverse)"
        if line < kstart:
            return "helpers (envelope closed forms, cp.async)"
        p = "prologue (index region load)"
        for name, l0 in marks:
            if line >= l0:
                p = name
        return p

    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(collections.Counter)
    for r in data:
        li = lines.get(int(r[0], 16) - base)
        if li is None or li[1].split()[0] not in r[1]:
            raise SystemExit("the report does not match the library's code: rebuild the state that was profiled")
        a = agg[phase(li[0])]
        a["instr"] += int(r[col["Instructions Executed"]] or 0)
        a["samples"] += int(r[col["# Samples"]] or 0)
        a["smem"] += int(r[col["L1 Wavefronts Shared"]] or 0)
        a["excess"] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        for sc in stalls:
            a[sc] += int(r[col[sc]] or 0)
    if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
        byline = collections.defaultdict(collections.Counter)
        for r in data:
            li = lines.get(int(r[0], 16) - base)
            a = byline[li[0]]
            a["instr"] += int(r[col["Instructions Executed"]] or 0)
            a["samples"] += int(r[col["# Samples"]] or 0)
            a["smem"] += int(r[col["L1 Wavefronts Shared"]] or 0)
            for sc in stalls:
                a[sc] += int(r[col[sc]] or 0)
        tot = sum(a["samples"] for a in byline.values())
        print("| line | source | stall samples | warp instr / candidate | smem wavefronts / candidate | top stall reasons |")
        print("|---|---|---|---|---|---|")
        for ln, a in sorted(byline.items(), key=lambda kv: -kv[1]["samples"])[:int(sys.argv[3])]:
            top = sorted(((sc, a[sc]) for sc in stalls), key=lambda kv: -kv[1])[:3]
            text = src[ln - 1].strip()[:90] if ln else "(no line)"
            print("| %s | `%s` | %.2f %% | %.2f k | %.2f k | %s |" % (
                ln, text.replace("|", "\\|"), 100 * a["samples"] / tot, a["instr"] / per / 1e3, a["smem"] / per / 1e3,
                ", ".join("%s %.0f %%" % (k.replace("stall_", ""), 100 * v / max(1, a["samples"])) for k, v in top)))
        print()
    ti = sum(a["instr"] for a in agg.values())
    ts = sum(a["samples"] for a in agg.values())
    print("| phase | warp instr / candidate | share | stall samples | smem wavefronts / candidate (excess) | top stall reasons |")
    print("|---|---|---|---|---|---|")
    for ph, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        if a["samples"] < 0.004 * ts:
            continue
        top = sorted(((sc, a[sc]) for sc in stalls), key=lambda kv: -kv[1])[:3]
        print("| %s | %.1f k | %.1f %% | %.1f %% | %.1f k (%.1f k) | %s |" % (
            ph, a["instr"] / per / 1e3, 100 * a["instr"] / ti, 100 * a["samples"] / ts, a["smem"] / per / 1e3, a["excess"] / per / 1e3,
            ", ".join("%s %.0f %%" % (k.replace("stall_", ""), 100 * v / max(1, a["samples"])) for k, v in top)))


if __name__ == "__main__":
    main()
