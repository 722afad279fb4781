"""Static SASS instruction mix of the fused kernels, from the built objects (no GPU needed):
    python tools/sass_mix.py > profiles/r2_sass_mix.txt
Per kernel: packed-fp32 instructions split by how many DISTINCT general registers they read (the register file, not the
fma pipe, sets their cost: tools/ubench/fp32x2_operands.cu), shared/global memory instructions, and the async-proxy / TMA
mnemonics that prove the Blackwell path (UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = ["fast_path.o", "sf_fast.o", "tacaw_fast.o"]
FP2 = ("FFMA2", "FADD2", "FMUL2")
# measured issue cost in cycles per warp-instruction and scheduler (profiles/r2_ubench_fp32x2_operands.txt): the register
# file feeds ~two 32-bit operands per lane and cycle, so a packed instruction costs max(2, reads / 2) + ~6 %
COST = {4: 2.13, 5: 2.68, 6: 3.24}
PIXELS_PER_THREAD = 16       # every fused kernel: 16 points of a line per thread and tile


def classify(op, operands):
    srcs = [a.strip() for a in operands.split(",")][1:]
    regs, other = set(), []
    for a in srcs:
        m = re.match(r"-?\|?(U?R\d+)", a)
        if m and not m.group(1).startswith("UR"):
            regs.add(m.group(1))
        elif m:
            other.append("uniform")
        elif a.startswith("c[") or a.startswith("-c["):
            other.append("const")
        else:
            other.append("imm")
    # .F32 = one 32-bit register broadcast to both halves, otherwise a 64-bit pair
    n32 = sum(1 if re.search(re.escape(r) + r"(\.reuse)?\.F32(\W|$)", operands) and ".F32x2" not in operands.split(r)[1][:12] else 2 for r in regs)
    return f"{op} {len(regs)} reg src" + (f"+{'/'.join(sorted(set(other)))}" if other else "") + f" ({n32} x 32-bit reads)"


def cost_of(label):
    reads = int(re.search(r"\((\d+) x 32-bit", label).group(1))
    return COST.get(reads, 2.13 if reads < 4 else 3.24)


def main():
    for obj in OBJS:
        path = os.path.join(ROOT, "pyslice_b200", "csrc", "build", obj)
        txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        for f in re.split(r"\n\s*Function : ", txt)[1:]:
            name = f.split("\n")[0]
            short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            short = re.sub(r"psb::\(anonymous namespace\)::", "", short).split("(")[0]
            ops, fp2 = collections.Counter(), collections.Counter()
            for line in f.split("\n"):
                m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?\s*(.*?);", line)
                if not m:
                    continue
                op = m.group(1)
                ops[op] += 1
                if op in FP2:
                    fp2[classify(op, m.group(3))] += 1
            total = sum(ops.values())
            fp = sum(ops[o] for o in FP2)
            mem = {k: ops[k] for k in ("LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMS", "RED") if ops[k]}
            tma = {k: ops[k] for k in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMAPF", "BAR", "MUFU") if ops[k]}
            print(f"{short}\n    {total} instructions: packed fp32 {fp} ({100 * fp // max(total, 1)} %)  scalar fp32 "
                  f"{ops['FFMA'] + ops['FADD'] + ops['FMUL']}  memory {mem}  async/sync {tma}")
            for k, v in sorted(fp2.items()):
                print(f"        {v:5d}  {k}")
            if "fast_" in short and fp:
                # one trip of the persistent loop = the whole listing minus a short prologue: register-file-limited fp32
                # issue time of a tile, per pixel and SM (4 schedulers), scalar fp32 at ~1.5 cycles
                cyc = sum(v * cost_of(k) for k, v in fp2.items()) + 1.5 * (ops["FFMA"] + ops["FADD"] + ops["FMUL"])
                print(f"        -> fp32 issue floor {cyc:7.0f} cycles per warp and tile = {cyc / (PIXELS_PER_THREAD * 32 * 4):.3f} "
                      f"SM-cycles per pixel (idealised 2 cycles per packed instruction: {2 * fp / (PIXELS_PER_THREAD * 32 * 4):.3f})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
