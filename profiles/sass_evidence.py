"""Per-kernel SASS mnemonic counts of libwifi_b200.so (runs here, no GPU needed).
    python profiles/sass_evidence.py > profiles/rNN_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "80211parallelestimation_b200", "libwifi_b200.so")
KEEP = ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UBLKPF", "SYNCS", "DMMA", "HMMA", "FFMA2", "FADD2", "FFMA", "DFMA", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "MUFU", "CREDUX", "REDUX", "LDL", "STL")

HEADER = """SASS evidence (cuobjdump -sass libwifi_b200.so, sm_100a): per kernel, counts of the mnemonics that show which hardware path it uses.
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (tensor memory), UBLKCP = cp.async.bulk (TMA bulk copy), UBLKPF = cp.async.bulk.prefetch.L2,
SYNCS = mbarrier, DMMA = FP64 tensor-core mma.sync, HMMA = warp-level mma.sync (here: TF32 m16n8k8 of the batched inverse), FFMA2 = fma.rn.f32x2 (packed FP32 pairs: the low-rank per-frame MMSE).  (B200_PROFILING.md: the PTX names never appear in SASS.)
"""


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    print(HEADER)
    for name, blk in zip(names, blocks):
        ops = collections.Counter()
        n = 0
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", blk, re.M):
            n += 1
            op = m.group(1)
            if op in KEEP:
                ops[op] += 1
        short = re.sub(r"\(.*", "", name)
        print("%-72s %5d instr | %s" % (short[:72], n, ", ".join("%s %d" % kv for kv in ops.most_common(8))))


if __name__ == "__main__":
    sys.exit(main())
