"""Per-object SASS opcode census of the built library (cuobjdump -sass on cuda-qr_b200/csrc/*.o): the mnemonics that prove
which hardware paths a kernel uses (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP /
UBLKPF, mma.sync -> HMMA, packed fp32 -> FFMA2).   python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cuda-qr_b200", "csrc")
ops = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "HMMA", "FFMA2", "FMUL2", "FFMA", "DFMA", "SHFL",
       "MUFU", "SYNCS", "LDGSTS", "BAR", "UCGABAR", "ATOM", "RED"]
print("# cuobjdump -sass opcode counts per object (static instruction counts), nvcc -gencode arch=compute_100a,code=sm_100a")
print(f"{'object':22s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in ops))
for obj in sorted(glob.glob(os.path.join(root, "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    total = 0
    for line in txt.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
        if not m:
            continue
        total += 1
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + ".") or (o in ("BAR",) and op == "BAR"):
                cnt[o] += 1
        base = op.split(".")[0]
        if base in ops and base != op:
            pass
    # opcode field includes modifiers after the first dot: count by base mnemonic
    cnt = collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
        if m and m.group(1) in ops:
            cnt[m.group(1)] += 1
    print(f"{os.path.basename(obj):22s} {total:7d} " + " ".join(f"{cnt[o]:8d}" for o in ops))
print("\n# kernels per object with a tensor / TMA / bulk-copy instruction")
for obj in sorted(glob.glob(os.path.join(root, "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur = None
    per = collections.defaultdict(collections.Counter)
    for line in txt.splitlines():
        f = re.match(r"\s+Function : (\S+)", line)
        if f:
            cur = f.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
        if m and cur and m.group(1) in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "HMMA", "UBLKPF", "UBLKCP", "UTMAPF"):
            per[cur][m.group(1)] += 1
    for fn, c in per.items():
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        print(f"{os.path.basename(obj):18s} {name[:110]:110s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
