"""Development tool (not a test): per-opcode instruction histogram of the hottest loop of a
kernel, read from `cuobjdump -sass`.  Usage:
    python tests/dev_sass.py simulst_b200/csrc/build/mma_fwd_bf16.o 'mma_fwd_kernelILi128ELi8E13__nv_bfloat16Li1ELb1E'
The "loop" is the largest backward-branch span (the target-step loop of the MMA kernels)."""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    cand = [f for f in funcs if pat in f]
    assert cand, f"no function matching {pat}"
    fn = cand[0]
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    # find backward branches
    best = None
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    print(" backward branches:", ", ".join(f"{a:#x}..{b:#x} ({(b - a) // 16 + 1})" for a, b in loops))
    if len(sys.argv) > 3:
        best = loops[int(sys.argv[3])]
    print(f"{fn}\n total instructions {len(ins)}; loop {best[0]:#x}..{best[1]:#x}")
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    hist = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        hist[op.split(".")[0]] += 1
    print(f" loop body: {len(body)} instructions")
    for op, c in hist.most_common():
        print(f"   {op:12s} {c}")


if __name__ == "__main__":
    main()
