#!/usr/bin/env python
"""tools/sass_count.py OBJ KERNEL_SUBSTR [--dump] — instructions per attempt on the HOT PATH of a kernel's main loop.

The hot path is found statically from `cuobjdump -sass`: the main loop is the backward branch spanning the most FP64
instructions; the hot path is the SHORTEST trip from the loop head back to it among the trips that hold (within 10 %)
the most FP64 instructions — rare blocks (retire/refill, exact controller tests, checkpoints, regrouping) only add
instructions, trips that skip the stages hold next to no FP64.  Prints FP64 / other instruction counts for one trip round the loop (the RK
fast kernels make TWO attempts per trip: drive.cuh) and, with --dump, the hot path itself.
"""
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def load(obj, kernel):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funs = re.findall(r"Function : (\S+)", names)
    match = [f for f in funs if all(k in f for k in kernel.split("+"))]
    if not match:
        sys.exit(f"no kernel matching {kernel!r}")
    fun = match[0]
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
    ins = []
    for l in out.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return fun, ins


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]


def main():
    obj, kernel = sys.argv[1], sys.argv[2]
    dump = "--dump" in sys.argv
    fun, ins = load(obj, kernel)
    addr = {a: i for i, (a, _) in enumerate(ins)}
    isfp = [opcode(t) in FP64 for _, t in ins]
    pre = [0]
    for f in isfp:
        pre.append(pre[-1] + f)
    bra = re.compile(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)")
    # main loop: the backward branch whose span holds the most FP64 instructions
    best = None
    for i, (a, t) in enumerate(ins):
        m = bra.search(t)
        if m and int(m.group(1), 16) in addr and int(m.group(1), 16) < a:
            j = addr[int(m.group(1), 16)]
            n = pre[i + 1] - pre[j]
            if best is None or n > best[0]:
                best = (n, j, i)
    _, head, tail = best

    def succ(i):
        t = ins[i][1]
        op = opcode(t)
        if op in ("EXIT", "RET") and not t.startswith("@"):
            return []
        m = bra.search(t) if op == "BRA" else None
        out = []
        if m and int(m.group(1), 16) in addr:
            out.append(addr[int(m.group(1), 16)])
            cond = t.startswith("@") or re.search(r"BRA(?:\.\w+)*\s+!?U?P\d+,", t)
            if cond and i + 1 < len(ins):
                out.append(i + 1)
        elif i + 1 < len(ins):
            out.append(i + 1)
        return out

    # Hot path = the SHORTEST trip head -> head among those that hold (nearly) the most FP64 instructions: the rare
    # blocks (exact controller tests, retire/refill, checkpoints, regrouping) only ever add instructions, and the trips
    # that skip the stages (Done, attempt cap) hold next to no FP64.  Pareto table per instruction: {fp64: (len, next)}.
    sys.setrecursionlimit(100000)
    memo, on_stack = {}, set()

    def table(i):
        if i in memo:
            return memo[i]
        if i in on_stack:
            return {}
        on_stack.add(i)
        res = {}
        for j in succ(i):
            sub = {0: (0, None)} if j == head else table(j)
            for fp, (ln, _) in sub.items():
                key = fp + isfp[i]
                if key not in res or ln + 1 < res[key][0]:
                    res[key] = (ln + 1, j)
        on_stack.discard(i)
        memo[i] = res
        return res

    tab = table(head)
    top = max(tab)
    fp_hot = min((fp for fp in tab if fp >= 0.9 * top), key=lambda fp: tab[fp][0])
    path, i, fp = [], head, fp_hot
    while True:
        path.append(i)
        j = memo[i][fp][1]
        fp -= isfp[i]
        if j == head or j is None:
            break
        i = j
    ops = [opcode(ins[k][1]) for k in path]
    n_fp = sum(o in FP64 for o in ops)
    by = {}
    for o in ops:
        by[o] = by.get(o, 0) + 1
    print(f"kernel: {fun}")
    print(f"main loop: {ins[head][0]:#06x} .. {ins[tail][0]:#06x}; hot path = {len(path)} instructions per trip: "
          f"{n_fp} FP64 + {len(path) - n_fp} other")
    print("by opcode: " + ", ".join(f"{k} {v}" for k, v in sorted(by.items(), key=lambda kv: -kv[1])))
    if dump:
        for k in path:
            print(f"  {ins[k][0]:04x}  {ins[k][1]}")


if __name__ == "__main__":
    main()
