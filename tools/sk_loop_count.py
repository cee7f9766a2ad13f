"""Static size of the dequant loop of wna16_sk_kernel<4,32,EPI_NONE> in an object file (no GPU needed):
instructions from the loop head to its back edge, minus the slow paths of the barrier waits."""
import re, subprocess, sys, collections
obj = sys.argv[1] if len(sys.argv) > 1 else "build/obj/wna16_sk.o"
fun = "_ZN4b2002sk15wna16_sk_kernelILi4ELi32ELi0EEEv14CUtensorMap_stNS0_4ArgsE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2idx = {a: i for i, (a, _) in enumerate(ins)}
sttm = [i for i, (_, t) in enumerate(ins) if "STTM" in t]
# back edge: the last backward BRA after the last STTM whose target lies before the first STTM
head = tail = None
for i in range(sttm[-1], len(ins)):
    m = re.search(r"BRA\s+(?:P\d,\s*)?0x([0-9a-f]+)", ins[i][1])
    if m:
        tgt = int(m.group(1), 16)
        if tgt in addr2idx and addr2idx[tgt] < sttm[0]:
            head, tail = addr2idx[tgt], i
            break
body = ins[head:tail + 1]
# slow-path blocks: from a CS2R (clock read of wait_bar) to the following BSYNC / trap
slow = 0
i = 0
while i < len(body):
    if "CS2R" in body[i][1]:
        j = i
        while j < len(body) and "BPT.TRAP" not in body[j][1]:
            j += 1
        slow += j - i + 1
        i = j + 1
    else:
        i += 1
ops = collections.Counter()
for _, t in body:
    t = re.sub(r"^@!?U?P\d\s+", "", t)
    ops[t.split()[0].split(".")[0]] += 1
print(f"dequant loop: {len(body)} static instructions, ~{len(body) - slow} on the fast path ({slow} in wait slow paths)")
print(ops.most_common(24))
