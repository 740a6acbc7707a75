"""Instruction mix of a kernel's innermost pair loop from its SASS (no GPU needed):
   python scripts/sass_mix.py <object.o> <kernel-name-substring> [occurrence]
The loop is the shortest backward-branch region that contains MUFU.EX2; opcodes are binned by issue pipe
(B300_MICROARCH.md: FFMA/FMUL/FADD/IMAD and the packed fp32x2 forms -> fma pipe; IADD3/LOP3/SHF/ISETP/FSETP/FSEL/
SEL/MOV/FMNMX/PRMT -> alu pipe; MUFU/I2F/F2I/FRND -> xu; LDG/STG/RED/LDS/STS/ATOM/SHFL -> lsu)."""
import re
import subprocess
import sys

obj, name = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
body = [f for f in funcs if name in f.split("\n")[0]][0]
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for i, (addr, text) in enumerate(ins):
    m = re.search(r"BRA\s+(?:\.U\s+)?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) < addr:
        tgt = int(m.group(1), 16)
        j = next(k for k, (a, _) in enumerate(ins) if a >= tgt)
        seg = ins[j:i + 1]
        if any("MUFU.EX2" in t for _, t in seg):
            loops.append(seg)
loops.sort(key=len)
seg = loops[occ]
PIPE = [("xu", r"^(MUFU|I2F|F2I|FRND|I2FP|F2FP)"), ("lsu", r"^(LDG|STG|RED|REDG|LDS|STS|ATOM|SHFL|LDC|LDCU|REDUX|VOTE|MATCH)"),
        ("fma", r"^(FFMA2?|FMUL2?|FADD2?|IMAD|HFMA2|DFMA)"), ("alu", r"^(IADD3?|VIADD|LOP3|SHF|ISETP|FSETP|FSEL|SEL|MOV|FMNMX|PRMT|LEA|CS2R|PLOP3|VIMNMX|IABS|POPC|FLO|BREV)"),
        ("ctl", r"^(BRA|BSSY|BSYNC|BREAK|EXIT|WARPSYNC|NOP|BAR|CALL|RET)")]
mix, ops = {}, {}
for _, t in seg:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0]
    pipe = next((p for p, rx in PIPE if re.match(rx, op)), "other")
    mix[pipe] = mix.get(pipe, 0) + 1
    ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
print(f"{name}: innermost EX2 loop #{occ}: {len(seg)} instructions (per 4-pixel chunk)")
print("  by pipe:", ", ".join(f"{k} {v}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])))
print("  by opcode:", ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))
