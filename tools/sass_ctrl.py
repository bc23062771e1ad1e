"""Decode the scheduling control fields of sm_100 SASS (cuobjdump -sass output): stall count, yield, write / read
scoreboard index and the wait mask of every instruction. Development aid for reading which loads share a scoreboard.

    cuobjdump -sass lib.so | python tools/sass_ctrl.py [regex-of-opcodes-to-show]
"""
import re
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else "."
lines = sys.stdin.read().split("\n")
i = 0
out = []
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = hi >> 41   # bits 105.. of the 128-bit word
            stall = ctrl & 0xf
            yld = (ctrl >> 4) & 1
            wb = (ctrl >> 5) & 7
            rb = (ctrl >> 8) & 7
            wait = (ctrl >> 11) & 0x3f
            txt = m.group(2).strip()
            if re.search(pat, txt):
                w = "".join(str(k) for k in range(6) if wait >> k & 1)
                print(f"{m.group(1)} st{stall:2d} {'Y' if yld else ' '} W{wb if wb != 7 else '-'} R{rb if rb != 7 else '-'} wait[{w:6s}] {txt}")
            i += 2
            continue
    i += 1
