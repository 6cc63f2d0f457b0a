"""dev: decode the scheduling control bits (stall, yield, write/read barrier, wait mask) of SASS instructions.
usage: cuobjdump -sass file.o | python scripts/sass_ctrl.py <function-substring> [first_addr last_addr]"""
import re
import sys

want = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
infn = False
pend = None
tot = 0
for line in sys.stdin:
    if "Function :" in line:
        infn = want in line
        continue
    if not infn:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line)
    if m:
        pend = (int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16))
        continue
    m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", line)
    if m and pend:
        hiw = int(m.group(1), 16)
        addr, text, _ = pend
        pend = None
        if addr < lo or addr > hi:
            continue
        ctrl = hiw >> 41
        stall = ctrl & 0xf
        yld = (ctrl >> 4) & 1
        wbar = (ctrl >> 5) & 7
        rbar = (ctrl >> 8) & 7
        wait = (ctrl >> 11) & 0x3f
        tot += stall
        print("%05x  st%2d %s w%s r%s wait%02x  %s" % (addr, stall, "Y" if yld else "-", "-" if wbar == 7 else wbar, "-" if rbar == 7 else rbar, wait, text))
print("sum of stall counts:", tot)
