"""dev: SASS of one kernel with the control fields decoded (stall count, yield, write/read scoreboard, wait mask).
usage: sass_ctrl.py <object or .so> <substring of the mangled kernel name> [from_addr to_addr]"""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
on = False
i = 0
total = 0
while i < len(txt):
    ln = txt[i]
    if "Function :" in ln:
        on = pat in ln
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", ln) if on else None
    if m and i + 1 < len(txt):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", txt[i + 1])
        if m2:
            c = (int(m2.group(1), 16) >> 41) & 0x1FFFFF
            stall, yld, wb, rb, wait = c & 0xF, (c >> 4) & 1, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3F
            a = int(m.group(1), 16)
            if lo <= a <= hi:
                total += stall
                print("%05x st=%2d wb=%s rb=%s wait=%s  %s" % (a, stall, wb if wb != 7 else "-", rb if rb != 7 else "-",
                                                             format(wait, "06b"), m.group(2)))
            i += 2
            continue
    i += 1
print("sum of stall counts in range:", total)
