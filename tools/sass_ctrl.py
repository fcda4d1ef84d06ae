"""Developer helper: list the SASS of one kernel with the decoded scheduling control fields (stall count, yield,
write/read scoreboard index, wait mask) -- cuobjdump prints them only as hex.
   python tools/sass_ctrl.py <lib.so> <kernel-name-substring> [first last]"""
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 30
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    on, idx, i = False, 0, 0
    while i < len(out):
        ln = out[i]
        if "Function :" in ln:
            on = pat in ln
            idx = 0
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if on and m:
            m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
            hi64 = int(m2.group(1), 16)
            ctrl = (hi64 >> 41) & 0x1FFFFF
            stall, yld, wbar, rbar, wait = ctrl & 15, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 63
            if lo <= idx <= hi:
                print("%5d %-78s st=%2d %s W=%s R=%s wait=%s" % (idx, m.group(2)[:78], stall, "Y" if yld else " ",
                      "-" if wbar == 7 else wbar, "-" if rbar == 7 else rbar,
                      "".join(str(b) for b in range(6) if wait >> b & 1) or "-"))
            idx += 1
            i += 1
        i += 1


if __name__ == "__main__":
    main()
