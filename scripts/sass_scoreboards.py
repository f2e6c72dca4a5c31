"""Decode the scoreboard fields of a kernel's SASS (cuobjdump -sass): for every global load / reduction
the barrier it sets, and for every instruction the barriers it waits on. Used to check that a
look-ahead load does not share its barrier with the wait of the row it is supposed to run ahead of.
   python scripts/sass_scoreboards.py <mangled-name-substring> [lo_hex hi_hex]"""
import re, subprocess, sys
ROOT = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
so = ROOT + "/node2vec_by_ecc_b200/libn2v_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
name = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
start = next(i for i, l in enumerate(txt) if "Function :" in l and name in l)
end = next((i for i in range(start + 1, len(txt)) if "Function :" in txt[i]), len(txt))
L = txt[start:end]
i = 0
while i < len(L) - 1:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", L[i])
    m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", L[i + 1]) if m else None
    if m and m2:
        a, t, w = int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)
        stall, wr, rd, wait = (w >> 41) & 0xF, (w >> 46) & 7, (w >> 49) & 7, (w >> 52) & 0x3F
        if lo <= a <= hi and (wait or wr != 7 or "BRA" in t):
            ws = "".join(str(b) for b in range(6) if wait >> b & 1)
            print("%05x %-72s set %s  wait [%s]" % (a, t[:72], "-" if wr == 7 else wr, ws))
        i += 2
    else:
        i += 1
