"""What proves the kernels are Blackwell-native: per kernel, counts of the SASS mnemonics for tcgen05
(`UTC*MMA`, of which `.2CTA`), TMEM loads (`LDTM`), TMA tensor / bulk copies (`UTMALDG`, `UBLKCP`), tcgen05
commits (`UTCBAR`), and the legacy tensor path (`HMMA`, must be 0).  Reads the objects the build left in
perceive_b200/_obj/; writes profiles/<round>_sass_summary.txt.

    python tools/sass_summary.py profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PATS = OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("of which .2CTA", r"\bUTC\w*MMA\S*\.2CTA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"),
                    ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (legacy)", r"\bHMMA"),
                    ("FFMA", r"\bFFMA"), ("LDS.128", r"\bLDS\S*\.128"), ("ST.E (peer/global stores)", r"\bST\.E")])


def main():
    out_path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "sass_summary.txt"
    lines = ["# SASS mnemonic counts per kernel (cuobjdump -sass of perceive_b200/_obj/*.o, sm_100a)", ""]
    for obj in sorted((ROOT / "perceive_b200" / "_obj").glob("*.o")):
        sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
        kernels = re.split(r"\n\s*Function : ", sass)[1:]
        for blk in kernels:
            name = blk.split("\n", 1)[0].strip()
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
            dem = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", ""))
            counts = {k: len(re.findall(p, blk)) for k, p in PATS.items()}
            if not any(counts[k] for k in ("UTC*MMA", "LDTM", "UTMALDG", "UBLKCP")) and "scan_kernel" not in dem and "rescore" not in dem and "p2p" not in dem:
                continue
            lines.append(f"{obj.name}: {dem}")
            lines.append("    " + "  ".join(f"{k}={v}" for k, v in counts.items() if v or k in ("UTC*MMA", "HMMA (legacy)")))
    out_path.write_text("\n".join(lines) + "\n")
    print(f"wrote {out_path} ({len(lines)} lines)")


if __name__ == "__main__":
    main()
