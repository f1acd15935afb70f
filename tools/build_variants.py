"""A/B builds of the board sweep kernel: python tools/build_variants.py
Compiles csrc/cfr_board.cu with different PRL_BV_* switches into pokerrl_b200/lib/variants/lib_<name>.so (the other
translation units are taken from the regular build); run one with PRL_LIB_PATH=<that file> python tools/board_probe.py."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pokerrl_b200.csrc import build as B  # noqa: E402

VARIANTS = {
    "base": dict(RED=0, P1PIPE=0, FOLDLIN=0, SCAN7=0, SPLITB3=0),
    "red": dict(RED=1, P1PIPE=0, FOLDLIN=0, SCAN7=0, SPLITB3=0),
    "p1pipe": dict(RED=0, P1PIPE=1, FOLDLIN=0, SCAN7=0, SPLITB3=0),
    "foldlin": dict(RED=0, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=0),
    "noscan": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=0),
    "scan7": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=1, SPLITB3=0),
    "splitb3": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=1),
    "ert": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=0, ERT=1),
    "splitb3_ert": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=1, ERT=1),
    "final": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=0, ERT=1),
    "rowtotf": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=0, ERT=1, ROWTOTF=1),
    "nonewton": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=0, ERT=1, NEWTON=0),
    "ert_p1a2": dict(RED=1, P1PIPE=2, FOLDLIN=1, SCAN7=0, SPLITB3=0, ERT=1),
    "p3bal": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=0, P3BAL=1),
    "splitb3_ert_p3bal": dict(RED=1, P1PIPE=1, FOLDLIN=1, SCAN7=0, SPLITB3=1, ERT=1, P3BAL=1),
}


def main(names):
    B.build()
    out_dir = os.path.join(B.LIB_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    obj_dir = os.path.join(B.LIB_DIR, "obj")
    others = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s in B.SOURCES if s != "cfr_board.cu"]
    for name in names:
        defs = ["-DPRL_BV_%s=%d" % kv for kv in VARIANTS[name].items()]
        o = os.path.join(out_dir, "cfr_board_%s.o" % name)
        subprocess.check_call([B.NVCC] + B.ARCH + B.COMMON + B.SOURCES["cfr_board.cu"] + defs +
                              ["-c", os.path.join(B.HERE, "cfr_board.cu"), "-o", o])
        lib = os.path.join(out_dir, "lib_%s.so" % name)
        subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", lib, o] + others)
        print(lib)


if __name__ == "__main__":
    main(sys.argv[1:] or list(VARIANTS))
