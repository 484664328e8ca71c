"""Verbose parity report on the GPU (debug aid): python tests/gpu_report.py [sac|sacper|td3|big]"""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_utils as PU

which = sys.argv[1:] or ["sac"]
for w in which:
    print("=" * 30, w)
    try:
        if w == "sac":
            r = PU.sac_parity_report(S=24, A=2, B=32, C=64, steps=2, seed=7, verbose=True)
        elif w == "sacper":
            r = PU.sac_parity_report(S=24, A=2, B=32, C=64, steps=3, seed=8, prioritized=True, verbose=True)
        elif w == "sachint":
            r = PU.sac_parity_report(S=22, A=2, B=32, C=64, steps=12, seed=5, use_hint=True, verbose=True, trace=False)
        elif w == "td3":
            r = PU.td3_parity_report(S=24, A=2, B=32, C=64, steps=4, seed=9, prioritized=True, verbose=True)
        elif w == "td3hint":
            r = PU.td3_parity_report(S=24, A=2, B=32, C=64, steps=4, seed=13, prioritized=True, use_hint=True, admm_rho=1.0, verbose=True)
        elif w == "big":
            r = PU.sac_parity_report(S=420, A=2, B=256, C=1024, steps=2, seed=3, verbose=True)
        elif w == "bigper":
            r = PU.sac_parity_report(S=420, A=2, B=256, C=4096, steps=2, seed=4, prioritized=True, verbose=True)
        print("RESULT", w, "ok" if r["ok"] else "FAILED", "failures:", len(r["failures"]), "launches", r["launches"])
        for f in r["failures"][:12]:
            print("   ", f)
    except Exception:
        traceback.print_exc()
