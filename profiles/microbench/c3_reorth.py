"""K6 where it is bandwidth-bound (bench.py's C3 case: news20 shape, d = 1.36 M, m = 50), per KCRN_REORTH_STAGE setting."""
import json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import bench
    torch.cuda.set_device(0)
    peak, _ = bench.measured_peak_gbs()
    out = bench.strong_case("c3 news20-shape", 19996, 1355191, 450.0, 50, 4, 0, 1, 0, peak)
    print("KCRN_REORTH_STAGE=" + os.environ.get("KCRN_REORTH_STAGE", "default"), json.dumps(out["reorth"]), "us/HVP in step", round(out["us_per_hvp_in_step"], 1))
else:
    for kb in sys.argv[1:] or ("0", "72", "184"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, KCRN_REORTH_STAGE=kb))
