"""Time of the re-orthogonalised recurrence (K6) against the plain one, per KCRN_REORTH_STAGE setting (child processes)."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
    import conftest  # noqa
    import torch
    from kcrn_b200.engine import Engine
    from kcrn_b200.synth import make_problem
    from oracle import oracle as O
    name, m = sys.argv[2], int(sys.argv[3])
    P = make_problem(name, seed=0)
    eng = Engine(P.indptr, P.indices, P.data, O.map_labels(P.b), (P.n, P.d))
    eng.set_x(torch.as_tensor(P.x0, device="cuda"))
    eng.gradient(keep_internal=True)

    def rec_ms(ro, reps=3):
        eng.lanczos(m, ro); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                eng.lanczos(m, ro)
            b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / reps)
        return best
    t0, t1 = rec_ms(False), rec_ms(True)
    vbytes = 3 * 8 * P.d * sum(j + 1 for j in range(m - 1))
    print(f"{name} m={m} KCRN_REORTH_STAGE={os.environ.get('KCRN_REORTH_STAGE', 'default')}: recurrence {t0:.3f} ms, re-orthogonalised {t1:.3f} ms, "
          f"K6 {t1 - t0:.3f} ms = {vbytes / ((t1 - t0) * 1e-3) / 1e9:.0f} GB/s of 3 basis reads per step")
    eng.close()
else:
    for name, m in (("news20", 50), ("rcv1", 50), ("rcv1", 10)):
        for kb in ("0", "72", "120", "184"):
            e = dict(os.environ, KCRN_REORTH_STAGE=kb)
            subprocess.run([sys.executable, __file__, "child", name, str(m)], env=e)
