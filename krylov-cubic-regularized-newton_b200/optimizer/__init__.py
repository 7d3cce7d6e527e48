"""Host-side mirror of the reference's ``optimizer`` package for the Krylov-CRN hot path.

Same module and class names as the reference (``optimizer.loss.LogisticRegression``,
``optimizer.cubic.Lanczos / cubic_solver_root / Cubic_Krylov_LS``, ``optimizer.optimizer.Optimizer``,
``optimizer.opt_trace.Trace``) so a driver written against the reference imports unchanged; the
arithmetic runs in libkcrn.so on a B200.
"""
