"""Worker for the row-sharded (multi-GPU) parity check; launched by torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/mgpu_worker.py

Every rank holds a contiguous, nnz-balanced row block of ONE problem; results must match the CPU oracle on
the whole matrix (value / gradient / HVP <= 1e-12) and the reference trajectory prefix (<= 1e-8)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import conftest  # noqa: E402,F401  (import paths)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kcrn_b200.dist import join_communicator, shard_rows  # noqa: E402
from kcrn_b200.synth import make_problem  # noqa: E402
from optimizer.cubic import Cubic_Krylov_LS  # noqa: E402
from optimizer.loss import LogisticRegression  # noqa: E402
from oracle import oracle as O  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    import faulthandler
    faulthandler.dump_traceback_later(75, exit=True)      # a hang prints where, instead of eating the box time
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    name = sys.argv[1] if len(sys.argv) > 1 else "rcv1"
    P = make_problem(name, seed=0)
    S = shard_rows(P, rank, world)
    loss = LogisticRegression(S.scipy_csr(), P_labels(S), l1=0, l2=0, store_mat_vec_prod=True, n_global=P.n)
    join_communicator(loss.engine, rank, world)
    L = O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b, threads=4)
    print(f"rank {rank}: shard ready, communicator joined", flush=True)
    rng = np.random.default_rng(5)
    x, v = rng.standard_normal(P.d), rng.standard_normal(P.d)
    xd, vd = torch.as_tensor(x, device="cuda"), torch.as_tensor(v, device="cuda")
    errs = {
        "value": abs(loss.value(xd) - L.value(x)) / abs(L.value(x)),
        "gradient": rel(loss.gradient(xd).cpu().numpy(), L.gradient(x)),
        "hvp": rel(loss.hess_vec_prod(xd, vd).cpu().numpy(), L.hess_vec_prod(x, v)),
    }
    print(f"rank {rank}: oracle methods {errs}", flush=True)
    # replicas must hold identical bits after the allreduce
    h = loss.hess_vec_prod(xd, vd)
    ref = h.clone()
    dist.broadcast(ref, 0)
    errs["replica_bits_equal"] = 0.0 if torch.equal(ref, h) else 1.0
    # fused Lanczos: the epilogue kernel that sums the ranks' partial A't over NVLink peer memory against the NCCL
    # allreduce in front of the same epilogue.  Odd m replayed back to back (no host sync) exercises the
    # read-done acknowledgement between runs; m = 10 with re-orthogonalisation the second epilogue stage.
    eng = loss.engine
    eng.set_x(xd)
    gd = loss.gradient(xd)
    for m, reorth in ((5, False), (10, True), (1, False)):
        got = {}
        for mode in (True, False):
            try:
                eng.p2p_enable(mode)
            except Exception as e:       # peer exchange not attached (no peer access): nothing to compare
                print(f"rank {rank}: peer exchange unavailable ({e})", flush=True)
                break
            for _ in range(12):
                eng.lanczos(m, reorth=reorth, g=gd)
            V, a, bt, info = eng.lanczos_view()
            torch.cuda.synchronize()
            got[mode] = (a.cpu().numpy().copy(), bt.cpu().numpy().copy(), V.cpu().numpy().copy())
            if mode:
                mine = torch.cat([a.flatten(), bt.flatten(), V.flatten()])
                ref0 = mine.clone()
                dist.broadcast(ref0, 0)
                if not torch.equal(ref0, mine):
                    errs["peer_replica_bits_equal"] = 1.0
        if len(got) == 2:
            key = f"peer_vs_nccl_m{m}{'_reorth' if reorth else ''}"
            errs[key] = max(rel(got[True][0], got[False][0]), rel(got[True][2][:2], got[False][2][:2]))
            if errs[key] > 1e-10:
                errs["peer_fail"] = 1.0
    try:
        eng.p2p_enable(True)
    except Exception:
        pass
    opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=1e-9, label="mg", tqdm=False)
    opt.run(P.x0, it_max=8)
    # the step loop kept on the device (nested WHILE graph nodes, peer exchange inside) = the host-driven loop, bit
    # for bit, on every rank; batched values of stored iterates = the whole-matrix oracle
    try:
        eng.p2p_enable(True)
        dev = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=1e-9, label="mgd", tqdm=False,
                              device_loop=True, device_chunk=3)
        dev.run(P.x0, it_max=8)
        same = len(dev.history) == len(opt.history) and all(a == b for a, b in zip(dev.history, opt.history))
        same = same and torch.equal(dev.x, opt.x)
        errs["device_loop_equals_host_loop"] = 0.0 if same else 1.0
    except Exception as e:
        print(f"rank {rank}: device loop unavailable ({e})", flush=True)
    # d-sharded re-orthogonalisation (KCRN_DSHARD=1: every rank orthogonalises its slice of d; coefficients, |w'|^2 and
    # w' are sums over the ranks) against the replicated one, on a second handle over the same shard
    os.environ["KCRN_DSHARD"] = "1"
    loss2 = LogisticRegression(S.scipy_csr(), P_labels(S), l1=0, l2=0, store_mat_vec_prod=True, n_global=P.n)
    join_communicator(loss2.engine, rank, world)
    os.environ.pop("KCRN_DSHARD")
    res = []
    for eng_k in (eng, loss2.engine):
        eng_k.set_x(xd)
        gk = eng_k.gradient()
        eng_k.lanczos(12, reorth=True, g=gk)
        Vk, ak, bk, _ = eng_k.lanczos_view()
        res.append((ak.cpu().numpy().copy(), bk.cpu().numpy().copy(), Vk.cpu().numpy().copy()))
    G = res[1][2] @ res[1][2].T
    errs["dshard_reorth_vs_replicated"] = max(rel(res[1][0], res[0][0]), rel(res[1][1][:12], res[0][1][:12]), rel(res[1][2], res[0][2]))
    errs["dshard_orthogonality"] = float(np.max(np.abs(G - np.eye(12))))
    mine = torch.as_tensor(res[1][2], device="cuda")
    ref0 = mine.clone()
    dist.broadcast(ref0, 0)
    if not torch.equal(ref0, mine) or errs["dshard_reorth_vs_replicated"] > 1e-9 or errs["dshard_orthogonality"] > 1e-12:
        errs["peer_fail"] = 1.0
    loss2.engine.close()
    Xs = torch.stack([loss.to_device(x) for x in opt.trace.xs[:5]])
    vb = eng.values_batch(Xs).cpu().numpy()
    vo = np.array([L.value(x.cpu().numpy()) for x in opt.trace.xs[:5]])
    errs["values_batch"] = float(np.max(np.abs(vb - vo) / np.abs(vo)))
    g = np.load(os.path.join(HERE, "golden", f"traj_{name}.npz"))
    f = np.array([hh["f"] for hh in opt.history])
    errs["trajectory_prefix"] = float(np.max(np.abs(f - g["m10_ref_f"][:len(f)]) / np.abs(g["m10_ref_f"][:len(f)])))
    ok = (errs["value"] <= 1e-12 and errs["gradient"] <= 1e-12 and errs["hvp"] <= 1e-12
          and errs["replica_bits_equal"] == 0.0 and errs["trajectory_prefix"] <= 1e-8
          and "peer_fail" not in errs and "peer_replica_bits_equal" not in errs
          and errs.get("device_loop_equals_host_loop", 0.0) == 0.0 and errs["values_batch"] <= 1e-12)
    cnt = loss.engine.counters()
    print(f"rank {rank}/{world}: rows {S.n} nnz {S.nnz} errs {errs} allreduces {cnt['allreduces']} -> {'OK' if ok else 'FAIL'}",
          flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    loss.engine.close()            # graphs and the NCCL communicator go first
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


def P_labels(S):
    return S.b


if __name__ == "__main__":
    main()
