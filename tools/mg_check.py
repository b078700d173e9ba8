"""torchrun job: distributed fit + LML + gradient on WORLD_SIZE GPUs, checked on rank 0 against the oracle."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npoints", dest="n", type=int, default=3000)
    ap.add_argument("--block", dest="nb", type=int, default=256)
    ap.add_argument("--dim", dest="d", type=int, default=16)
    a = ap.parse_args()
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from gaussian_process_b200 import get_engine
    from gaussian_process_b200._lib import COV_SE
    from oracle import gp_oracle as O
    eng = get_engine(lr)
    eng.mg_init()
    X, y = O.synth_c5(a.n, a.d)
    lml, grad, alpha = eng.mg_fit_grad(COV_SE, X, y, [1.0, 4.0], 5e-4, nb=a.nb)
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        lml_o, grad_o, alpha_o = O.rbf_fit_lml_grad(X, y, 1.0, 4.0)
        e1 = abs(lml - lml_o) / abs(lml_o)
        e2 = abs(grad[1] - grad_o) / abs(grad_o)
        e3 = float(np.max(np.abs(eng.to_host(alpha[:a.n]) - alpha_o)) / np.max(np.abs(alpha_o)))
        print("world=%d lml rel %.2e grad rel %.2e alpha rel %.2e" % (dist.get_world_size(), e1, e2, e3))
        assert e1 < 1e-8 and e2 < 1e-7 and e3 < 1e-7
        print("MG_CHECK_OK")
    # ---- sharded prediction (test points split over ranks) from the REPLICATED factor of the distributed fit
    from gaussian_process_b200.distributed import BinaryLaplaceDistributed, multiclass_newton_sharded, predict_sharded
    Xc, yc, Xs = O.synth_c1(200, 333)
    np.random.seed(0)
    mu_o, sd_o, _ = O.regression_prediction(Xc, Xs, yc, 'rbf', 1, 1)
    fit = eng.mg_fit(COV_SE, Xc, yc, [1.0, 1.0], 5e-4, nb=128)
    for mu, var in (eng.mg_predict(fit, Xs), predict_sharded(eng, eng.fit(COV_SE, Xc, yc, [1.0, 1.0], 5e-4), Xs)):
        assert np.max(np.abs(mu - mu_o)) < 1e-8 * np.max(np.abs(mu_o)) and np.max(np.abs(var - sd_o ** 2)) < 1e-8 * np.max(sd_o ** 2)
    # ---- binary Laplace with B factored by the distributed Cholesky
    Xb, yb, _ = O.synth_c3(1500, 8)
    mb = BinaryLaplaceDistributed(eng, Xb, 1.0, 1.0, nb=128)
    itb = mb.fit_newton(yb, tolerance=1e-9)
    f_ob = O.binary_training_newton(O.rbf_kernel(Xb, Xb, 1, 1), yb, tolerance=1e-9)
    eb = float(np.max(np.abs(eng.to_host(mb.f[:1500]) - f_ob[0])) / np.max(np.abs(f_ob[0])))
    assert itb == f_ob[4] and eb < 1e-6, (itb, f_ob[4], eb)
    if dist.get_rank() == 0:
        print("distributed binary Laplace rel err %.2e after %d Newton steps" % (eb, itb))
        print("MG_LAPLACE_OK")
    Xm, labels, ym, Xt, tl = O.synth_c4(n=300, C=5, D=6, n_test=10)
    Xd = eng.to_device(Xm)
    Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
    model = multiclass_newton_sharded(eng, Kd, ym, 5, 300, tolerance=1e-9)
    p_o, f_o, it_o = O.multi_training_newton(O.rbf_kernel(Xm, Xm, 1, 1), ym, 5, 300, tolerance=1e-9)
    ef = float(np.max(np.abs(eng.to_host(model.f) - f_o)) / np.max(np.abs(f_o)))
    assert ef < 1e-6, ef
    if dist.get_rank() == 0:
        print("sharded prediction ok; class-sharded multiclass Laplace rel err %.2e (classes on rank 0: %s)" % (ef, model.classes))
        print("MG_SHARD_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
