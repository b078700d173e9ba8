"""torchrun job: class-sharded multiclass Laplace (C=10, n=8192, D=16) on WORLD_SIZE GPUs.
Class c lives on rank c mod P; sum_c E_c, R^T c and f are all-reduced.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    C, D = 10, 16
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from gaussian_process_b200 import get_engine
    from gaussian_process_b200._lib import COV_SE
    from oracle import gp_oracle as O
    eng = get_engine(lr)
    X, labels, y, Xt, tl = O.synth_c4(n, C, D, 2048)
    Xd = eng.to_device(X)
    Kd = eng.cov(COV_SE, Xd, Xd, [1.0, 1.0], same_x=True)
    from gaussian_process_b200 import parallel as P
    from gaussian_process_b200.laplace import MultiLaplaceNewton
    world, rank = dist.get_world_size(), dist.get_rank()
    # the model multiclass_newton_sharded builds, kept so that the warm-up pays for allocations / NCCL channels
    model = MultiLaplaceNewton(eng, Kd, C, n, classes=P.shard_classes(C, rank, world),
                               allreduce=(P.allreduce_sum_ if world > 1 else None))
    model.fit(y, 1e-6, 1)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    model.fit(y, 1e-6, 30)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=eng.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    it = len(model.errors)
    fm = model.predict(Xd, eng.to_device(Xt), y)
    acc = float(np.mean(np.argmax(fm, axis=1) == tl))
    if dist.get_rank() == 0:
        print(json.dumps({"config": "C4 multiclass Laplace C=%d n=%d D=%d, classes sharded over %d GPU(s)" % (C, n, D, dist.get_world_size()),
                          "fit_total_s": float(t.item()), "iterations": it, "per_iteration_s": float(t.item()) / it, "classes_on_rank0": list(model.classes),
                          "test_accuracy": acc}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
