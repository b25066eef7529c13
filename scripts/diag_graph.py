import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
import dsgan_oracle as O
from test_step_gpu import _make_model, _load
PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
batches = [O.synthetic_pair(2, 64, 64, seed=10 + i) for i in range(6)]
for prec in ("fp32", "bf16"):
    runs = {}
    for mode in ("eager1", "eager2", "graph"):
        model = _make_model(prec, ["--cuda_graph", "1" if mode == "graph" else "0"])
        _load(model, PG, PD, PV)
        traj = []
        for A, B in batches:
            model.set_input({"A": A, "B": B, "A_paths": [""], "B_paths": [""]})
            model.optimize_parameters()
            torch.cuda.synchronize()
            traj.append([float(model._loss[i]) for i in range(7)])
        runs[mode] = traj
    for step in range(6):
        d_ee = max(abs(a - b) for a, b in zip(runs["eager1"][step], runs["eager2"][step]))
        d_eg = max(abs(a - b) for a, b in zip(runs["eager1"][step], runs["graph"][step]))
        print(prec, step, "eager-eager %.2e   eager-graph %.2e" % (d_ee, d_eg), ["%.5f" % v for v in runs["graph"][step]])
