import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, bench, osmo_trx_b200
trx = osmo_trx_b200.Trx(0)
n = 4096
rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", n, seed=7, device=trx.device)
trx.detect_config(16, 1)
r = trx.detect_demod(rx, typ, tsc, mt, bound, n_gmsk_soft=148)
torch.cuda.synchronize()
print("ok", int((r["rc"] > 0).sum()))
